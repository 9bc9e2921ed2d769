"""Host-side library input for the search path.

Mirrors the public surface of the reference's BarCodeLibrary module (BarCodeLibrary.py:9-102):
`BarCodeLibraryReader(filename, column=...).read_barcodes()`, `BarCodeLibrary(filename=None,
barcodes=None, **kwargs)` with `.barcodes` (a de-duplicated, unordered set of strings), `.size`,
`.add/.remove/.load/.load_from_list`, and `BarCodeLibraryError`.
"""
import csv
import os

from .Logger import Logger
from . import seqio


class BarCode:
    def __init__(self, sequence):
        self.sequence = sequence


def _spacers_from_fasta(path, _column):
    return [str(rec.seq) for rec in seqio.read_fasta(path)]


def _spacers_from_tsv(path, column):
    if column is None:
        raise ValueError("A barcode column must be specified for TSV files")
    with open(path, newline="") as handle:
        rows = csv.reader(handle, delimiter="\t")
        header = next(rows)
        try:
            at = header.index(column)
        except ValueError:
            raise ValueError(f"Column '{column}' not found in file") from None
        return [row[at] for row in rows]


_READERS = {".fasta": _spacers_from_fasta, ".tsv": _spacers_from_tsv}


class BarCodeLibraryReader:
    def __init__(self, filename, column=None):
        self.filename = filename
        self.column = column

    def read_barcodes(self):
        ext = os.path.splitext(self.filename)[1]
        if ext not in _READERS:
            raise ValueError(f"Unsupported file format: {self.filename}")
        return _READERS[ext](self.filename, self.column)


class BarCodeLibraryError(Exception):
    """Raised when a library cannot be loaded; `.message` explains why."""

    def __init__(self, message):
        super().__init__(message)
        self.message = message


class BarCodeLibrary(Logger):
    def __init__(self, filename=None, barcodes=None, **kwargs):
        super().__init__()
        self._barcodes = set()
        self.kwargs = kwargs
        self.reader = None
        if filename is not None:
            self.reader = BarCodeLibraryReader(filename, **kwargs)
            self.load()
        if barcodes is not None:
            self.load_from_list(barcodes)

    @property
    def barcodes(self):
        return self._barcodes

    @property
    def size(self):
        return len(self._barcodes)

    def add(self, sequence):
        self._barcodes.add(sequence)

    def remove(self, sequence):
        self._barcodes.remove(sequence)

    def load(self):
        # the class path searches linear contigs (BarCodeLibrary.py:73-75)
        self.warn("Genome circularity is not yet implemented. Barcodes spanning the origin will be missed!")
        try:
            self._barcodes.update(self.reader.read_barcodes())
        except Exception as exc:
            raise BarCodeLibraryError("Failed to load barcodes") from exc
        self.info(f"Loaded {self.size} barcodes from {os.path.abspath(self.reader.filename)} ...")

    def load_from_list(self, barcodes):
        self._barcodes.update(barcodes)
        self.info(f"Loaded {self.size} barcodes from list ...")
