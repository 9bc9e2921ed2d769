"""PySamParser - the reference's SAM -> hit-frame step (PySamParser.py:8-52) without pysam.

`PySamParser(filename).ranges` yields the same columns in the same meaning:
Chromosome, Start (0-based leftmost '+'-strand position), End, Mapped, Strand, Barcode (always
the library spacer: SEQ is reverse-complemented back for '-' alignments), Mismatches (NM, or
the string "0" when the tag is absent).

If `filename` is the sam_path of a BowtieRunner that has aligned in this process, the frame
comes straight from its device hit records - no SAM text is parsed.  Any other SAM file is read
with the text reader in samio.py, row by row like the reference does.
"""
import pandas as pd

from . import samio
from ._state import RESULTS
from .ranges import PyRanges
from .seqio import reverse_complement


def rev_comp(seq):
    return reverse_complement(seq)


class PySamParser:
    def __init__(self, filename):
        self.filename = filename
        self._ranges = None

    def read_sam(self):
        yield from samio.read_sam(self.filename)

    def _rows_from_text(self):
        cols = {k: [] for k in ("Chromosome", "Start", "End", "Mapped", "Strand", "Barcode", "Mismatches")}
        for read in self.read_sam():
            minus = read.is_reverse
            seq = read.query_sequence
            cols["Chromosome"].append(read.reference_name)
            cols["Start"].append(read.reference_start)
            cols["End"].append(read.reference_end)
            cols["Mapped"].append(not read.is_unmapped)
            cols["Strand"].append("-" if minus else "+")
            cols["Barcode"].append(rev_comp(seq) if (minus and seq is not None) else seq)
            cols["Mismatches"].append(read.get_tag("NM") if read.has_tag("NM") else "0")
        return pd.DataFrame(cols)

    @property
    def ranges(self):
        if self._ranges is None:
            runner = RESULTS.get(self.filename)
            df = runner.frame if runner is not None and runner.frame is not None else self._rows_from_text()
            pr = PyRanges(df)
            pr.df.attrs.update(getattr(df, "attrs", {}))
            self._ranges = pr
        return self._ranges
