"""BowtieRunner - the reference's aligner front-end, backed by the CUDA search instead of bowtie.

Public surface kept from BowtieRunner.py:13-150 of the reference:

    with BowtieRunner() as bowtie:
        bowtie.make_fasta(records)        # dict id -> record with .id/.seq/.description
        bowtie.make_fastq(barcodes)       # iterable of spacer strings; calls accumulate (:67 append mode)
        bowtie.create_index()             # BowtieError if no genome was written (:79-80)
        bowtie.align(num_mismatches=1, num_threads=12)
        sam = PySamParser(bowtie.sam_path)

plus the properties index_path / fasta_path / fastq_path / sam_path and BowtieError(message).

What happens instead of the two subprocesses (BowtieRunner.py:87-97,111-136):
  make_fasta    keeps the contigs in memory (and writes the FASTA file, as the reference does)
  create_index  bc_create + bc_set_genome: H2D copy and 2-bit/ambiguity packing on the GPU
  align         per spacer length: bc_set_library, bc_build_index(k), bc_search(k), bc_copy_hits;
                `bowtie -a -v k` semantics (every alignment, both strands), PAM fused when known
The result is kept as a hit table (`.hits`, `.frame`) registered under `sam_path`, so
PySamParser(bowtie.sam_path).ranges needs no text round trip; a bowtie-style SAM file is also
written to `sam_path` (always when write_sam=True, by default only up to SAM_AUTO_LIMIT lines).

Several GPUs: `BowtieRunner(devices=[0, 1, ...])` (or devices="all") keeps one search context per GPU,
each driven by its own host thread; every GPU holds the genome and the library and owns one slot
range of the seed directory (BC_PARAM_SLOT_PART), so index build, window sort and verification all
split N ways and the per-GPU hit sets are disjoint.  With devices="auto" the reference's
`num_threads` argument of align() (bowtie -p, BowtieRunner.py:104,120) picks how many GPUs are used.
There is no CPU fallback: any native error surfaces as BowtieError.
"""
from __future__ import annotations

import os
import tempfile

import numpy as np
import pandas as pd

from . import _native, samio
from ._state import ACTIVE_PAM, RESULTS
from .Logger import Logger
from .seqio import reverse_complement, write_fasta

SAM_AUTO_LIMIT = 2_000_000

try:  # string columns of 10^5..10^6 rows: Arrow kernels instead of one Python object per row
    import pyarrow as _pa
    import pyarrow.compute as _pc
except ImportError:  # pragma: no cover - pandas works without it, only slower
    _pa = _pc = None


def _upper_and_lengths(reads):
    """(upper-cased reads, their lengths as int64).  With pyarrow the first is an Arrow string array, else a list."""
    if _pa is not None:
        try:
            arr = _pa.array(reads, type=_pa.string())
            return _pc.utf8_upper(arr), _pc.utf8_length(arr).to_numpy(zero_copy_only=False).astype(np.int64)
        except (_pa.ArrowInvalid, _pa.ArrowTypeError, TypeError):
            pass
    up = [r.upper() for r in reads]
    return up, np.fromiter((len(r) for r in up), dtype=np.int64, count=len(up))


def _take_str(values, idx):
    """values[idx] as a pandas string column; values = an Arrow string array, or any sequence of str."""
    if _pa is not None:
        try:
            arr = values if isinstance(values, (_pa.Array, _pa.ChunkedArray)) else _pa.array(list(values), type=_pa.string())
            return pd.Series(arr.take(_pa.array(np.asarray(idx, dtype=np.int64))), dtype="str").array
        except (_pa.ArrowInvalid, _pa.ArrowTypeError, TypeError, ValueError):
            pass
    vals = np.asarray(values.to_pylist() if _pa is not None and isinstance(values, (_pa.Array, _pa.ChunkedArray)) else list(values),
                      dtype=object)
    return vals[np.asarray(idx, dtype=np.int64)]


def _rows_of(upper, idx, L):
    """The spacers upper[idx] (all of length L) as a uint8 matrix [n, L] for bc_set_library."""
    if _pa is not None and isinstance(upper, (_pa.Array, _pa.ChunkedArray)):
        tk = upper.take(_pa.array(np.asarray(idx, dtype=np.int64)))
        if isinstance(tk, _pa.ChunkedArray):
            tk = tk.combine_chunks()
        data = tk.buffers()[2]
        if tk.offset == 0 and data is not None and data.size >= len(idx) * L:
            flat = np.frombuffer(data, dtype=np.uint8, count=len(idx) * L)
            if tk.null_count == 0:
                return flat.reshape(len(idx), L)
        return [s for s in tk.to_pylist()]
    return [upper[i] for i in idx]



class BowtieError(Exception):
    """Raised for any failure of the search path; `.message` explains it (BowtieRunner.py:144-150)."""

    def __init__(self, message):
        super().__init__(message)
        self.message = message


class BowtieRunner(Logger):
    def __init__(self, device=0, write_sam="auto", write_files=True, devices=None):
        super().__init__()
        self.temp_dir = tempfile.TemporaryDirectory()
        self.device = device
        self.devices = devices          # None: [device]; list of ordinals; "all"; "auto" (num_threads decides)
        self._searchers = []
        self.write_sam = write_sam
        self.write_files = write_files
        self._index_path = None
        self._contig_ids, self._contigs = [], []
        self._reads = []
        self._upper = self._lens = None   # upper-cased reads (Arrow array or list) and their lengths, built by align()
        self._searcher = None
        self._pam = None  # (pam, direction) set explicitly through set_pam()
        self.hits = None
        self.frame = None
        self.stats = []

    # ---- paths (same names as the reference; files live in the temp dir)
    @property
    def index_path(self):
        if self._index_path is None:
            fd, self._index_path = tempfile.mkstemp(dir=self.temp_dir.name)
            os.close(fd)
        return self._index_path

    @property
    def fasta_path(self):
        return self.index_path + ".fasta"

    @property
    def fastq_path(self):
        return self.index_path + ".fastq"

    @property
    def sam_path(self):
        return self.index_path + ".sam"

    def __enter__(self):
        return self

    def __exit__(self, exc_type, exc_val, exc_tb):
        self.close()

    def close(self):
        RESULTS.pop(self.sam_path, None)
        for srch in self._searchers:
            srch.close()
        self._searchers = []
        self._searcher = None
        self.temp_dir.cleanup()

    # ---- inputs
    def make_fasta(self, records):
        self.info(f"Writing FASTA file {self.fasta_path} ...")
        try:
            recs = list(records.values())
            self._contig_ids = [r.id for r in recs]
            self._contigs = [str(r.seq) for r in recs]
            if self.write_files:
                with open(self.fasta_path, "w") as handle:
                    write_fasta(recs, handle)
        except Exception as exc:
            raise BowtieError("Failed to write FASTA file") from exc

    def make_fastq(self, barcodes):
        self.info(f"Writing FASTQ file {self.fastq_path} ...")
        try:
            new = [str(b) for b in barcodes]
            self._reads.extend(new)
            if self.write_files:
                with open(self.fastq_path, "a") as handle:
                    for s in new:  # Q40 for every base, unnamed records (BowtieRunner.py:68-74)
                        handle.write(f"@{samio.READ_NAME} <unknown description>\n{s}\n+\n{'I' * len(s)}\n")
        except Exception as exc:
            raise BowtieError("Failed to write FASTQ file") from exc

    def set_pam(self, pam, direction="downstream"):
        """Optional: tell the aligner which PAM to check in the same pass.  Without it, align()
        uses the most recently constructed PAMFinder, if any."""
        self._pam = (pam, direction)

    # ---- compute
    def _device_list(self):
        if self.devices is None:
            return [self.device]
        if isinstance(self.devices, str):
            n = _native.device_count()
            if n == 0:
                raise BowtieError("no usable CUDA device")
            return list(range(n))           # "all" / "auto": align() may use fewer ("auto" + num_threads)
        return list(self.devices)

    def create_index(self):
        if not self._contigs or sum(len(c) for c in self._contigs) == 0:
            raise BowtieError("BowtieRunner.fasta_path does not exist or is an empty")
        devs = self._device_list()
        self.info(f"Creating index on CUDA device(s) {devs} ...")
        try:
            if not self._searchers:
                self._searchers = [_native.Searcher(d) for d in devs]
                self._searcher = self._searchers[0]
            self._on_all(self._searchers, lambda srch, _i: srch.set_genome(self._contigs))
        except (_native.NativeError, _native.NativeLibraryError) as exc:
            raise BowtieError("Failed to index") from exc
        st = self._searcher.stats()
        self.subproc(f"packed {st['genome_bases']} bases in {st['ms_pack_genome']:.3f} ms on {len(devs)} GPU(s)")

    @staticmethod
    def _on_all(searchers, fn):
        """fn(searcher, index) on every context, one host thread per GPU (the C calls release the GIL)."""
        if len(searchers) == 1:
            return [fn(searchers[0], 0)]
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(len(searchers)) as ex:
            return list(ex.map(lambda a: fn(a[1], a[0]), enumerate(searchers)))

    def _pam_setting(self):
        if self._pam is not None:
            return self._pam
        finder = ACTIVE_PAM["finder"]
        if finder is not None and finder.device_checkable:
            # the class API slices the 3' side for both directions (PAMProcessor.py:69-87)
            return (finder.raw_pam, "downstream")
        return None

    def align(self, num_mismatches=0, num_threads=os.cpu_count()):
        if self._searcher is None:
            raise BowtieError("bowtie failed: create_index() has not been called")
        k = int(num_mismatches)
        self.info(f"Performing alignment on {len(self._searchers)} CUDA device(s) ...")
        self.json(["bc_search", "-a", f"-v{k}", f"reads={len(self._reads)}"])
        pam = self._pam_setting()
        reads = self._reads
        upper, lens = _upper_and_lengths(reads)
        self._upper, self._lens = upper, lens
        by_len = {int(L): np.nonzero(lens == L)[0] for L in np.unique(lens)}
        parts, self.stats = [], []
        active = self._searchers
        if self.devices == "auto" and num_threads:
            active = active[:max(1, min(len(active), int(num_threads)))]
        world = len(active)
        try:
            for L, idx in sorted(by_len.items()):
                if L == 0 or L > 32:
                    if L > 32:
                        raise BowtieError(f"spacers longer than 32 nt are not supported (got {L})")
                    continue
                idx = np.asarray(idx, dtype=np.int64)
                lib = _rows_of(upper, idx, L)

                def one(srch, rank):
                    srch.set_pam(pam[0] if pam else "", pam[1] if pam else "downstream")
                    srch.set_library(lib)
                    srch.set_param(_native.BC_PARAM_SLOT_PART, rank | (world << 16))
                    srch.search(k)
                    srch.sort_hits("best")   # bowtie --best: per read, fewest mismatches first (device radix sort)
                    return srch.hits(), srch.stats()

                for h, st in self._on_all(active, one):
                    h["spacer_id"] = idx[h["spacer_id"]].astype(np.uint32)  # back to read order
                    parts.append(h)
                    self.stats.append(st)
        except (_native.NativeError, _native.NativeLibraryError) as exc:
            raise BowtieError(f"bc_search failed: {exc}") from exc
        hits = np.concatenate(parts) if parts else np.zeros(0, dtype=_native.HIT_DTYPE)
        # bowtie --best: per read, fewest mismatches first; then by position for determinism.  One GPU and
        # one spacer length: the device already sorted the buffer; otherwise the sorted parts are merged
        if len(parts) > 1:
            order = np.lexsort((hits["meta"] & 1, hits["gpos"], (hits["meta"] >> 1) & 3, hits["spacer_id"]))
            hits = hits[order]
        self.hits = hits
        self._pam_used = pam
        self.frame = self._build_frame()
        RESULTS[self.sam_path] = self
        n_lines = len(self.frame)
        if self.write_sam is True or (self.write_sam == "auto" and n_lines <= SAM_AUTO_LIMIT):
            samio.write_sam(self.sam_path, self)
        for st in self.stats:
            self.subproc(f"L={st['spacer_len']} k={st['k']}: {st['hits']} alignments of {st['library_spacers']} "
                         f"reads, search {st['ms_search']:.3f} ms (index {st['ms_build_index']:.3f} ms)")

    # ---- result views
    def contig_of(self, gpos):
        off = self._searcher.contig_offsets if self._searcher is not None else self._offsets
        return np.searchsorted(off[1:], gpos, side="right")

    def _build_frame(self):
        """Hit table with the columns PySamParser.ranges produces (PySamParser.py:38-46), one row
        per alignment plus one unmapped row per read without alignments, and - when a PAM was
        fused - `PAM` / `Targeting` columns (CRISPRiLibrary.py:16-21)."""
        hits, reads = self.hits, self._reads
        off = np.asarray(self._searcher.contig_offsets, dtype=np.int64)
        self._offsets = off
        if getattr(self, "_upper", None) is None or len(self._lens) != len(reads):
            self._upper, self._lens = _upper_and_lengths(reads)
        barcode, L = self._upper, self._lens
        sid = hits["spacer_id"].astype(np.int64)
        ci = np.searchsorted(off[1:], hits["gpos"].astype(np.int64), side="right")
        start = hits["gpos"].astype(np.int64) - off[ci]
        df = pd.DataFrame({
            "Chromosome": _take_str(self._contig_ids, ci) if len(hits) else np.zeros(0, dtype=object),
            "Start": start,
            "End": start + L[sid] if len(hits) else start,
            "Mapped": np.ones(len(hits), dtype=bool),
            "Strand": _take_str(["+", "-"], hits["meta"] & 1) if len(hits) else np.zeros(0, dtype=object),
            "Barcode": _take_str(barcode, sid) if len(hits) else np.zeros(0, dtype=object),
            "Mismatches": ((hits["meta"] >> 1) & 3).astype(np.int64),
        })
        if self._pam_used:
            pam_str, targeting = self._pam_columns(ci, start, L[sid] if len(hits) else L[:0])
            df["PAM"] = pam_str
            df["Targeting"] = targeting
            # "class-api" = the 3' slice PAMFinder.get_pam_seq uses for BOTH directions
            # (PAMProcessor.py:69-87); an explicit set_pam(pam, "upstream") follows the script rule
            # (5' side, targets.py:266-307) instead, and CRISPRiLibrary then re-annotates with its finder
            df.attrs["pam_key"] = (self._pam_used[0].upper(),
                                   "class-api" if self._pam_used[1] == "downstream" else "upstream")
        aligned = np.zeros(len(reads), dtype=bool)
        aligned[sid] = True
        missing = np.nonzero(~aligned)[0]
        if len(missing):  # flag-4 SAM lines: PySamParser reports them with Mismatches "0" (:45)
            um = pd.DataFrame({
                "Chromosome": None, "Start": -1, "End": None, "Mapped": False, "Strand": "+",
                "Barcode": _take_str(barcode, missing), "Mismatches": "0",
            })
            if self._pam_used:
                um["PAM"] = ""
                um["Targeting"] = False
            attrs = dict(df.attrs)
            df = pd.concat([df.astype({"Mismatches": object, "End": object}), um], ignore_index=True)
            df.attrs.update(attrs)
        return df

    def _pam_columns(self, ci, start, lens):
        """PAM strings and match flags.  Full, unambiguous PAMs come decoded from the kernel's
        2-bit codes; truncated or ambiguous ones (contig ends, N in the genome) and lower-case
        genomes are resolved with the reference's own string rule (PAMProcessor.py:65-97)."""
        import re
        hits = self.hits
        pam = self._pam_used[0].upper()
        P = len(pam)
        meta = hits["meta"]
        codes = (meta >> 16).astype(np.uint32)
        if P:
            # at most 4^P distinct PAMs: decode each distinct code once, then index (one Python
            # string per distinct PAM instead of one per hit)
            uniq, inv = np.unique(codes & np.uint32((1 << (2 * P)) - 1), return_inverse=True)
            table = np.empty(len(uniq), dtype=object)
            for t, code in enumerate(uniq.tolist()):
                table[t] = "".join("ACGT"[(code >> (2 * i)) & 3] for i in range(P))
        else:
            table, inv = np.asarray([""], dtype=object), np.zeros(len(hits), dtype=np.int64)
        targeting = (meta & _native.META_PAM_OK) != 0
        lower = any(c != c.upper() for c in self._contigs) if not hasattr(self, "_has_lower") else self._has_lower
        self._has_lower = lower
        slow = ((meta & _native.META_PAM_FULL) == 0) | ((meta & _native.META_PAM_AMB) != 0)
        if lower:
            slow = np.ones(len(hits), dtype=bool)
        if not slow.any():
            return (_take_str(table, inv) if len(hits) else np.zeros(0, dtype=object)), targeting
        pam_str = table[inv] if len(hits) else np.zeros(0, dtype=object)
        pattern = re.compile(pam.replace("N", "[ATCG]"))
        minus = (meta & 1) != 0
        upstream = self._pam_used[1] == "upstream"
        for j in np.nonzero(slow)[0]:
            seq = self._contigs[ci[j]]
            s0, e0 = int(start[j]), int(start[j] + lens[j])
            if upstream:  # 5' side of the protospacer; a PAM that leaves the contig is no PAM (targets.py:266-307)
                if minus[j]:
                    s = reverse_complement(seq[e0:e0 + P]) if e0 + P <= len(seq) else ""
                else:
                    s = seq[s0 - P:s0] if s0 - P >= 0 else ""
            else:
                s = reverse_complement(seq[s0 - P:s0]) if minus[j] else seq[e0:e0 + P]
            pam_str[j] = s
            targeting[j] = bool(pattern.search(s))
        return pam_str, targeting
