"""SAM text in the shape `bowtie -S` emits, and a pysam-free reader for it.

Writer: header @HD/@SQ/@PG; one line per alignment with FLAG 0/16 (+256 on every alignment of
a read after its first), 1-based leftmost POS, MAPQ 255, CIGAR <L>M, SEQ reverse-complemented
for '-' alignments, Q40 qualities, tags XA:i (mismatch stratum), MD:Z, NM:i; reads without an
alignment get a flag-4 line with XM:i:0.  That is what PySamParser.py:16-48 and
targets.py:522-530 of the reference consume through pysam.

Reader: `read_sam(path)` yields objects exposing the pysam.AlignedSegment attributes those two
call sites touch (is_reverse, is_unmapped, is_mapped, query_sequence, query_name, reference_name,
reference_start, reference_end, get_tag, has_tag, get_reference_sequence).
"""
from __future__ import annotations

import re

import numpy as np

from .seqio import reverse_complement

READ_NAME = "<unknown"  # SeqRecord(Seq(barcode)) has id "<unknown id>"; bowtie keeps the first token


def md_string(ref_window: str, mismatch_positions):
    """MD:Z for an ungapped alignment: match run lengths separated by the REFERENCE base at every
    mismatch (SAM spec); consecutive mismatches are separated by a 0."""
    out, run, mm = [], 0, set(mismatch_positions)
    for j, base in enumerate(ref_window):
        if j in mm:
            out.append(str(run))
            out.append(base)
            run = 0
        else:
            run += 1
    out.append(str(run))
    return "".join(out)


def write_sam(path, runner):
    """Serialise runner.hits (sorted per read, best stratum first) as bowtie-style SAM."""
    reads = [r.upper() for r in runner._reads]
    contigs, ids = runner._contigs, runner._contig_ids
    off = np.asarray(runner._offsets, dtype=np.int64)
    hits = runner.hits
    ci = np.searchsorted(off[1:], hits["gpos"].astype(np.int64), side="right")
    start = hits["gpos"].astype(np.int64) - off[ci]
    with open(path, "w") as h:
        h.write("@HD\tVN:1.0\tSO:unsorted\n")
        for name, seq in zip(ids, contigs):
            h.write(f"@SQ\tSN:{name}\tLN:{len(seq)}\n")
        h.write('@PG\tID:barcoder_b200\tVN:1\tCL:"bc_search -a --best -S"\n')
        j, n = 0, len(hits)
        for rid, read in enumerate(reads):
            L = len(read)
            first = True
            while j < n and hits["spacer_id"][j] == rid:
                meta = int(hits["meta"][j])
                minus = meta & 1
                nmm = (meta >> 1) & 3
                mask = int(hits["mm_mask"][j])
                s0 = int(start[j])
                window = contigs[ci[j]][s0:s0 + L].upper()
                pos = [i for i in range(L) if mask >> i & 1]
                if minus:  # mask is in spacer orientation; SAM fields are in reference orientation
                    pos = sorted(L - 1 - i for i in pos)
                flag = (16 if minus else 0) | (0 if first else 256)
                seq = reverse_complement(read) if minus else read
                h.write(f"{READ_NAME}\t{flag}\t{ids[ci[j]]}\t{s0 + 1}\t255\t{L}M\t*\t0\t0\t{seq}\t{'I' * L}\t"
                        f"XA:i:{nmm}\tMD:Z:{md_string(window, pos)}\tNM:i:{nmm}\n")
                first = False
                j += 1
            if first:
                h.write(f"{READ_NAME}\t4\t*\t0\t0\t*\t*\t0\t0\t{read}\t{'I' * L}\tXM:i:0\n")


_MD_TOKEN = re.compile(r"(\d+)|([A-Za-z])|\^([A-Za-z]+)")


class AlignedRead:
    __slots__ = ("query_name", "flag", "reference_name", "reference_start", "reference_end", "mapping_quality",
                 "cigarstring", "query_sequence", "tags")

    def __init__(self, fields):
        self.query_name = fields[0]
        self.flag = int(fields[1])
        unmapped = bool(self.flag & 4)
        self.reference_name = None if unmapped or fields[2] == "*" else fields[2]
        self.reference_start = -1 if unmapped else int(fields[3]) - 1
        self.mapping_quality = int(fields[4])
        self.cigarstring = None if fields[5] == "*" else fields[5]
        self.query_sequence = None if fields[9] == "*" else fields[9]
        span = sum(int(n) for n, op in re.findall(r"(\d+)([MIDNSHP=X])", fields[5]) if op in "MDN=X")
        self.reference_end = None if unmapped else self.reference_start + span
        self.tags = {}
        for t in fields[11:]:
            name, typ, val = t.split(":", 2)
            self.tags[name] = int(val) if typ == "i" else float(val) if typ == "f" else val

    @property
    def is_reverse(self):
        return bool(self.flag & 16)

    @property
    def is_unmapped(self):
        return bool(self.flag & 4)

    @property
    def is_mapped(self):
        return not self.flag & 4

    @property
    def is_secondary(self):
        return bool(self.flag & 256)

    def has_tag(self, name):
        return name in self.tags

    def get_tag(self, name):
        if name not in self.tags:
            raise KeyError(f"tag '{name}' not present")
        return self.tags[name]

    def get_reference_sequence(self):
        """Reference bases under the alignment rebuilt from SEQ + MD, lower-case at mismatches
        (pysam's convention, which is why design_guides.py:111 upper-cases `target`)."""
        md = self.tags.get("MD")
        if md is None or self.query_sequence is None:
            raise ValueError("MD tag not present")
        out, qi = [], 0
        for num, base, _deleted in _MD_TOKEN.findall(md):
            if num:
                out.append(self.query_sequence[qi:qi + int(num)])
                qi += int(num)
            elif base:
                out.append(base.lower())
                qi += 1
        return "".join(out)


def read_sam(path):
    with open(path) as h:
        for line in h:
            if not line or line[0] == "@":
                continue
            fields = line.rstrip("\n").split("\t")
            if len(fields) >= 11:
                yield AlignedRead(fields)
