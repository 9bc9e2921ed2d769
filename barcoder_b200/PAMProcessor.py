"""PAMProcessor / GuideFinder / PAMFinder with the reference's surface (PAMProcessor.py:4-97).

String semantics are kept exactly, including the two quirks SURVEY.md A7/A8 records:
  * only `N` is expanded (to [ATCG]); any other letter of the PAM is a literal, and the PAM is
    used as a regular expression with re.search / re.finditer;
  * PAMFinder.get_pam_seq slices the same side for "upstream" and "downstream":
    '+' -> seq[End:End+P], '-' -> revcomp(seq[Start-P:Start]), with Python slice semantics at
    the contig ends (short or empty strings, which never match).

On the hot path these per-row functions are not called: constructing a PAMFinder registers it
as the active PAM, BowtieRunner.align() hands (pam, direction) to the CUDA search, and the hit
frame arrives with `PAM` and `Targeting` already filled (see CRISPRiLibrary._annotate_targets).
"""
import re

from ._state import ACTIVE_PAM
from .seqio import reverse_complement

_STRAND_WORDS = {"+": 1, "1": 1, "+1": 1, "fwd": 1, "forward": 1, "-": -1, "-1": -1, "rev": -1, "reverse": -1}


def _pam_regex(pam):
    return pam.replace("N", "[ATCG]")


class PAMProcessor:
    def __init__(self, records, pam, direction):
        self.records = records
        self.raw_pam = pam
        self.pam = _pam_regex(pam)
        self.direction = direction

    def get_sequence(self, row):
        sequence = self.records[row.Chromosome].seq[row.Start:row.End]
        return sequence.reverse_complement() if row.Strand == "-" else sequence

    def get_strand(self, strand_symbol):
        try:
            return _STRAND_WORDS[str(strand_symbol).lower().strip()]
        except KeyError:
            raise ValueError(f"Unrecognized strand symbol: {strand_symbol}") from None


class GuideFinder:
    """Guides next to every (non-overlapping) regex match of the PAM on both strands of every
    record; guides at a contig start may be shorter than `length` (PAMProcessor.py:27-57)."""

    def __init__(self, records, pam, direction, length):
        self.records = records
        self.pam = _pam_regex(pam)
        self.direction = direction
        self.length = length

    def find_guides_from_pam(self):
        if self.direction not in ("downstream", "upstream"):
            raise ValueError("Direction must be 'upstream' or 'downstream'")
        pattern = re.compile(self.pam)
        n = self.length
        guides = []
        for record in self.records.values():
            fwd = str(record.seq)
            for text in (fwd, reverse_complement(fwd)):
                if self.direction == "downstream":
                    guides.extend(text[max(0, m.start() - n):m.start()] for m in pattern.finditer(text))
                else:
                    guides.extend(text[m.end():m.end() + n] for m in pattern.finditer(text))
        return guides


class PAMFinder(PAMProcessor):
    def __init__(self, records, pam, direction):
        super().__init__(records, pam, direction)
        self.pam_length = len(pam)
        ACTIVE_PAM["finder"] = self

    @property
    def device_checkable(self):
        """True when the PAM is plain letters of at most 8 characters, i.e. the regex is exactly
        a per-position set test the kernel can evaluate."""
        return 0 < len(self.raw_pam) <= 8 and self.raw_pam.isalpha() and self.raw_pam.isupper()

    def get_pam_seq(self, row):
        if self.direction not in ("upstream", "downstream"):
            raise ValueError("direction must be 'upstream' or 'downstream'")
        seq = self.records[row.Chromosome].seq
        p = self.pam_length
        if self.get_strand(row.Strand) == 1:
            return str(seq[row.End:row.End + p])
        return str(seq[row.Start - p:row.Start].reverse_complement())

    def pam_matches(self, sequence):
        return bool(re.search(self.pam, sequence))
