"""Python face of the CPU oracle - TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module; the product package (barcoder_b200/) never does.

Parity pin: see oracle.c.  bowtie 1.3.1 (environment.yml:26) is absent, so the oracle
is pinned to the reference's data fixture (CN-32-zmo.tsv plasmid rows) and to the
SURVEY.md section 8c known answer, not to bowtie output: "parity unpinned" vs bowtie.

Contents
  * ctypes bindings for liboracle.so (orc_search_brute / orc_search_seeded /
    orc_annotate_pam), returning structured numpy arrays laid out like bc_hit;
  * a pure-Python twin (`py_search`, `py_pam_class_api`, `py_pam_script`) that follows
    the reference's string code literally, for tiny cases:
      PySamParser.py:26-48, PAMProcessor.py:65-97, targets.py:184-190, 219-307.
"""
from __future__ import annotations

import ctypes
import os
import re
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

HIT_DTYPE = np.dtype([("spacer_id", "<u4"), ("gpos", "<u4"), ("mm_mask", "<u4"), ("meta", "<u4")])

META_PAM_OK = 1 << 3
META_PAM_FULL = 1 << 4
META_PAM_AMB = 1 << 5
PAM_FLAG_IUPAC = 1
PAM_FLAG_GATE = 2


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True,
                       stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build())
        sig = [ctypes.c_char_p, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_char_p, ctypes.c_uint32,
               ctypes.c_uint32, ctypes.c_int, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int]
        for name in ("orc_search_brute", "orc_search_seeded"):
            f = getattr(L, name)
            f.argtypes = sig
            f.restype = ctypes.c_int64
        L.orc_search_seeded_b.argtypes = sig + [ctypes.c_int]
        L.orc_search_seeded_b.restype = ctypes.c_int64
        L.orc_annotate_pam.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_char_p, ctypes.c_void_p,
                                       ctypes.c_uint32, ctypes.c_uint32, ctypes.c_char_p, ctypes.c_int,
                                       ctypes.c_uint]
        L.orc_annotate_pam.restype = ctypes.c_int64
        _LIB = L
    return _LIB


def concat_genome(contigs):
    """list[str|bytes] -> (bytes, uint64 offsets[n+1])"""
    bs = [c.encode() if isinstance(c, str) else bytes(c) for c in contigs]
    off = np.zeros(len(bs) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([len(b) for b in bs])
    return b"".join(bs), off


def canonical_sort(hits):
    """Canonical order used by every parity comparison: (spacer_id, gpos, strand)."""
    order = np.lexsort((hits["meta"] & 1, hits["gpos"], hits["spacer_id"]))
    return hits[order]


def search(contigs, spacers, k, pam="", direction="downstream", flags=0, threads=None,
           mode="seeded", cap=None, blocks=0):
    """Run the C oracle.  `spacers`: list[str] of equal length, or a uint8 array [n, L] of ASCII
    rows.  Returns canonical-sorted hits.
    `blocks` forces the number of pigeonhole blocks of the seeded strategy (0 = cost model)."""
    genome, off = concat_genome(contigs)
    n = len(spacers)
    if isinstance(spacers, np.ndarray):   # uint8 ASCII rows [n, L]: no per-spacer Python strings
        L = spacers.shape[1] if n else 0
        sp = np.ascontiguousarray(spacers, dtype=np.uint8).tobytes()
    else:
        L = len(spacers[0]) if n else 0
        assert all(len(s) == L for s in spacers), "oracle.search needs equal-length spacers"
        sp = "".join(spacers).encode()
    threads = threads or os.cpu_count() or 1
    if mode == "seeded":
        def f(*a):
            return lib().orc_search_seeded_b(*a, int(blocks))
    else:
        f = lib().orc_search_brute
    cap = cap or max(1 << 16, 64 * n)
    while True:
        out = np.zeros(cap, dtype=HIT_DTYPE)
        total = f(genome, off.ctypes.data, len(contigs), sp, n, L, int(k), out.ctypes.data, cap, threads)
        if total < 0:
            raise ValueError("oracle rejected the arguments")
        if total <= cap:
            out = out[:total]
            break
        cap = int(total)
    kept = lib().orc_annotate_pam(out.ctypes.data, len(out), genome, off.ctypes.data, len(contigs), L,
                                  pam.encode(), 0 if direction == "downstream" else 1, flags)
    if kept < 0:
        raise ValueError("oracle rejected the PAM")
    return canonical_sort(out[:kept])


# ----------------------------------------------------------------------------- pure-Python twin

_COMP = str.maketrans("ACGTN", "TGCAN")


def revcomp(s):
    return s.translate(_COMP)[::-1]


def py_search(contigs, spacers, k):
    """Literal restatement for tiny inputs.  Returns a sorted list of tuples
    (spacer_index, contig_index, start0, strand, nmm, mismatch_positions_in_spacer_orientation)."""
    out = []
    for si, sp in enumerate(spacers):
        L = len(sp)
        if L <= k:
            continue
        for strand, q in (("+", sp), ("-", revcomp(sp))):
            for ci, ref in enumerate(contigs):
                for p in range(0, len(ref) - L + 1):
                    w = ref[p:p + L]
                    if any(ch not in "ACGT" for ch in w):
                        continue
                    mm = [j for j in range(L) if q[j] != w[j] or q[j] not in "ACGT"]
                    if len(mm) <= k:
                        pos = mm if strand == "+" else sorted(L - 1 - j for j in mm)
                        out.append((si, ci, p, strand, len(mm), tuple(pos)))
    return sorted(out)


def py_pam_class_api(seq, start, end, strand, pam):
    """PAMFinder.get_pam_seq + pam_matches (PAMProcessor.py:65-97): same slice for
    'upstream' and 'downstream'; python slicing with no bounds check; re.search with
    only N expanded to [ATCG]."""
    P = len(pam)
    if strand == "+":
        s = seq[end:end + P]
    else:
        s = revcomp(seq[start - P:start]) if True else ""
    return s, bool(re.search(pam.replace("N", "[ATCG]"), s))


def py_pam_script(seq, start, end, strand, pam, direction="downstream"):
    """extract_downstream_pam / extract_upstream_pam + pam_matches (targets.py:219-307).
    `seq` is the (topological) contig the bounds are checked against.  Returns
    (pam_or_None, matches)."""
    P = len(pam)
    right = (direction == "downstream") == (strand == "+")
    if right:
        if end + P > len(seq):
            return None, False
        s = seq[end:end + P].upper()
    else:
        if start - P < 0:
            return None, False
        s = seq[start - P:start].upper()
    if strand == "-":
        s = revcomp(s)
    if not s:
        return s, False
    if pam == "N" * len(pam) or not pam:
        return s, True
    return s, bool(re.match(pam.replace("N", "."), s))


def py_enumerate_guides(contigs, L, pam, direction="downstream"):
    """Literal restatement of design_guides.find_sequences_with_barcode_and_pam
    (design_guides.py:22-49) on plain strings: both strands, regex with N -> [ATGC] matched at the
    start of the slice, spacer must be pure GATC, result is a set.  Keeps the reference's loop
    bound (range(len - L - len(pam) + 1)) and its negative-index slices for the upstream PAM."""
    out = set()
    rx = re.compile(pam.replace("N", "[ATGC]"))
    for ref in contigs:
        for sequence in (ref, revcomp(ref)):
            for i in range(len(sequence) - L - len(pam) + 1):
                if direction == "downstream":
                    hit = rx.match(sequence[i + L:i + L + len(pam)])
                else:
                    hit = rx.match(sequence[i - len(pam):i])
                if hit:
                    spacer = sequence[i:i + L]
                    if all(b in "GATC" for b in spacer):
                        out.add(spacer)
    return out
