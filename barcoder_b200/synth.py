"""Synthetic genomes and libraries of the shapes BASELINE.json names.

RNG is ``numpy.random.Generator(PCG64(seed))`` throughout so every box generates the same
bytes.  Everything is returned as uint8 ASCII arrays (what the C ABI consumes).
"""
from __future__ import annotations

import numpy as np

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.zeros(256, dtype=np.uint8)
for _a, _b in zip(b"ACGTNacgtn", b"TGCANtgcan"):
    _COMP[_a] = _b


def rng_for(seed):
    return np.random.Generator(np.random.PCG64(seed))


def random_bases(n, rng):
    return _ACGT[rng.integers(0, 4, size=n, dtype=np.uint8)]


def random_genome(total_bp, seed, n_contigs=1, n_fraction=0.0, n_run=200):
    """-> (ascii uint8[total_bp], offsets uint64[n_contigs+1]).  `n_fraction` of the bases are
    overwritten with 'N' in runs of `n_run` (exercises the ambiguity plane)."""
    rng = rng_for(seed)
    seq = random_bases(total_bp, rng)
    if n_contigs == 1:
        offsets = np.array([0, total_bp], dtype=np.uint64)
    else:
        cuts = np.sort(rng.choice(np.arange(1, total_bp), size=n_contigs - 1, replace=False))
        offsets = np.concatenate([[0], cuts, [total_bp]]).astype(np.uint64)
    if n_fraction > 0:
        runs = max(1, int(total_bp * n_fraction / n_run))
        starts = rng.integers(0, max(1, total_bp - n_run), size=runs)
        for s in starts:
            seq[s:s + n_run] = ord("N")
    return seq, offsets


def codes_to_rows(codes, L):
    """uint64 base-4 codes (base j at bits 2j) -> uint8[n, L] ASCII."""
    out = np.empty((len(codes), L), dtype=np.uint8)
    for j in range(L):
        out[:, j] = _ACGT[((codes >> np.uint64(2 * j)) & np.uint64(3)).astype(np.uint8)]
    return out


def random_library(n, L, seed, distinct=True):
    """-> uint8[n, L] of uniform random spacers (distinct rows by default).  Drawn as base-4
    integers so that 10^7 rows dedupe with one integer sort."""
    rng = rng_for(seed)
    hi = (1 << (2 * L)) - 1
    if distinct and L <= 12 and n > hi + 1:
        distinct = False  # more rows requested than distinct spacers exist
    if distinct and n > 1 and L <= 12 and n > (hi + 1) // 4:
        codes = rng.permutation(np.arange(hi + 1, dtype=np.uint64))[:n]
        return np.ascontiguousarray(codes_to_rows(codes, L))
    codes = rng.integers(0, hi, size=n, dtype=np.uint64, endpoint=True)
    if distinct and n > 1:
        for _ in range(16):
            codes = np.sort(codes)
            codes = codes[np.concatenate([[True], codes[1:] != codes[:-1]])]
            if len(codes) == n or len(codes) > hi:
                break
            extra = rng.integers(0, hi, size=n - len(codes), dtype=np.uint64, endpoint=True)
            codes = np.concatenate([codes, extra])
        codes = rng.permutation(codes[:n])
    return np.ascontiguousarray(codes_to_rows(codes, L))


def revcomp_rows(rows):
    return _COMP[rows[:, ::-1]]


def plant(library, genome, fraction, k, seed):
    """Overwrite `fraction` of the library rows with genome windows (random strand) carrying
    0..k random substitutions, so that the expected hit set is non-trivial.  Windows touching a
    non-ACGT base are skipped.  Returns the indices that were planted."""
    rng = rng_for(seed)
    n, L = library.shape
    m = int(n * fraction)
    if m == 0 or len(genome) < L:
        return np.zeros(0, dtype=np.int64)
    idx = rng.choice(n, size=m, replace=False)
    pos = rng.integers(0, len(genome) - L + 1, size=m)
    win = genome[pos[:, None] + np.arange(L)[None, :]].copy()
    ok = np.isin(win, _ACGT).all(axis=1)
    nsub = rng.integers(0, k + 1, size=m)
    for j in range(k):
        rows = np.nonzero(nsub > j)[0]
        cols = rng.integers(0, L, size=len(rows))
        win[rows, cols] = _ACGT[rng.integers(0, 4, size=len(rows), dtype=np.uint8)]
    flip = rng.integers(0, 2, size=m).astype(bool)
    win[flip] = revcomp_rows(win[flip])
    library[idx[ok]] = win[ok]
    return idx[ok]


def enumerate_pam_guides(genome, offsets, L, pam="NGG"):
    """Every distinct ACGT-only L-mer immediately 5' of a PAM match on either strand, per
    contig, linear (design_guides.py:22-49 semantics without the circular overhang).
    Only `N` is a wildcard in `pam`.  -> uint8[n, L] sorted."""
    P = len(pam)
    out = []
    for c in range(len(offsets) - 1):
        seq = genome[int(offsets[c]):int(offsets[c + 1])]
        for strand_seq in (seq, _COMP[seq[::-1]]):
            n = len(strand_seq)
            if n < L + P:
                continue
            ok = np.ones(n - L - P + 1, dtype=bool)
            for i, ch in enumerate(pam.encode()):
                col = strand_seq[L + i:n - P + 1 + i]
                ok &= np.isin(col, _ACGT) if ch == ord("N") else (col == ch)
            starts = np.nonzero(ok)[0]
            if len(starts) == 0:
                continue
            win = strand_seq[starts[:, None] + np.arange(L)[None, :]]
            win = win[np.isin(win, _ACGT).all(axis=1)]
            out.append(win)
    if not out:
        return np.zeros((0, L), dtype=np.uint8)
    allw = np.ascontiguousarray(np.concatenate(out))
    view = allw.view(np.dtype((np.void, L))).ravel()
    return np.unique(view).view(np.uint8).reshape(-1, L)


def rows_to_strings(rows):
    return [bytes(r).decode() for r in rows]
