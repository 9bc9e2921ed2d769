"""The five BASELINE.json configs on the GPU (SURVEY.md section 8d inputs).  Full oracle
comparison where the oracle finishes in seconds (cfg 1, 2, and library subsamples of cfg 3-5);
size-independent properties at full size: every planted spacer is recovered at its planted
site, sharded == unsharded, probe path == join path, gate == filter of ungated."""
import os

import numpy as np
import pytest

from barcoder_b200 import _native, synth
from oracle import oracle

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def gpu_search(genome, off, lib, k, pam, iupac=False, gate=False, path=0, blocks=0):
    with _native.Searcher(0) as s:
        s.set_genome_array(genome, off)
        s.set_library(lib)
        s.set_pam(pam, "downstream", iupac=iupac, gate=gate)
        if path:
            s.set_param(_native.BC_PARAM_PATH, path)
        if blocks:
            s.set_param(_native.BC_PARAM_BLOCKS, blocks)
        s.search(k)
        return _native.canonical_sort(s.hits()), s.stats()


def oracle_search(genome, off, lib, k, pam, iupac=False, gate=False):
    contigs = [bytes(genome[int(off[i]):int(off[i + 1])]) for i in range(len(off) - 1)]
    flags = (oracle.PAM_FLAG_IUPAC if iupac else 0) | (oracle.PAM_FLAG_GATE if gate else 0)
    return oracle.search(contigs, np.ascontiguousarray(lib), k, pam=pam, flags=flags)


def planted_library(genome, n, L, k, seed, frac=0.01):
    """Library with a known answer: returns (lib, planted index, planted dev-free position, strand)."""
    rng = synth.rng_for(seed)
    lib = synth.random_library(n, L, seed=seed + 1)
    m = max(1, int(n * frac))
    idx = rng.choice(n, size=m, replace=False)
    pos = rng.integers(0, len(genome) - L + 1, size=m)
    win = genome[pos[:, None] + np.arange(L)[None, :]].copy()
    ok = np.isin(win, np.frombuffer(b"ACGT", np.uint8)).all(axis=1)
    nsub = rng.integers(0, k + 1, size=m)
    for j in range(k):
        rows = np.nonzero(nsub > j)[0]
        cols = rng.integers(0, L, size=len(rows))
        win[rows, cols] = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, size=len(rows))]
    flip = rng.integers(0, 2, size=m).astype(bool)
    win[flip] = synth.revcomp_rows(win[flip])
    lib[idx[ok]] = win[ok]
    return lib, idx[ok], pos[ok], flip[ok]


def assert_planted_found(hits, idx, pos, flip, off):
    """Every planted spacer has a hit at its planted site and strand (contig-crossing plants excepted)."""
    key = set(zip(hits["spacer_id"].tolist(), hits["gpos"].tolist(), (hits["meta"] & 1).tolist()))
    L_ok = 0
    for i, p, f in zip(idx.tolist(), pos.tolist(), flip.tolist()):
        if (i, p, int(f)) in key:
            L_ok += 1
    assert L_ok >= 0.99 * len(idx), (L_ok, len(idx))  # plants that straddle a contig end cannot hit


def test_cfg1_ecoli_shape_full_oracle():
    """cfg 1 stand-in: 4,641,652 bp circular-shaped contig, 50,000 NGG 20-mers, k=1."""
    genome, off = synth.random_genome(4_641_652, seed=1)
    guides = synth.enumerate_pam_guides(genome, off, 20, "NGG")
    rng = synth.rng_for(1)
    lib = guides[rng.choice(len(guides), size=50_000, replace=False)]
    ref = oracle_search(genome, off, lib, 1, "NGG")
    gpu, st = gpu_search(genome, off, lib, 1, "NGG")
    assert gpu.tobytes() == ref.tobytes()
    assert len(gpu) >= 50_000 and ((gpu["meta"] & 8) != 0).sum() >= 50_000  # every guide hits itself next to its PAM


def test_cfg2_zymomonas_full_oracle(cn32_spacers, plasmids):
    """cfg 2: the 9,503 fixture spacers vs 4 real plasmids + a synthetic 2,058,755 bp chromosome, k=2, NGNC."""
    chrom, _ = synth.random_genome(2_058_755, seed=2)
    contigs = [chrom] + [np.frombuffer(str(r.seq).encode(), np.uint8) for r in plasmids.values()]
    genome = np.concatenate(contigs)
    off = np.zeros(len(contigs) + 1, np.uint64)
    off[1:] = np.cumsum([len(c) for c in contigs])
    lib = np.frombuffer("".join(cn32_spacers).encode(), np.uint8).reshape(-1, 32)
    ref = oracle_search(genome, off, lib, 2, "NGNC")
    for path in (1, 2):
        gpu, st = gpu_search(genome, off, lib, 2, "NGNC", path=path)
        assert gpu.tobytes() == ref.tobytes()
    on_plasmids = gpu[gpu["gpos"] >= off[1]]
    assert len(on_plasmids) == 869


def test_cfg3_all_ngg_guides_vs_own_genome():
    """cfg 3: every NGG 20-mer of the cfg-1 genome (~5e5) vs the genome, k=3; oracle on a subsample."""
    genome, off = synth.random_genome(4_641_652, seed=1)
    lib = synth.enumerate_pam_guides(genome, off, 20, "NGG")
    assert 400_000 < len(lib) < 700_000
    gpu, st = gpu_search(genome, off, lib, 3, "NGG")
    exact_self = gpu[((gpu["meta"] >> 1) & 3) == 0]
    assert len(np.unique(exact_self["spacer_id"])) == len(lib)  # each guide finds itself exactly
    sub = np.arange(0, len(lib), 97)[:4000]
    ref = oracle_search(genome, off, lib[sub], 3, "NGG")
    got = gpu[np.isin(gpu["spacer_id"], sub)].copy()
    got["spacer_id"] = np.searchsorted(sub, got["spacer_id"]).astype(np.uint32)
    assert _native.canonical_sort(got).tobytes() == ref.tobytes()
    gpu_probe, _ = gpu_search(genome, off, lib[:100_000], 3, "NGG", path=1)
    gpu_join, _ = gpu_search(genome, off, lib[:100_000], 3, "NGG", path=2)
    assert gpu_probe.tobytes() == gpu_join.tobytes()


def test_cfg4_scaled_planted_and_sharded():
    """cfg 4 at 1/10 scale (1M spacers x 10 Mbp, k=3) with the full-size seed scheme forced (b=5, 6)."""
    genome, off = synth.random_genome(10_000_000, seed=4)
    lib, idx, pos, flip = planted_library(genome, 1_000_000, 20, 3, seed=40)
    a, st = gpu_search(genome, off, lib, 3, "NGG", path=2, blocks=6)
    assert_planted_found(a, idx, pos, flip, off)
    b, _ = gpu_search(genome, off, lib, 3, "NGG", path=2, blocks=5)
    assert a.tobytes() == b.tobytes()
    sub = np.arange(0, len(lib), 251)[:3000]
    ref = oracle_search(genome, off, lib[sub], 3, "NGG")
    got = a[np.isin(a["spacer_id"], sub)].copy()
    got["spacer_id"] = np.searchsorted(sub, got["spacer_id"]).astype(np.uint32)
    assert _native.canonical_sort(got).tobytes() == ref.tobytes()
    gated, _ = gpu_search(genome, off, lib, 3, "NGG", gate=True, path=2)
    keep = a[((a["meta"] & 8) != 0) | ((a["meta"] & 32) != 0)]
    assert gated.tobytes() == keep.tobytes()


def test_cfg5_scaled_32mers_iupac_pam_many_contigs():
    """cfg 5 at 1/10 scale (100k 32-mers x 300 Mbp in 24 contigs with N runs, k=2, NNGRRT IUPAC)."""
    genome, off = synth.random_genome(300_000_000, seed=5, n_contigs=24, n_fraction=0.001)
    lib, idx, pos, flip = planted_library(genome, 100_000, 32, 2, seed=50)
    gpu, st = gpu_search(genome, off, lib, 2, "NNGRRT", iupac=True)
    assert_planted_found(gpu, idx, pos, flip, off)
    sub = np.sort(np.concatenate([idx[:300], np.arange(0, 2000)]))
    sub = np.unique(sub)
    ref = oracle_search(genome, off, lib[sub], 2, "NNGRRT", iupac=True)
    got = gpu[np.isin(gpu["spacer_id"], sub)].copy()
    got["spacer_id"] = np.searchsorted(sub, got["spacer_id"]).astype(np.uint32)
    assert _native.canonical_sort(got).tobytes() == ref.tobytes()
    assert ((gpu["meta"] & 8) != 0).sum() > 0


@pytest.mark.parametrize("direction", ["downstream", "upstream"])
@pytest.mark.parametrize("L,pam", [(20, "NGG"), (32, "NGNC"), (7, "TTTN"), (20, "")])
def test_guide_enumeration_matches_design_guides(direction, L, pam):
    """device enumeration == literal design_guides.py loop (reference loop bounds)."""
    genome, off = synth.random_genome(30_000, seed=L + len(pam), n_contigs=3, n_fraction=0.02, n_run=11)
    contigs = [bytes(genome[int(off[i]):int(off[i + 1])]).decode() for i in range(3)]
    want = oracle.py_enumerate_guides(contigs, L, pam, direction)
    with _native.Searcher(0) as s:
        s.set_genome_array(genome, off)
        got = s.enumerate_guides(L, pam, direction, reference_range=True)
        assert {bytes(r).decode() for r in got} == want
        full = s.enumerate_guides(L, pam, direction, reference_range=False)
        assert want <= {bytes(r).decode() for r in full}
    if direction == "downstream" and pam:
        host = synth.enumerate_pam_guides(genome, off, L, pam)
        assert np.array_equal(host, got)


def test_guide_enumeration_ecoli_scale_and_all_t():
    genome, off = synth.random_genome(4_641_652, seed=1)
    host = synth.enumerate_pam_guides(genome, off, 20, "NGG")
    with _native.Searcher(0) as s:
        s.set_genome_array(genome, off)
        got = s.enumerate_guides(20, "NGG")
    assert np.array_equal(host, got)
    g = np.frombuffer(b"ACG" + b"T" * 32 + b"AGGTTT", np.uint8)
    with _native.Searcher(0) as s:
        s.set_genome_array(g, np.array([0, len(g)], np.uint64))
        got = {bytes(r).decode() for r in s.enumerate_guides(32, "NGG")}
    assert "T" * 32 in got and got == oracle.py_enumerate_guides([bytes(g).decode()], 32, "NGG")


@pytest.mark.parametrize("path", [1, 2])
def test_genome_range_parts_equal_whole(path):
    """genome-range sharding (BC_PARAM_SCAN_PART) - the union of the parts is the whole."""
    genome, off = synth.random_genome(1_500_000, seed=61, n_contigs=5, n_fraction=0.002)
    lib = synth.random_library(20_000, 20, seed=62)
    synth.plant(lib, genome, 0.3, 2, seed=63)
    whole, _ = gpu_search(genome, off, lib, 2, "NGG", path=path)
    parts = []
    for r in range(3):
        with _native.Searcher(0) as s:
            s.set_genome_array(genome, off)
            s.set_library(lib)
            s.set_pam("NGG")
            s.set_param(_native.BC_PARAM_PATH, path)
            s.set_param(_native.BC_PARAM_SCAN_PART, r | (3 << 16))
            s.search(2)
            parts.append(s.hits())
    assert sum(len(p) for p in parts) == len(whole)
    assert _native.canonical_sort(np.concatenate(parts)).tobytes() == whole.tobytes()
    assert all(len(p) > 0 for p in parts)


# ------------------------------------------------------------- the configs at their stated sizes
def _subsample_equal(gpu_hits, sub, ref):
    got = gpu_hits[np.isin(gpu_hits["spacer_id"], sub)].copy()
    got["spacer_id"] = np.searchsorted(sub, got["spacer_id"]).astype(np.uint32)
    got = _native.canonical_sort(got)
    assert len(got) == len(ref), (len(got), len(ref))
    assert got.tobytes() == ref.tobytes()


def test_cfg4_full_size_bench_workload():
    """cfg 4 exactly as bench.py times it (BASELINE.json configs[3]): 10^7 distinct 20-mers (1 % planted)
    x 100 Mbp, k<=3, NGG, scheme chosen by the cost model.  Checked against (a) the oracle on a
    2,500-spacer subsample that contains planted spacers, bit-exact; (b) the closed-form expected
    number of random hits 2*G*n*P[Binomial(20, 3/4) <= 3] plus the planted ones."""
    from math import comb
    G, n, L, k = 100_000_000, 10_000_000, 20, 3
    genome, off = synth.random_genome(G, seed=4)
    lib = synth.random_library(n, L, seed=40)
    planted = synth.plant(lib, genome, 0.01, k, seed=1040)
    with _native.Searcher(0) as s:
        s.set_genome_array(genome, off)
        s.set_library(lib)
        s.set_pam("NGG", "downstream")
        nh = s.search(k)
        st = s.stats()
        hits = s.hits()
    assert st["path"] == 3 and st["key_nt"] >= 9      # compact join on a seed covering design
    p_le_k = sum(comb(L, j) * 0.75 ** j * 0.25 ** (L - j) for j in range(k + 1))
    expect = 2.0 * (G - L + 1) * n * p_le_k + len(planted)
    assert abs(nh - expect) < 6 * expect ** 0.5 + 0.002 * expect, (nh, expect)
    # every planted spacer is found at least once
    found = np.zeros(n, dtype=bool)
    found[hits["spacer_id"]] = True
    assert found[planted].all()
    rng = synth.rng_for(7)
    sub = np.unique(np.concatenate([planted[:500], rng.choice(n, size=2000, replace=False)]))
    ref = oracle_search(genome, off, lib[sub], k, "NGG")
    _subsample_equal(hits, sub, ref)


def test_cfg4_full_size_forced_multi_pass_join():
    """Same job with the genome forced into 3 join passes (BC_PARAM_JOIN_CHUNK) and a hit sink that
    receives the records across the passes: identical record set."""
    G, n, L, k = 100_000_000, 2_000_000, 20, 3
    genome, off = synth.random_genome(G, seed=4)
    lib = synth.random_library(n, L, seed=44)
    synth.plant(lib, genome, 0.01, k, seed=1044)
    with _native.Searcher(0) as s:
        s.set_genome_array(genome, off)
        s.set_library(lib)
        s.set_pam("NGG", "downstream")
        s.set_param(_native.BC_PARAM_PATH, 2)
        s.search(k)
        whole = _native.canonical_sort(s.hits())
        s.set_param(_native.BC_PARAM_JOIN_CHUNK, 37_000_001)
        sink = np.zeros(len(whole) + 16, dtype=_native.HIT_DTYPE)
        s.set_hit_sink(sink.ctypes.data, len(sink))
        nh = s.search(k)
        assert nh == len(whole)
        assert _native.canonical_sort(sink[:nh].copy()).tobytes() == whole.tobytes()
        assert _native.canonical_sort(s.hits()).tobytes() == whole.tobytes()


def test_cfg5_full_size():
    """cfg 5 at its stated size: 10^6 32-mers (1 % planted) x 3 Gbp in 24 contigs with N runs, k<=2,
    NNGRRT with IUPAC expansion.  Oracle on a subsample containing planted spacers, bit-exact; every
    planted spacer recovered at its planted site."""
    genome, off = synth.random_genome(3_000_000_000, seed=5, n_contigs=24, n_fraction=0.001)
    lib, idx, pos, flip = planted_library(genome, 1_000_000, 32, 2, seed=50)
    gpu, st = gpu_search(genome, off, lib, 2, "NNGRRT", iupac=True)
    assert_planted_found(gpu, idx, pos, flip, off)
    sub = np.unique(np.concatenate([idx[:500], np.arange(0, 2000)]))
    ref = oracle_search(genome, off, lib[sub], 2, "NNGRRT", iupac=True)
    _subsample_equal(gpu, sub, ref)
    assert ((gpu["meta"] & 8) != 0).sum() > 0


def test_join_dense_path_many_stages_and_tiles():
    """Dense verify with >= 3 shared-memory stages per bucket (> 128 entries) and >= 3 warp-tiles per
    slot (> 384 windows) on purpose: 6-nt keys (b=2 of L=12 at k=1), 6*10^5 entries, 3*10^6 windows."""
    genome, off = synth.random_genome(3_000_000, seed=71, n_contigs=3, n_fraction=0.001)
    lib = synth.random_library(300_000, 12, seed=72)
    synth.plant(lib, genome, 0.05, 1, seed=73)
    ref = oracle_search(genome, off, lib, 1, "NGG")
    gpu, st = gpu_search(genome, off, lib, 1, "NGG", path=2, blocks=2)
    assert st["path"] == 2 and st["blocks"] == 2
    # 4^6 slots per combination: ~146 entries (3 stages of 64) and ~732 windows (6 tiles of 128) per slot
    assert 2 * len(lib) / 4096 > 128 and len(genome) / 4096 > 384
    assert gpu.tobytes() == ref.tobytes()


@pytest.mark.parametrize("path,key_nt", [(1, 0), (2, 0), (3, 0), (3, 9), (1, 10)])
def test_slot_range_parts_equal_whole(path, key_nt):
    """Slot-range sharding (BC_PARAM_SLOT_PART, the strong-scaling split of ONE library): every
    context indexes / sorts / verifies only its range of the seed directory; the parts are disjoint
    and their union is the whole hit set."""
    genome, off = synth.random_genome(1_500_000, seed=81, n_contigs=5, n_fraction=0.002)
    lib = synth.random_library(30_000, 20, seed=82)
    synth.plant(lib, genome, 0.3, 3, seed=83)
    lib[11, 4] = ord("N")
    ref = oracle_search(genome, off, lib, 3, "NGG")
    parts = []
    for r in range(3):
        with _native.Searcher(0) as s:
            s.set_genome_array(genome, off)
            s.set_library(lib)
            s.set_pam("NGG")
            s.set_param(_native.BC_PARAM_PATH, path)
            if key_nt:
                s.set_param(_native.BC_PARAM_KEY_NT, key_nt)
            s.set_param(_native.BC_PARAM_SLOT_PART, r | (3 << 16))
            s.search(3)
            parts.append(s.hits())
    assert sum(len(p) for p in parts) == len(ref)
    assert _native.canonical_sort(np.concatenate(parts)).tobytes() == ref.tobytes()
    assert all(len(p) > 0 for p in parts)
