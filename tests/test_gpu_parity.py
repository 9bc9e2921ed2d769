"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle on the
same seeded inputs.  Bit-exact: the canonical-sorted 16-byte hit records must be identical."""
import os

import numpy as np
import pytest

from barcoder_b200 import _native, synth
from oracle import oracle

pytestmark = pytest.mark.gpu


def run_gpu(genome, off, lib, k, pam="", direction="downstream", iupac=False, gate=False, blocks=0, path=0,
            hit_cap=0, window_sort=0, key_nt=0, join_chunk=0):
    with _native.Searcher(0) as s:
        s.set_genome_array(genome, off)
        s.set_library(lib)
        s.set_pam(pam, direction, iupac=iupac, gate=gate)
        if blocks:
            s.set_param(_native.BC_PARAM_BLOCKS, blocks)
        if path:
            s.set_param(_native.BC_PARAM_PATH, path)
        if hit_cap:
            s.set_param(_native.BC_PARAM_HIT_CAPACITY, hit_cap)
        if window_sort:
            s.set_param(_native.BC_PARAM_WINDOW_SORT, window_sort)
        if key_nt:
            s.set_param(_native.BC_PARAM_KEY_NT, key_nt)
        if join_chunk:
            s.set_param(_native.BC_PARAM_JOIN_CHUNK, join_chunk)
        n = s.search(k)
        hits = s.hits()
        assert len(hits) == n
        return _native.canonical_sort(hits), s.stats()


def run_oracle(genome, off, lib, k, pam="", direction="downstream", iupac=False, gate=False, mode="seeded"):
    contigs = [bytes(genome[int(off[i]):int(off[i + 1])]) for i in range(len(off) - 1)]
    spacers = synth.rows_to_strings(lib)
    flags = (oracle.PAM_FLAG_IUPAC if iupac else 0) | (oracle.PAM_FLAG_GATE if gate else 0)
    return oracle.search(contigs, spacers, k, pam=pam, direction=direction, flags=flags, mode=mode)


def assert_same(gpu, ref):
    assert len(gpu) == len(ref), (len(gpu), len(ref))
    assert gpu.tobytes() == ref.tobytes()


def small_case(L, k, seed, n=300, G=60000, n_contigs=4, nfrac=0.01):
    genome, off = synth.random_genome(G, seed=seed, n_contigs=n_contigs, n_fraction=nfrac, n_run=9)
    lib = synth.random_library(n, L, seed=seed + 1)
    synth.plant(lib, genome, 0.5, k, seed=seed + 2)
    lib[1, L // 3] = ord("N")
    lib[2, :] = ord("N")
    lib[5, 0] = ord("a") + (lib[5, 0] - ord("A")) if lib[5, 0] < 97 else lib[5, 0]  # lower-case base
    return genome, off, lib


@pytest.mark.parametrize("L", [1, 2, 5, 12, 19, 20, 21, 31, 32])
@pytest.mark.parametrize("k", [0, 1, 2, 3])
def test_probe_matches_oracle_all_lengths(L, k):
    genome, off, lib = small_case(L, k, seed=100 * L + k, n=120, G=20000)
    ref = run_oracle(genome, off, lib, k, pam="NGG", mode="brute" if L < 4 else "seeded")
    gpu, st = run_gpu(genome, off, lib, k, pam="NGG", path=1)
    assert_same(gpu, ref)
    if L > k:
        assert st["path"] == 1 and st["scan_launches"] >= 1


@pytest.mark.parametrize("k,blocks", [(0, 1), (0, 2), (0, 4), (1, 2), (1, 3), (1, 5), (2, 3), (2, 4), (2, 6),
                                      (3, 4), (3, 5), (3, 6), (3, 7)])
@pytest.mark.parametrize("L", [20, 32])
def test_probe_every_seed_scheme(k, blocks, L):
    genome, off, lib = small_case(L, k, seed=7 * blocks + k + L)
    ref = run_oracle(genome, off, lib, k, pam="NGG")
    gpu, st = run_gpu(genome, off, lib, k, pam="NGG", blocks=blocks, path=1)
    assert st["blocks"] == blocks
    assert_same(gpu, ref)


@pytest.mark.parametrize("direction", ["downstream", "upstream"])
@pytest.mark.parametrize("pam,iupac", [("NGG", False), ("NGNC", False), ("NNGRRT", True), ("NNGRRT", False),
                                       ("TTTV", True), ("", False), ("NNNNNNNN", False)])
def test_pam_modes(direction, pam, iupac):
    genome, off, lib = small_case(20, 2, seed=len(pam) + 31, n=400, G=40000, nfrac=0.03)
    for gate in (False, True):
        ref = run_oracle(genome, off, lib, 2, pam=pam, direction=direction, iupac=iupac, gate=gate)
        gpu, _ = run_gpu(genome, off, lib, 2, pam=pam, direction=direction, iupac=iupac, gate=gate, path=1)
        assert_same(gpu, ref)


def test_golden_g2_through_abi(plasmids, cn32_spacers, golden_dir):
    """SURVEY.md 8c G2 through the CUDA path: 869 hits, identical records to the oracle, and the
    committed canonical tuple list."""
    ids = list(plasmids)
    contigs = [str(plasmids[i].seq).encode() for i in ids]
    genome = np.frombuffer(b"".join(contigs), dtype=np.uint8)
    off = np.zeros(len(contigs) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([len(c) for c in contigs])
    lib = np.frombuffer("".join(cn32_spacers).encode(), dtype=np.uint8).reshape(-1, 32)
    ref = run_oracle(genome, off, lib, 2, pam="NGNC")
    gpu, _ = run_gpu(genome, off, lib, 2, pam="NGNC")
    assert len(gpu) == 869
    assert_same(gpu, ref)
    tup = []
    for h in gpu:
        ci = int((off[1:] <= h["gpos"]).sum())
        meta = int(h["meta"])
        pam = "".join("ACGT"[(meta >> (16 + 2 * i)) & 3] for i in range(4)) if meta & 16 else ""
        tup.append((cn32_spacers[h["spacer_id"]], ids[ci], int(h["gpos"] - off[ci]), "-" if meta & 1 else "+",
                    (meta >> 1) & 3, pam))
    text = "\n".join("\t".join(str(x) for x in t) for t in sorted(tup))
    with open(os.path.join(golden_dir, "g2_known_answer.tsv")) as h:
        assert h.read().rstrip("\n") == text


def test_edge_cases():
    # empty library, genome shorter than L, contig shorter than L, all-N contig, empty contig
    genome = np.frombuffer(b"ACGTACGTAC" + b"NNNNNNNNNNNNNNNNNNNNNNNNN" + b"ACG" + b"ACGTTGCAACGTTGCAACGTAAGG", dtype=np.uint8)
    off = np.array([0, 10, 35, 35, 38, 62], dtype=np.uint64)
    lib = np.frombuffer(b"ACGTTGCAACGTTGCAACGT" + b"ACGTACGTACACGTACGTAC", dtype=np.uint8).reshape(2, 20)
    ref = run_oracle(genome, off, lib, 1, pam="NGG", mode="brute")
    gpu, _ = run_gpu(genome, off, lib, 1, pam="NGG")
    assert len(ref) >= 1
    assert_same(gpu, ref)
    with _native.Searcher(0) as s:
        s.set_genome_array(genome, off)
        s.set_library(np.zeros((0, 20), dtype=np.uint8))
        assert s.search(2) == 0
        s.set_library(lib)
        assert s.search(0) == len(run_oracle(genome, off, lib, 0))
        with pytest.raises(_native.NativeError):
            s.search(4)
        with pytest.raises(_native.NativeError):
            s.set_library(np.zeros((1, 33), dtype=np.uint8))
    with _native.Searcher(0) as s:
        with pytest.raises(_native.NativeError):
            s.search(1)  # nothing loaded


def test_hit_buffer_overflow_retries():
    genome, off = synth.random_genome(200000, seed=3)
    lib = synth.random_library(2000, 12, seed=4)
    ref = run_oracle(genome, off, lib, 2)
    assert len(ref) > 5000
    gpu, st = run_gpu(genome, off, lib, 2, hit_cap=1000)
    assert st["scan_launches"] == 2
    assert_same(gpu, ref)


def test_palindrome_reported_on_both_strands():
    genome = np.frombuffer(b"TTTTACGTACGTACGTACGTAAAAGG", dtype=np.uint8)
    off = np.array([0, len(genome)], dtype=np.uint64)
    lib = np.frombuffer(b"ACGTACGTACGTACGT", dtype=np.uint8).reshape(1, 16)
    gpu, _ = run_gpu(genome, off, lib, 0)
    assert_same(gpu, run_oracle(genome, off, lib, 0, mode="brute"))
    assert sorted((int(h["gpos"]), int(h["meta"] & 1)) for h in gpu) == [(4, 0), (4, 1)]


def test_medium_random_many_contigs():
    genome, off = synth.random_genome(3_000_000, seed=11, n_contigs=40, n_fraction=0.002)
    lib = synth.random_library(20000, 20, seed=12)
    synth.plant(lib, genome, 0.3, 3, seed=13)
    ref = run_oracle(genome, off, lib, 3, pam="NGG")
    gpu, st = run_gpu(genome, off, lib, 3, pam="NGG")
    assert_same(gpu, ref)
    assert len(gpu) > 6000


# ------------------------------------------------------------------ bucket-join path (K3-join)
@pytest.mark.parametrize("k,blocks", [(0, 1), (0, 3), (1, 2), (1, 4), (2, 3), (2, 5), (3, 4), (3, 5), (3, 6), (3, 7)])
@pytest.mark.parametrize("L", [20, 32])
def test_join_every_seed_scheme(k, blocks, L):
    genome, off, lib = small_case(L, k, seed=11 * blocks + k + L)
    ref = run_oracle(genome, off, lib, k, pam="NGG")
    gpu, st = run_gpu(genome, off, lib, k, pam="NGG", blocks=blocks, path=2)
    assert st["blocks"] == blocks and st["path"] == 2
    assert_same(gpu, ref)


@pytest.mark.parametrize("L", [1, 3, 8, 17, 20, 32])
@pytest.mark.parametrize("k", [0, 1, 3])
def test_join_matches_oracle_lengths(L, k):
    genome, off, lib = small_case(L, k, seed=200 * L + k, n=120, G=20000)
    ref = run_oracle(genome, off, lib, k, pam="NGNC", mode="brute" if L < 4 else "seeded")
    gpu, _ = run_gpu(genome, off, lib, k, pam="NGNC", path=2)
    assert_same(gpu, ref)


def test_join_dense_buckets_exceed_shared_tile():
    """Library buckets larger than the shared-memory tile (1024 entries) and duplicate spacers."""
    genome, off = synth.random_genome(3000, seed=21, n_contigs=2, n_fraction=0.01, n_run=5)
    lib = synth.random_library(50000, 6, seed=22, distinct=False)
    ref = run_oracle(genome, off, lib, 1, pam="NGG")
    gpu, st = run_gpu(genome, off, lib, 1, pam="NGG", blocks=2, path=2)
    assert len(ref) > 500000
    assert_same(gpu, ref)
    gpu2, _ = run_gpu(genome, off, lib, 1, pam="NGG", blocks=2, path=1)
    assert_same(gpu2, ref)


def test_join_medium_both_paths_agree():
    genome, off = synth.random_genome(2_000_000, seed=31, n_contigs=7, n_fraction=0.002)
    lib = synth.random_library(100000, 20, seed=32)
    synth.plant(lib, genome, 0.2, 3, seed=33)
    ref = run_oracle(genome, off, lib, 3, pam="NGG")
    for path in (1, 2):
        gpu, st = run_gpu(genome, off, lib, 3, pam="NGG", path=path)
        assert st["path"] == path
        assert_same(gpu, ref)
    gate_ref = run_oracle(genome, off, lib, 3, pam="NGG", gate=True)
    gpu, _ = run_gpu(genome, off, lib, 3, pam="NGG", gate=True, path=2)
    assert_same(gpu, gate_ref)


@pytest.mark.parametrize("path", [1, 2])
def test_hit_sink_streams_the_same_records(path):
    """bc_set_hit_sink: records delivered to host memory during the search equal bc_copy_hits'."""
    genome, off = synth.random_genome(1_500_000, seed=41, n_contigs=5, n_fraction=0.002)
    lib = synth.random_library(60000, 20, seed=42)
    synth.plant(lib, genome, 0.3, 2, seed=43)
    ref = run_oracle(genome, off, lib, 2, pam="NGG")
    with _native.Searcher(0) as s:
        s.set_genome_array(genome, off)
        s.set_library(lib)
        s.set_pam("NGG", "downstream")
        s.set_param(_native.BC_PARAM_PATH, path)
        sink = np.zeros(len(ref) + 100, dtype=_native.HIT_DTYPE)
        s.set_hit_sink(sink.ctypes.data, len(sink))
        for _ in range(2):              # the sink persists across searches
            sink[:] = 0
            n = s.search(2)
            assert n == len(ref)
            assert_same(_native.canonical_sort(sink[:n].copy()), ref)
        assert_same(_native.canonical_sort(s.hits()), ref)   # the device copy is still there
        # a sink that is too small is an error, but the records stay retrievable
        small = np.zeros(len(ref) - 1, dtype=_native.HIT_DTYPE)
        s.set_hit_sink(small.ctypes.data, len(small))
        with pytest.raises(_native.NativeError):
            s.search(2)
        assert_same(_native.canonical_sort(s.hits()), ref)
        s.set_hit_sink(None, 0)
        assert s.search(2) == len(ref)


def test_hit_sink_with_device_buffer_overflow_retry():
    """The device buffer is grown and the search repeated; the sink must end up with the full set."""
    genome, off, lib = small_case(20, 2, seed=77, n=2000, G=200000)
    ref = run_oracle(genome, off, lib, 2, pam="NGG")
    assert len(ref) > 64
    with _native.Searcher(0) as s:
        s.set_genome_array(genome, off)
        s.set_library(lib)
        s.set_pam("NGG", "downstream")
        s.set_param(_native.BC_PARAM_PATH, 2)
        s.set_param(_native.BC_PARAM_HIT_CAPACITY, 16)
        sink = np.zeros(len(ref), dtype=_native.HIT_DTYPE)
        s.set_hit_sink(sink.ctypes.data, len(sink))
        assert s.search(2) == len(ref)
        assert_same(_native.canonical_sort(sink.copy()), ref)


# ------------------------------------------------ radix form of the genome-side sort (pass A/B)
@pytest.mark.parametrize("k,blocks", [(0, 1), (1, 2), (1, 4), (2, 3), (2, 5), (3, 4), (3, 5), (3, 6)])
@pytest.mark.parametrize("L", [20, 32])
def test_join_radix_window_sort_every_seed_scheme(k, blocks, L):
    """BC_PARAM_WINDOW_SORT=2 must give the same records as the direct scatter wherever it applies
    (keys of 4..8 nt) and fall back silently elsewhere."""
    genome, off, lib = small_case(L, k, seed=13 * blocks + k + L, n=400, G=90000)
    ref = run_oracle(genome, off, lib, k, pam="NGG")
    gpu, st = run_gpu(genome, off, lib, k, pam="NGG", blocks=blocks, path=2, window_sort=2)
    assert st["path"] == 2
    assert_same(gpu, ref)


@pytest.mark.parametrize("gate", [False, True])
def test_join_radix_window_sort_medium(gate):
    """Several 2048-record chunks per bin, ragged last chunk, N runs, many contigs, PAM gate."""
    genome, off = synth.random_genome(2_500_000, seed=51, n_contigs=9, n_fraction=0.003)
    lib = synth.random_library(150000, 20, seed=52)
    synth.plant(lib, genome, 0.2, 3, seed=53)
    ref = run_oracle(genome, off, lib, 3, pam="NGG", gate=gate)
    for ws in (1, 2):
        gpu, st = run_gpu(genome, off, lib, 3, pam="NGG", gate=gate, blocks=5, path=2, window_sort=ws)
        assert_same(gpu, ref)


def test_join_radix_window_sort_short_keys_and_dense_slots():
    """4-nt keys (one bin per combination), huge slots, duplicate spacers."""
    genome, off = synth.random_genome(300000, seed=61, n_contigs=3, n_fraction=0.01, n_run=5)
    lib = synth.random_library(3000, 8, seed=62, distinct=False)
    ref = run_oracle(genome, off, lib, 1, pam="NGG")
    gpu, st = run_gpu(genome, off, lib, 1, pam="NGG", blocks=2, path=2, window_sort=2)
    assert len(ref) > 100000
    assert_same(gpu, ref)


# ------------------------------------------------ compact bucket join (8-byte window records, path 3)
@pytest.mark.parametrize("k,blocks", [(0, 1), (0, 3), (1, 2), (1, 4), (2, 3), (2, 5), (3, 4), (3, 5), (3, 6), (3, 7)])
@pytest.mark.parametrize("L", [12, 17, 20])
def test_cjoin_every_block_scheme(k, blocks, L):
    genome, off, lib = small_case(L, k, seed=17 * blocks + k + L, n=400, G=90000)
    ref = run_oracle(genome, off, lib, k, pam="NGG")
    gpu, st = run_gpu(genome, off, lib, k, pam="NGG", blocks=blocks, path=3)
    assert st["blocks"] == blocks and st["path"] == 3
    assert_same(gpu, ref)


@pytest.mark.parametrize("path", [1, 2, 3])
@pytest.mark.parametrize("key_nt", [9, 10])
def test_covering_designs_every_path(path, key_nt):
    """Seed covering designs (bc_designs.inc) instead of block schemes: L=20, k=3, keys of 9 / 10 nt.
    Same records from the probe kernel, the 16-byte join and the compact join."""
    genome, off, lib = small_case(20, 3, seed=900 + key_nt, n=3000, G=400000, n_contigs=5, nfrac=0.004)
    ref = run_oracle(genome, off, lib, 3, pam="NGG")
    gpu, st = run_gpu(genome, off, lib, 3, pam="NGG", key_nt=key_nt, path=path)
    assert st["blocks"] == 0 and st["key_nt"] == key_nt and st["path"] == path
    assert st["combos"] == {9: 12, 10: 15}[key_nt]
    assert_same(gpu, ref)


@pytest.mark.parametrize("gate", [False, True])
@pytest.mark.parametrize("key_nt,blocks", [(0, 5), (9, 0), (10, 0), (0, 6)])
def test_cjoin_medium_multi_chunk_bins(gate, key_nt, blocks):
    """Several 8192-record chunks per bin, chunks straddling bins, ragged tiles, N runs, PAM gate,
    two forced passes over the genome."""
    genome, off = synth.random_genome(3_000_000, seed=151, n_contigs=9, n_fraction=0.003)
    lib = synth.random_library(200000, 20, seed=152)
    synth.plant(lib, genome, 0.2, 3, seed=153)
    ref = run_oracle(genome, off, lib, 3, pam="NGG", gate=gate)
    gpu, st = run_gpu(genome, off, lib, 3, pam="NGG", gate=gate, blocks=blocks, key_nt=key_nt, path=3)
    assert st["path"] == 3
    assert_same(gpu, ref)
    gpu2, _ = run_gpu(genome, off, lib, 3, pam="NGG", gate=gate, blocks=blocks, key_nt=key_nt, path=3,
                      join_chunk=1_700_001)
    assert_same(gpu2, ref)


@pytest.mark.parametrize("path,key_nt", [(3, 10), (3, 9), (1, 0)])
def test_low_complexity_genome_skewed_bins(path, key_nt):
    """Homopolymer runs and tandem repeats: a few seed bins / slots hold most of the windows (skewed pass-A bins, slots
    of thousands of windows against buckets of near-identical spacers, item-queue bursts: whole warps pass at once)."""
    rng = np.random.default_rng(191)
    parts = []
    for i in range(40):
        kind = i % 4
        if kind == 0:
            parts.append(rng.integers(0, 4, 4000))                        # random
        elif kind == 1:
            parts.append(np.full(3000, i % 3))                            # homopolymer
        elif kind == 2:
            parts.append(np.tile(rng.integers(0, 4, 2 + i % 5), 1500)[:3000])   # short tandem repeat
        else:
            parts.append(np.tile(rng.integers(0, 4, 23), 200)[:3000])     # 23-mer repeat
    codes = np.concatenate(parts)
    genome = np.frombuffer(b"ACGT", dtype=np.uint8)[codes].copy()
    off = np.array([0, len(genome) // 2, len(genome)], dtype=np.uint64)
    lib = synth.random_library(1500, 20, seed=192)
    # spacers taken from the repeats themselves (with up to 2 edits), plus pure homopolymers
    for j in range(0, 600):
        p0 = int(rng.integers(0, len(genome) - 20))
        lib[j] = genome[p0:p0 + 20]
        for _ in range(j % 3):
            lib[j, int(rng.integers(0, 20))] = b"ACGT"[int(rng.integers(0, 4))]
    lib[600] = ord("A"); lib[601] = ord("C"); lib[602] = ord("T")
    ref = run_oracle(genome, off, lib, 3, pam="NGG")
    assert len(ref) > 200_000
    gpu, st = run_gpu(genome, off, lib, 3, pam="NGG", path=path, key_nt=key_nt)
    assert st["path"] == path
    assert_same(gpu, ref)


def test_cjoin_dense_slots_duplicates_and_big_buckets():
    """Huge slots and buckets (4-nt keys), duplicate spacers, short spacers."""
    genome, off = synth.random_genome(300000, seed=161, n_contigs=3, n_fraction=0.01, n_run=5)
    lib = synth.random_library(3000, 8, seed=162, distinct=False)
    ref = run_oracle(genome, off, lib, 1, pam="NGG")
    gpu, st = run_gpu(genome, off, lib, 1, pam="NGG", blocks=2, path=3)
    assert len(ref) > 100000 and st["path"] == 3
    assert_same(gpu, ref)
    lib = synth.random_library(300_000, 12, seed=172)
    synth.plant(lib, genome, 0.05, 1, seed=173)
    ref = run_oracle(genome, off, lib, 1, pam="NGG")
    gpu, st = run_gpu(genome, off, lib, 1, pam="NGG", blocks=2, path=3)
    assert_same(gpu, ref)


def test_cjoin_hit_sink_and_overflow():
    genome, off, lib = small_case(20, 2, seed=177, n=2000, G=200000)
    ref = run_oracle(genome, off, lib, 2, pam="NGG")
    with _native.Searcher(0) as s:
        s.set_genome_array(genome, off)
        s.set_library(lib)
        s.set_pam("NGG", "downstream")
        s.set_param(_native.BC_PARAM_PATH, 3)
        s.set_param(_native.BC_PARAM_HIT_CAPACITY, 16)
        sink = np.zeros(len(ref), dtype=_native.HIT_DTYPE)
        s.set_hit_sink(sink.ctypes.data, len(sink))
        assert s.search(2) == len(ref)
        assert_same(_native.canonical_sort(sink.copy()), ref)
        assert s.stats()["path"] == 3


def test_cjoin_item_queue_overflow_retries():
    """The verify kernel has no slow path: when the global item queue between k_cverify and k_cfinish is too small
    batches are dropped, the demand is counted and bc_search repeats the search with a queue of that size."""
    genome, off = synth.random_genome(300000, seed=181, n_contigs=3, n_fraction=0.01, n_run=5)
    lib = synth.random_library(300_000, 12, seed=182)
    synth.plant(lib, genome, 0.05, 1, seed=183)
    ref = run_oracle(genome, off, lib, 1, pam="NGG")
    # 1024-record hit buffer -> item queue of 2 * 1024 + 65536 entries; this job queues several 10^5 items
    gpu, st = run_gpu(genome, off, lib, 1, pam="NGG", blocks=2, path=3, hit_cap=1024)
    assert st["path"] == 3 and st["search_attempts"] >= 2
    assert_same(gpu, ref)
    gpu, st = run_gpu(genome, off, lib, 1, pam="NGG", blocks=2, path=3)
    assert st["search_attempts"] == 1
    assert_same(gpu, ref)


@pytest.mark.parametrize("index_sort", [1, 2])
def test_cjoin_index_builders_agree(index_sort):
    """The compact path's library index can be built by the two radix passes (default) or by the
    two-level atomic scatter of the other paths (BC_PARAM_INDEX_SORT): same records; spacers with
    non-ACGT characters are skipped per combination in both."""
    genome, off, lib = small_case(20, 3, seed=333, n=5000, G=500000, n_contigs=4, nfrac=0.004)
    lib[17, 3] = ord("N")
    ref = run_oracle(genome, off, lib, 3, pam="NGG")
    with _native.Searcher(0) as s:
        s.set_genome_array(genome, off)
        s.set_library(lib)
        s.set_pam("NGG", "downstream")
        s.set_param(_native.BC_PARAM_PATH, 3)
        s.set_param(_native.BC_PARAM_KEY_NT, 9)
        s.set_param(_native.BC_PARAM_INDEX_SORT, index_sort)
        n = s.search(3)
        assert n == len(ref)
        assert_same(_native.canonical_sort(s.hits()), ref)


@pytest.mark.parametrize("order", ["canonical", "best"])
def test_device_hit_sort(order):
    """bc_sort_hits: the device buffer comes back ordered like the host lexsort the parity tests use
    ("canonical") or like `bowtie --best` (fewest mismatches first within a read)."""
    genome, off, lib = small_case(20, 3, seed=444, n=4000, G=600000, n_contigs=5, nfrac=0.004)
    with _native.Searcher(0) as s:
        s.set_genome_array(genome, off)
        s.set_library(lib)
        s.set_pam("NGG", "downstream")
        n = s.search(3)
        raw = s.hits()
        s.sort_hits(order)
        got = s.hits()
        assert s.stats()["ms_sort_hits"] > 0
    assert n == len(got) > 1000
    if order == "canonical":
        want = raw[np.lexsort((raw["meta"] & 1, raw["gpos"], raw["spacer_id"]))]
    else:
        want = raw[np.lexsort((raw["meta"] & 1, raw["gpos"], (raw["meta"] >> 1) & 3, raw["spacer_id"]))]
    assert got.tobytes() == want.tobytes()


def test_device_hit_sort_large_and_degenerate():
    """> 1 tile per digit pass, many equal keys (stability), single-record and empty buffers."""
    genome, off = synth.random_genome(3000, seed=21, n_contigs=2, n_fraction=0.01, n_run=5)
    lib = synth.random_library(50000, 6, seed=22, distinct=False)
    with _native.Searcher(0) as s:
        s.set_genome_array(genome, off)
        s.set_library(lib)
        s.set_pam("NGG", "downstream")
        n = s.search(1)
        raw = s.hits()
        s.sort_hits("best")
        got = s.hits()
        assert n == len(got) > 100000
        want = raw[np.lexsort((raw["meta"] & 1, raw["gpos"], (raw["meta"] >> 1) & 3, raw["spacer_id"]))]
        assert got.tobytes() == want.tobytes()
        s.set_library(np.frombuffer(b"ACGTAC", np.uint8).reshape(1, 6))
        s.search(0)
        s.sort_hits("canonical")
        assert len(s.hits()) == s.stats()["hits"]


@pytest.mark.parametrize("packed", [0, 1])
@pytest.mark.parametrize("L,k,blocks", [(32, 2, 3), (20, 3, 4), (12, 0, 1)])
def test_probe_packed_directory(L, k, blocks, packed):
    """Probe path with the packed directory (one 4-byte load per probe: start | count << 26; the
    default) and without it (BC_PARAM_COMPACT_DIR = 1): same records, including buckets of 63 or
    more equal entries, which fall back to the 32-bit directory."""
    genome, off, lib = small_case(L, k, seed=555 + L + k, n=3000, G=300000, n_contigs=5, nfrac=0.004)
    lib[100:240] = lib[99]          # a bucket with 141 equal entries (> 63)
    ref = run_oracle(genome, off, lib, k, pam="NGG")
    with _native.Searcher(0) as s:
        s.set_genome_array(genome, off)
        s.set_library(lib)
        s.set_pam("NGG", "downstream")
        s.set_param(_native.BC_PARAM_PATH, 1)
        s.set_param(_native.BC_PARAM_BLOCKS, blocks)
        s.set_param(_native.BC_PARAM_COMPACT_DIR, packed)
        n = s.search(k)
        assert n == len(ref) and s.stats()["path"] == 1
        assert_same(_native.canonical_sort(s.hits()), ref)
