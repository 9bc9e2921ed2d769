"""CPU tests of the oracle itself: pinned to the golden vectors, brute vs seeded vs pure Python."""
import hashlib
import os

import numpy as np
import pytest

from barcoder_b200 import synth
from oracle import oracle

SURVEY_SHA = "5dbbc9c4be09b08acddf816ffebfa751eff16374e50da6054f2854c03aa493da"


def _tuples(hits, contigs, spacers, ids, P):
    _, off = oracle.concat_genome(contigs)
    out = []
    for h in hits:
        ci = int((off[1:] <= h["gpos"]).sum())
        meta = int(h["meta"])
        pam = "".join("ACGT"[(meta >> (16 + 2 * i)) & 3] for i in range(P)) if meta & oracle.META_PAM_FULL else ""
        out.append((spacers[h["spacer_id"]], ids[ci], int(h["gpos"] - off[ci]), "-" if meta & 1 else "+",
                    (meta >> 1) & 3, pam))
    return sorted(out)


def test_g2_known_answer_sha(plasmids, cn32_spacers, golden_dir):
    """869 hits = 757/68/44, sha256 pinned (seeded strategy; the committed file
    was produced by the brute strategy in tests/golden/make_golden.py)."""
    ids = list(plasmids)
    contigs = [str(plasmids[i].seq) for i in ids]
    hits = oracle.search(contigs, cn32_spacers, 2, pam="NGNC", mode="seeded")
    tup = _tuples(hits, contigs, cn32_spacers, ids, 4)
    assert len(tup) == 869
    assert [sum(1 for t in tup if t[4] == m) for m in range(3)] == [757, 68, 44]
    text = "\n".join("\t".join(str(x) for x in t) for t in tup)
    assert hashlib.sha256(text.encode()).hexdigest() == SURVEY_SHA
    with open(os.path.join(golden_dir, "g2_known_answer.tsv")) as h:
        assert h.read().rstrip("\n") == text
    ngnc = [t for t in tup if len(t[5]) == 4 and t[5][1] == "G" and t[5][3] == "C"]
    assert len(ngnc) == 839
    assert sum(1 for h in hits if h["meta"] & oracle.META_PAM_OK) == 839


def test_g1_fixture_rows_are_hits(plasmids, cn32_spacers, golden_dir):
    """Every plasmid row of the reference's CN-32-zmo.tsv is an exact hit with the stated PAM,
    and the fixture's `target` equals the reference slice in spacer orientation (G1/G3)."""
    import csv
    ids = list(plasmids)
    contigs = [str(plasmids[i].seq) for i in ids]
    hits = oracle.search(contigs, cn32_spacers, 2, pam="NGNC", mode="seeded")
    have = set(_tuples(hits, contigs, cn32_spacers, ids, 4))
    with open(os.path.join(golden_dir, "cn32_plasmid_rows.tsv")) as h:
        rows = list(csv.DictReader(h, delimiter="\t"))
    assert len(rows) == 772
    sites = set()
    for r in rows:
        strand = "+" if r["sp_dir"] == "F" else "-"
        t = (r["spacer"], r["chr"], int(r["tar_start"]), strand, int(r["mismatches"]), r["pam"])
        assert t in have
        sites.add(t[:4])
        ref = str(plasmids[r["chr"]].seq)[int(r["tar_start"]):int(r["tar_end"])]
        assert r["target"] == (ref if strand == "+" else oracle.revcomp(ref))
    assert len(sites) == 750
    exact_ngnc = {t[:4] for t in have if t[4] == 0 and len(t[5]) == 4 and t[5][1] == "G" and t[5][3] == "C"}
    assert exact_ngnc == sites


@pytest.mark.parametrize("L,k", [(20, 0), (20, 1), (20, 3), (32, 2), (7, 3), (3, 3), (1, 0)])
def test_brute_vs_seeded_vs_python(L, k):
    genome, off = synth.random_genome(3000, seed=L * 10 + k, n_contigs=3, n_fraction=0.02, n_run=7)
    lib = synth.random_library(40, L, seed=5)
    synth.plant(lib, genome, 0.6, k, seed=6)
    lib[3, L // 2] = ord("N")
    contigs = [bytes(genome[int(off[i]):int(off[i + 1])]).decode() for i in range(3)]
    spacers = synth.rows_to_strings(lib)
    a = oracle.search(contigs, spacers, k, mode="brute", threads=3)
    b = oracle.search(contigs, spacers, k, mode="seeded", threads=2)
    assert np.array_equal(a, b)
    py = oracle.py_search(contigs, spacers, k)
    _, o = oracle.concat_genome(contigs)
    got = sorted((int(h["spacer_id"]), int((o[1:] <= h["gpos"]).sum()),
                  int(h["gpos"] - o[int((o[1:] <= h["gpos"]).sum())]), "-" if h["meta"] & 1 else "+",
                  int(h["meta"] >> 1) & 3,
                  tuple(j for j in range(L) if h["mm_mask"] >> j & 1)) for h in a)
    assert got == py
    if L > k:
        assert len(a) > 0


def test_pam_annotation_matches_string_rules():
    genome, off = synth.random_genome(5000, seed=77, n_contigs=2, n_fraction=0.03, n_run=5)
    lib = synth.random_library(60, 20, seed=8)
    synth.plant(lib, genome, 0.9, 1, seed=9)
    contigs = [bytes(genome[int(off[i]):int(off[i + 1])]).decode() for i in range(2)]
    spacers = synth.rows_to_strings(lib)
    _, o = oracle.concat_genome(contigs)
    for direction in ("downstream", "upstream"):
        hits = oracle.search(contigs, spacers, 1, pam="NGG", direction=direction)
        assert len(hits)
        for h in hits:
            ci = int((o[1:] <= h["gpos"]).sum())
            start = int(h["gpos"] - o[ci])
            strand = "-" if h["meta"] & 1 else "+"
            s, ok = oracle.py_pam_script(contigs[ci], start, start + 20, strand, "NGG", direction)
            meta = int(h["meta"])
            full = bool(meta & oracle.META_PAM_FULL)
            assert full == (s is not None)
            if full and not meta & oracle.META_PAM_AMB:
                assert s == "".join("ACGT"[(meta >> (16 + 2 * i)) & 3] for i in range(3))
                assert bool(meta & oracle.META_PAM_OK) == ok
            if direction == "downstream":
                s2, ok2 = oracle.py_pam_class_api(contigs[ci], start, start + 20, strand, "NGG")
                if full and not meta & oracle.META_PAM_AMB:
                    assert s2 == s and ok2 == ok
                elif not full:
                    assert not ok2
        gated = oracle.search(contigs, spacers, 1, pam="NGG", direction=direction, flags=oracle.PAM_FLAG_GATE)
        keep = hits[((hits["meta"] & oracle.META_PAM_OK) != 0) | ((hits["meta"] & oracle.META_PAM_AMB) != 0)]
        assert np.array_equal(gated, keep)


@pytest.mark.parametrize("L,k", [(20, 3), (20, 2), (32, 2), (12, 1), (9, 3)])
def test_seeded_every_block_count_equals_brute(L, k):
    """The seeded strategy must give the exhaustive answer under every seed scheme it can choose
    (b = k+1 .. k+4 blocks), including spacers with non-ACGT characters and N runs in the genome."""
    genome, off = synth.random_genome(6000, seed=L + 3 * k, n_contigs=4, n_fraction=0.02, n_run=6)
    lib = synth.random_library(80, L, seed=15)
    synth.plant(lib, genome, 0.7, k, seed=16)
    lib[2, 0] = ord("N")
    lib[7, L - 1] = ord("n")
    lib[9, :] = ord("N")
    contigs = [bytes(genome[int(off[i]):int(off[i + 1])]).decode() for i in range(4)]
    spacers = synth.rows_to_strings(lib)
    want = oracle.search(contigs, spacers, k, mode="brute", threads=4)
    assert len(want) > 0
    for b in range(k + 1, min(k + 4, 8, L) + 1):
        got = oracle.search(contigs, spacers, k, mode="seeded", threads=3, blocks=b)
        assert np.array_equal(want, got), b
