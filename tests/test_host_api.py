"""CPU tests of the host-side mirror of the reference interface (no GPU needed)."""
import os
import types

import numpy as np
import pandas as pd
import pytest

from barcoder_b200 import (BarCodeLibrary, BowtieError, BowtieRunner, CRISPRiLibrary, GenBankParser, GuideFinder,
                           PAMFinder, PySamParser, samio, seqio)
from barcoder_b200.ranges import PyRanges, overlap_pairs
from oracle import oracle


def test_genbank_reader_golden_plasmids(golden_dir, plasmids):
    assert list(plasmids) == ["CP023716.1", "CP023717.1", "CP023718.1", "CP023719.1"]
    assert [len(r.seq) for r in plasmids.values()] == [32791, 33006, 36494, 39266]
    gb = GenBankParser(os.path.join(golden_dir, "zmo_plasmids.gb"))
    assert set(gb.topologies.values()) == {"circular"}
    assert gb.overhangs["CP023716.1"] == 100_000
    assert list(gb.num_genes.values()) == [31, 40, 52, 37]
    assert gb.seq_lens["CP023719.1"] == 39266
    assert "Zymomonas" in gb.organisms["CP023716.1"]
    df = gb.ranges.df
    assert set(df.columns) == {"Chromosome", "Start", "End", "Strand", "Locus_Tag", "Gene", "Type"}
    assert (df["Type"] == "source").sum() == 4
    # origin-spanning gene 36440..284 becomes two parts
    parts = df[df["Locus_Tag"] == "ZMO1_ZMOp36x053"]
    assert sorted(zip(parts["Start"], parts["End"])) == [(0, 284), (36439, 36494)]
    assert gb.find_gene_name_for_locus("ZMO1_ZMOp36x053") == "ZMO1_ZMOp36x053"
    assert gb.find_gene_name_for_locus("nope") is None


def test_genbank_roundtrip_and_locations(tmp_path):
    rec = seqio.SeqRecord(seqio.Seq("ACGTNNACGT" * 30), id="X1.2", name="X1", description="test contig")
    rec.annotations.update(topology="linear", organism="Testus maximus")
    rec.features = [
        seqio.SeqFeature(seqio.SimpleLocation(0, 300, 1), "source", {"organism": ["Testus maximus"]}),
        seqio.SeqFeature(seqio.SimpleLocation(9, 50, -1), "gene", {"locus_tag": ["T_1"], "gene": ["abc"]}),
        seqio.SeqFeature(seqio.CompoundLocation([seqio.SimpleLocation(60, 70, 1), seqio.SimpleLocation(80, 99, 1)]),
                         "gene", {"locus_tag": ["T_2"]}),
    ]
    path = tmp_path / "x.gb"
    seqio.write_genbank([rec], str(path))
    back = seqio.genbank_to_dict(str(path))["X1.2"]
    assert str(back.seq) == str(rec.seq) and back.annotations["topology"] == "linear"
    assert back.annotations["organism"] == "Testus maximus"
    g1, g2 = [f for f in back.features if f.type == "gene"]
    assert (g1.location.start, g1.location.end, g1.location.strand) == (9, 50, -1)
    assert [(p.start, p.end) for p in g2.location.parts] == [(60, 70), (80, 99)]
    loc = seqio._parse_location("complement(join(5..10,20..30))")
    assert [(p.start, p.end, p.strand) for p in loc.parts] == [(19, 30, -1), (4, 10, -1)]
    assert seqio.Seq("AACG").reverse_complement() == "CGTT"


def test_overlap_join_matches_bruteforce():
    rng = np.random.default_rng(0)
    a = pd.DataFrame({"Chromosome": rng.choice(["c1", "c2"], 200), "Start": rng.integers(0, 1000, 200)})
    a["End"] = a["Start"] + rng.integers(1, 40, 200)
    a["Strand"] = rng.choice(["+", "-"], 200)
    b = pd.DataFrame({"Chromosome": rng.choice(["c1", "c2", "c3"], 60), "Start": rng.integers(0, 1000, 60)})
    b["End"] = b["Start"] + rng.integers(1, 300, 60)
    b["Strand"] = rng.choice(["+", "-"], 60)
    b["Type"] = "gene"
    li, ri = overlap_pairs(a, b)
    want = sorted((i, j) for i in range(200) for j in range(60)
                  if a.Chromosome[i] == b.Chromosome[j] and a.Start[i] < b.End[j] and b.Start[j] < a.End[i])
    assert sorted(zip(li.tolist(), ri.tolist())) == want
    joined = PyRanges(a).join(PyRanges(b)).df
    assert len(joined) == len(want)
    assert {"Start_b", "End_b", "Strand_b", "Type"} <= set(joined.columns)
    same = PyRanges(a).join(PyRanges(b), strandedness="same").df
    assert (same["Strand"] == same["Strand_b"]).all() and len(same) < len(joined)


def test_sam_writer_reader_roundtrip(tmp_path):
    contigs = ["ACGTACGTAAGGTTTTCCCCGGGGAAAATTTT", "TTTTGGGGACGTACGTAAGG"]
    runner = types.SimpleNamespace(
        _reads=["ACGTACGTAA", "CCTTACGTAC", "GGGGGGGGGG"], _contigs=contigs, _contig_ids=["c1", "c2"],
        _offsets=np.array([0, 32, 52]),
        hits=np.array([(0, 0, 0, 0), (0, 40, 0b100, (1 << 1)), (1, 0, 0, 1), (1, 40, 0b10000000, 1 | (1 << 1))],
                      dtype=[("spacer_id", "<u4"), ("gpos", "<u4"), ("mm_mask", "<u4"), ("meta", "<u4")]))
    path = str(tmp_path / "x.sam")
    samio.write_sam(path, runner)
    reads = list(samio.read_sam(path))
    assert [r.flag for r in reads] == [0, 256, 16, 272, 4]
    assert reads[0].reference_name == "c1" and reads[0].reference_start == 0 and reads[0].reference_end == 10
    assert reads[1].get_tag("NM") == 1 and reads[1].get_tag("MD") == "2G7"
    assert reads[1].get_reference_sequence() == "ACgTACGTAA"
    assert reads[2].is_reverse and reads[2].query_sequence == "GTACGTAAGG"
    assert reads[3].get_tag("MD") == "2G7"  # spacer position 7 on '-' is window position 2
    assert reads[4].is_unmapped and reads[4].reference_name is None and not reads[4].has_tag("NM")
    df = PySamParser(path).ranges.df
    assert list(df["Barcode"]) == ["ACGTACGTAA", "ACGTACGTAA", "CCTTACGTAC", "CCTTACGTAC"]
    assert list(df["Strand"]) == ["+", "+", "-", "-"] and list(df["Start"]) == [0, 8, 0, 8]


def test_pam_finder_string_rules(plasmids):
    recs = {"c": seqio.SeqRecord(seqio.Seq("AAACCCGGGTTTAGGCATCGATCGTTAGCCCTA"), id="c")}
    pf = PAMFinder(recs, "NGG", "downstream")
    row = types.SimpleNamespace(Chromosome="c", Start=2, End=12, Strand="+")
    assert pf.get_pam_seq(row) == "AGG" and pf.pam_matches("AGG")
    row = types.SimpleNamespace(Chromosome="c", Start=31, End=33, Strand="+")
    assert pf.get_pam_seq(row) == "" and not pf.pam_matches("")
    row = types.SimpleNamespace(Chromosome="c", Start=9, End=19, Strand="-")
    assert pf.get_pam_seq(row) == "CCC" and not pf.pam_matches("CCC")
    row = types.SimpleNamespace(Chromosome="c", Start=2, End=12, Strand="-")  # negative slice start -> ""
    assert pf.get_pam_seq(row) == ""
    up = PAMFinder(recs, "NGG", "upstream")  # same slice as downstream (reference quirk)
    row = types.SimpleNamespace(Chromosome="c", Start=2, End=12, Strand="+")
    assert up.get_pam_seq(row) == "AGG"
    assert not PAMFinder(recs, "NGR", "downstream").pam_matches("AGG")  # R is a literal
    assert pf.get_strand("fwd") == 1 and pf.get_strand(-1) == -1
    with pytest.raises(ValueError):
        pf.get_strand("sideways")
    for s, (st, en, strand) in {"x": (2, 12, "+"), "y": (9, 19, "-")}.items():
        got = pf.get_pam_seq(types.SimpleNamespace(Chromosome="c", Start=st, End=en, Strand=strand))
        assert got == oracle.py_pam_class_api(str(recs["c"].seq), st, en, strand, "NGG")[0]


def test_guide_finder():
    recs = {"c": seqio.SeqRecord(seqio.Seq("TTGGACGTACGTAGGCC"), id="c")}
    down = GuideFinder(recs, "NGG", "downstream", 5).find_guides_from_pam()
    # forward matches: TGG@1 (guide 'T', truncated), AGG@12; reverse strand GGCCTACGTACGTCCAA: none with NGG? 'TGG'...
    fwd = "TTGGACGTACGTAGGCC"
    import re
    want = []
    for text in (fwd, oracle.revcomp(fwd)):
        for m in re.finditer("[ATCG]GG", text):
            want.append(text[max(0, m.start() - 5):m.start()])
    assert down == want and "T" in down
    up = GuideFinder(recs, "NGG", "upstream", 5).find_guides_from_pam()
    assert all(len(g) <= 5 for g in up)
    with pytest.raises(ValueError):
        GuideFinder(recs, "NGG", "sideways", 5).find_guides_from_pam()


def test_barcode_library(tmp_path):
    tsv = tmp_path / "lib.tsv"
    tsv.write_text("name\tspacer\nx\tACGT\ny\tACGT\nz\tTTTT\n")
    lib = BarCodeLibrary(str(tsv), column="spacer")
    assert lib.barcodes == {"ACGT", "TTTT"} and lib.size == 2
    fa = tmp_path / "lib.fasta"
    fa.write_text(">a\nACGT\n>b\nGGGG\nCC\n")
    assert BarCodeLibrary(str(fa)).barcodes == {"ACGT", "GGGGCC"}
    lib2 = BarCodeLibrary()
    lib2.load_from_list(["A", "A", "C"])
    lib2.add("G")
    lib2.remove("A")
    assert lib2.barcodes == {"C", "G"}
    from barcoder_b200 import BarCodeLibraryError
    with pytest.raises(BarCodeLibraryError):
        BarCodeLibrary(str(tsv), column="nope")
    with pytest.raises(BarCodeLibraryError):
        BarCodeLibrary(str(tsv))
    with pytest.raises(BarCodeLibraryError):
        BarCodeLibrary(str(tmp_path / "lib.txt"))


def test_crispri_library_tables():
    recs = {"c": seqio.SeqRecord(seqio.Seq("A" * 100), id="c")}
    pf = PAMFinder(recs, "NGG", "downstream")
    rows = [
        # barcode b1: one site, source + one gene on '+'
        dict(Chromosome="c", Start=10, End=30, Mapped=True, Strand="+", Barcode="b1", Mismatches=0, PAM="AGG",
             Targeting=True, Start_b=0, End_b=100, Strand_b="+", Type="source", Locus_Tag=None, Gene=None),
        dict(Chromosome="c", Start=10, End=30, Mapped=True, Strand="+", Barcode="b1", Mismatches=0, PAM="AGG",
             Targeting=True, Start_b=5, End_b=25, Strand_b="+", Type="gene", Locus_Tag="g1", Gene="a"),
        # barcode b2: two sites -> not source-unique after the first
        dict(Chromosome="c", Start=40, End=60, Mapped=True, Strand="-", Barcode="b2", Mismatches=1, PAM="TGG",
             Targeting=True, Start_b=0, End_b=100, Strand_b="+", Type="source", Locus_Tag=None, Gene=None),
        dict(Chromosome="c", Start=40, End=60, Mapped=True, Strand="-", Barcode="b2", Mismatches=1, PAM="TGG",
             Targeting=True, Start_b=50, End_b=90, Strand_b="-", Type="gene", Locus_Tag="g2", Gene="b"),
        dict(Chromosome="c", Start=70, End=90, Mapped=True, Strand="+", Barcode="b2", Mismatches=0, PAM="TTT",
             Targeting=False, Start_b=50, End_b=90, Strand_b="-", Type="gene", Locus_Tag="g2", Gene="b"),
    ]
    df = pd.DataFrame(rows)
    df.attrs["pam_key"] = ("NGG", "class-api")
    lib = CRISPRiLibrary(df, pf)
    assert list(lib.source_unique_targets["Barcode"]) == ["b1", "b2"]
    mt = lib.mapped_targets
    assert list(mt["Locus_Tag"]) == ["g1", "g2"]
    assert list(mt["Offset"]) == [5, 30] and list(mt["Overlap"]) == [15, 10]
    assert list(lib.unique_targets["Barcode"]) == ["b1", "b2"]
    assert len(lib.unambiguous_targets) == 2
    # without fused columns the finder's string functions are used
    df2 = df.drop(columns=["PAM", "Targeting"])
    lib2 = CRISPRiLibrary(df2, pf)
    assert not lib2.targets_df["Targeting"].any()  # the record is all 'A': no NGG anywhere


def test_bowtie_runner_call_order_errors(plasmids):
    with BowtieRunner() as b:
        with pytest.raises(BowtieError) as ei:
            b.create_index()
        assert "fasta_path" in ei.value.message
        with pytest.raises(BowtieError):
            b.align(1)
        b.make_fasta(plasmids)
        assert os.path.getsize(b.fasta_path) > 140000
        b.make_fastq(["ACGT", "TTTT"])
        b.make_fastq(["GGGG"])
        assert b._reads == ["ACGT", "TTTT", "GGGG"]
        assert open(b.fastq_path).read().count("\n") == 12
        assert b.sam_path.endswith(".sam") and b.index_path in b.sam_path
        import torch
        if not torch.cuda.is_available():
            with pytest.raises(BowtieError) as ei:
                b.create_index()  # no CPU fallback
            assert ei.value.message == "Failed to index"
    assert not os.path.exists(b.fasta_path)  # temp dir removed on exit


def test_overlap_join_with_whole_contig_source_feature_stays_linear():
    """GenBankParser.ranges always emits a whole-contig `source` interval; it must not widen the
    candidate window of the thousands of short gene intervals (quadratic blow-up)."""
    import time
    rng = np.random.default_rng(1)
    n_genes, n_hits, G = 4500, 60_000, 4_600_000
    gs = np.sort(rng.integers(0, G - 3000, n_genes))
    b = pd.DataFrame({"Chromosome": "c1", "Start": np.concatenate([[0], gs]),
                      "End": np.concatenate([[G], gs + rng.integers(200, 3000, n_genes)]),
                      "Strand": "+", "Type": ["source"] + ["gene"] * n_genes})
    hs = rng.integers(0, G - 20, n_hits)
    a = pd.DataFrame({"Chromosome": "c1", "Start": hs, "End": hs + 20, "Strand": "+"})
    t0 = time.time()
    li, ri = overlap_pairs(a, b)
    assert time.time() - t0 < 5.0
    assert (ri == 0).sum() == n_hits                      # every hit joins the source feature
    assert len(li) < 3 * n_hits                           # and only the genes it really overlaps
    sub = np.arange(0, n_hits, 97)
    bs, be = b["Start"].to_numpy(), b["End"].to_numpy()
    want = {(int(i), int(j)) for i in sub for j in np.nonzero((bs < hs[i] + 20) & (be > hs[i]))[0]}
    got = {(int(i), int(j)) for i, j in zip(li, ri) if i % 97 == 0}
    assert got == want


def test_genbank_wrapped_contig_line(tmp_path):
    """A CONTIG join(...) that wraps over indented lines is not feature text."""
    text = (
        "LOCUS       TEST1                     24 bp    DNA     linear   BCT 01-JAN-2000\n"
        "DEFINITION  test record.\n"
        "VERSION     TEST1.1\n"
        "FEATURES             Location/Qualifiers\n"
        "     source          1..24\n"
        "                     /organism=\"x\"\n"
        "     gene            3..10\n"
        "                     /locus_tag=\"T_1\"\n"
        "CONTIG      join(AAAA01000001.1:1..1000,gap(100),AAAA01000002.1:1..2000,\n"
        "            gap(50),AAAA01000003.1:1..500)\n"
        "ORIGIN\n"
        "        1 acgtacgtac gtacgtacgt acgt\n"
        "//\n")
    path = tmp_path / "wrapped.gb"
    path.write_text(text)
    recs = list(seqio.read_genbank(str(path)))
    assert len(recs) == 1 and recs[0].id == "TEST1.1" and str(recs[0].seq) == "ACGT" * 6
    assert [f.type for f in recs[0].features] == ["source", "gene"]


@pytest.mark.parametrize("pam,direction", [("NGNC", "downstream"), ("NGG", "upstream"), ("", "downstream"),
                                           ("NNNN", "downstream")])
def test_targets_rows_vectorised_equal_literal_loop(plasmids, cn32_spacers, pam, direction):
    """targets.build_rows (column operations) == targets.build_rows_loop (the literal per-alignment
    restatement of targets.py:310-464) on oracle hits over the circular plasmids with their 100 kb
    overhang: every intermediate column (target, coords, diff, type ...) and the shaped result."""
    from barcoder_b200 import targets
    from oracle import oracle
    sp = cn32_spacers[-1500:] + ["ACGTACGTACGTACGTACGTACGTACGTACGT", "acgtnacgtacgtacgtacg"]
    ids = list(plasmids)
    true_len = {r: len(plasmids[r].seq) for r in ids}
    topo = [str(plasmids[r].seq) + str(plasmids[r].seq)[:targets.OVERHANG] for r in ids]
    genes = targets.gene_intervals(plasmids)
    parts, by_len = [], {}
    for i, s in enumerate(sp):
        by_len.setdefault(len(s), []).append(i)
    for L, idx in sorted(by_len.items()):
        h = oracle.search(topo, [sp[i].upper() for i in idx], 2)
        h["spacer_id"] = np.asarray(idx)[h["spacer_id"]].astype(np.uint32)
        parts.append(h)
    hits = np.concatenate(parts)
    hits = hits[np.lexsort((hits["meta"] & 1, hits["gpos"], hits["spacer_id"]))]
    off = oracle.concat_genome(topo)[1].astype(np.int64)
    a = targets.build_rows(hits, off, sp, sp, ids, topo, true_len, genes, pam, direction)
    b = targets.build_rows_loop(hits, off, sp, sp, ids, topo, true_len, genes, pam, direction)
    cols = sorted(set(a.columns) | set(b.columns))

    def rows(df):
        df = df.reindex(columns=cols)

        def norm(v):  # the loop builds object columns that pandas widens to float when None is mixed in
            if pd.isna(v):
                return ""
            return str(int(v)) if isinstance(v, (float, np.floating)) and float(v).is_integer() else str(v)
        return sorted(tuple(norm(v) for v in r) for r in df.itertuples(index=False))
    assert len(a) == len(b) > 1000 and rows(a) == rows(b)
    fa, fb = targets.shape_results(a, true_len), targets.shape_results(b, true_len)
    assert list(fa.columns) == list(fb.columns)
    key = list(fa.columns)
    assert fa.sort_values(key).reset_index(drop=True).astype(str).equals(fb.sort_values(key).reset_index(drop=True).astype(str))


@pytest.mark.parametrize("with_arrow", [True, False])
def test_bowtierunner_string_helpers(with_arrow, monkeypatch):
    """The Arrow-backed column builders of BowtieRunner and their plain-Python fallbacks give the same values."""
    import importlib
    br = importlib.import_module("barcoder_b200.BowtieRunner")
    if not with_arrow:
        monkeypatch.setattr(br, "_pa", None)
        monkeypatch.setattr(br, "_pc", None)
    elif br._pa is None:
        pytest.skip("pyarrow not installed")
    reads = ["acgtACGTacgtACGTacgt", "TTTTTTTTTTTTTTTTTTTT", "GGGGGGGGGGGGGGGGGGGG", "acgtn", "ACGTACGTAC"]
    upper, lens = br._upper_and_lengths(reads)
    assert lens.tolist() == [20, 20, 20, 5, 10]
    as_list = upper.to_pylist() if hasattr(upper, "to_pylist") else list(upper)
    assert as_list == [r.upper() for r in reads]
    idx = np.array([2, 0, 0, 1], dtype=np.int64)
    col = br._take_str(upper, idx)
    assert [str(x) for x in col] == [reads[i].upper() for i in idx]
    assert [str(x) for x in br._take_str(["+", "-"], np.array([1, 0, 1]))] == ["-", "+", "-"]
    rows = br._rows_of(upper, np.array([0, 1, 2]), 20)
    if isinstance(rows, np.ndarray):
        assert rows.shape == (3, 20) and bytes(rows[0]).decode() == reads[0].upper()
    else:
        assert rows == [r.upper() for r in reads[:3]]
    df = pd.DataFrame({"Barcode": br._take_str(upper, idx), "n": np.arange(4)})
    assert df["Barcode"].tolist() == [reads[i].upper() for i in idx]
