"""GPU tests of the reference-facing class API: the call sequence of testing_grounds.py:16-43
through BowtieRunner -> PySamParser -> join -> CRISPRiLibrary, checked against the oracle and the
reference's own string rules, and against the CN-32-zmo.tsv plasmid rows (G1/G3)."""
import csv
import os
import re

import numpy as np
import pandas as pd
import pytest

from barcoder_b200 import (BarCodeLibrary, BowtieRunner, CRISPRiLibrary, GenBankParser, GuideFinder, PAMFinder,
                           PySamParser, _native, multi_gpu, samio, seqio, synth)
from barcoder_b200.ranges import PyRanges
from oracle import oracle

pytestmark = pytest.mark.gpu


def expected_frame(records, reads, k, pam):
    """The frame the reference would build: bowtie hit set (oracle) -> PySamParser columns ->
    PAMFinder.get_pam_seq / pam_matches per row, literally with strings."""
    ids = list(records)
    contigs = [str(records[i].seq) for i in ids]
    rows = []
    by_len = {}
    for i, r in enumerate(reads):
        by_len.setdefault(len(r), []).append(i)
    for L, idx in by_len.items():
        hits = oracle.search(contigs, [reads[i].upper() for i in idx], k)
        _, off = oracle.concat_genome(contigs)
        for h in hits:
            ci = int((off[1:] <= h["gpos"]).sum())
            start = int(h["gpos"] - off[ci])
            strand = "-" if h["meta"] & 1 else "+"
            s, ok = oracle.py_pam_class_api(contigs[ci], start, start + L, strand, pam)
            rows.append((ids[ci], start, start + L, strand, reads[idx[h["spacer_id"]]].upper(),
                         int(h["meta"] >> 1) & 3, s, ok))
    return sorted(rows)


def frame_rows(df):
    df = df[df["Mapped"]]
    return sorted(zip(df["Chromosome"], df["Start"].astype(int), df["End"].astype(int), df["Strand"], df["Barcode"],
                      df["Mismatches"].astype(int), df["PAM"], df["Targeting"].astype(bool)))


def test_class_api_flow_on_plasmids(golden_dir, cn32_spacers):
    genbank = GenBankParser(os.path.join(golden_dir, "zmo_plasmids.gb"))
    barcodes = BarCodeLibrary()
    barcodes.load_from_list(cn32_spacers[-3000:] + ["ACGTACGTACGTACGTACGTACGTACGTACGT", "acgtnacgtacgtacgtacg"])
    pam = PAMFinder(genbank.records, "NGNC", "downstream")
    with BowtieRunner() as bowtie:
        bowtie.make_fasta(genbank.records)
        bowtie.make_fastq(barcodes.barcodes)
        bowtie.create_index()
        bowtie.align(num_mismatches=2, num_threads=12)
        reads = list(bowtie._reads)
        sam = PySamParser(bowtie.sam_path)
        frame = sam.ranges.df.copy()
        targets = sam.ranges.join(genbank.ranges)
        # the SAM file written for pysam users carries the same alignments
        text = PySamParser.__new__(PySamParser)
        text.filename, text._ranges = bowtie.sam_path + ".copy", None
        os.replace(bowtie.sam_path, text.filename)
        from_text = text.ranges.df
    want = expected_frame(genbank.records, reads, 2, "NGNC")
    assert frame_rows(frame) == want
    assert sorted(zip(from_text["Chromosome"], from_text["Start"], from_text["End"], from_text["Strand"],
                      from_text["Barcode"], from_text["Mismatches"])) == [w[:6] for w in want]
    guides = CRISPRiLibrary(targets.df, pam)
    assert {"PAM", "Targeting", "Type", "Locus_Tag", "Start_b", "End_b", "Strand_b"} <= set(guides.targets_df.columns)
    mt = guides.mapped_targets
    assert len(mt) and (mt["Type"] == "gene").all() and mt["Targeting"].all()
    # Offset / Overlap follow CRISPRiLibrary.py:61-83
    for r in mt.itertuples():
        off = r.Start - r.Start_b if r.Strand_b == "+" else r.End_b - r.End
        assert r.Offset == off and r.Overlap == max(min(r.End, r.End_b) - max(r.Start, r.Start_b), 0)
    assert set(guides.unique_targets["Barcode"]) <= set(guides.source_unique_targets["Barcode"])
    assert not guides.unambiguous_targets["Barcode"].duplicated().any()


def test_fixture_rows_reproduced_through_class_api(golden_dir, cn32_spacers, plasmids):
    """G1/G3: every plasmid row of CN-32-zmo.tsv comes out of the pipeline: same site, PAM, and for
    genes that do not span the origin the same offset / overlap / gene strand."""
    genbank = GenBankParser(os.path.join(golden_dir, "zmo_plasmids.gb"))
    pam = PAMFinder(genbank.records, "NGNC", "downstream")
    with BowtieRunner(write_sam=False) as bowtie:
        bowtie.make_fasta(genbank.records)
        bowtie.make_fastq(cn32_spacers)
        bowtie.create_index()
        bowtie.align(num_mismatches=0)
        targets = PySamParser(bowtie.sam_path).ranges.join(genbank.ranges)
    lib = CRISPRiLibrary(targets.df, pam)
    mt = lib.mapped_targets
    key = {}
    for r in mt.itertuples():
        key[(r.Barcode, r.Chromosome, int(r.Start), r.Strand, r.Locus_Tag)] = r
    with open(os.path.join(golden_dir, "cn32_plasmid_rows.tsv")) as h:
        rows = list(csv.DictReader(h, delimiter="\t"))
    checked = 0
    for row in rows:
        strand = "+" if row["sp_dir"] == "F" else "-"
        r = key.get((row["spacer"], row["chr"], int(row["tar_start"]), strand, row["locus_tag"]))
        assert r is not None, row
        assert r.PAM == row["pam"] and r.Mismatches == 0
        if row["locus_tag"] == "ZMO1_ZMOp36x053":
            continue  # origin-spanning gene: the script path extends it past the end (targets.py:102-128)
        assert int(row["offset"]) == r.Offset and int(row["overlap"]) == r.Overlap
        assert (row["tar_dir"] == "F") == (r.Strand_b == "+")
        checked += 1
    assert checked >= 760


def test_mixed_lengths_unmapped_and_late_pam(plasmids):
    recs = {k: plasmids[k] for k in list(plasmids)[:2]}
    seq0 = str(recs["CP023716.1"].seq)
    reads = [seq0[100:120], seq0[500:532], seq0[900:912], "ACGTTTGACCGATTAGCAGT", seq0[100:120]]
    with BowtieRunner() as b:
        b.make_fasta(recs)
        b.make_fastq(reads[:2])
        b.make_fastq(reads[2:])  # accumulates (append mode)
        b.create_index()
        b._pam = None
        from barcoder_b200._state import ACTIVE_PAM
        ACTIVE_PAM["finder"] = None
        b.align(1)
        df = PySamParser(b.sam_path).ranges.df
        full = b.frame
        lines = list(samio.read_sam(b.sam_path))
    assert "PAM" not in df.columns  # no finder known at align time
    assert (full["Mapped"] == False).sum() >= 1 and list(full[~full["Mapped"].astype(bool)]["Mismatches"]) == ["0"]
    assert sum(1 for r in lines if r.is_unmapped) == 1
    # duplicate read reported once per copy, as bowtie aligns every read independently
    assert (df["Barcode"] == reads[0]).sum() == 2 * (df.drop_duplicates()["Barcode"] == reads[0]).sum() or \
        (df["Barcode"] == reads[0]).sum() >= 2
    late = PAMFinder(recs, "NGG", "downstream")
    joined = PyRanges(df).join(GenBankParserLike(recs).ranges)
    lib = CRISPRiLibrary(joined.df, late)  # falls back to the finder's string functions
    for r in lib.targets_df.itertuples():
        s, ok = oracle.py_pam_class_api(str(recs[r.Chromosome].seq), r.Start, r.End, r.Strand, "NGG")
        assert r.PAM == s and r.Targeting == ok


class GenBankParserLike:
    def __init__(self, records):
        from barcoder_b200.GenBankParser import feature_intervals
        self.ranges = PyRanges(feature_intervals(records))


def test_guide_finder_library_targets_itself():
    """testing_grounds.py flow on a synthetic record: guides enumerated next to a PAM must all map
    back with 0 mismatches and a matching PAM."""
    genome, off = synth.random_genome(30_000, seed=3)
    rec = seqio.SeqRecord(seqio.Seq(bytes(genome).decode()), id="syn.1", description="synthetic")
    rec.features = [seqio.SeqFeature(seqio.SimpleLocation(0, 30_000, 1), "source", {}),
                    seqio.SeqFeature(seqio.SimpleLocation(1000, 9000, 1), "gene", {"locus_tag": ["g1"]}),
                    seqio.SeqFeature(seqio.SimpleLocation(12000, 20000, -1), "gene", {"locus_tag": ["g2"]})]
    records = {"syn.1": rec}
    guides = [g for g in GuideFinder(records, "NGG", "downstream", 20).find_guides_from_pam() if len(g) == 20]
    lib = BarCodeLibrary(barcodes=guides)
    pam = PAMFinder(records, "NGG", "downstream")
    with BowtieRunner(write_sam=False) as b:
        b.make_fasta(records)
        b.make_fastq(lib.barcodes)
        b.create_index()
        b.align(num_mismatches=1)
        df = PySamParser(b.sam_path).ranges.df
        targets = PySamParser(b.sam_path).ranges.join(GenBankParserLike(records).ranges)
    exact = df[(df["Mismatches"] == 0) & df["Targeting"]]
    assert set(exact["Barcode"]) == set(guides)
    out = CRISPRiLibrary(targets.df, pam)
    assert len(out.unique_targets) > 0
    assert set(out.unique_targets["Locus_Tag"]) <= {"g1", "g2"}


def test_two_shards_on_one_gpu_equal_one_shard():
    """Library sharding emulated on one GPU: two shards searched with global spacer
    ids and concatenated must equal the unsharded result."""
    genome, off = synth.random_genome(500_000, seed=41, n_contigs=3, n_fraction=0.002)
    lib = synth.random_library(30_000, 20, seed=42)
    synth.plant(lib, genome, 0.3, 3, seed=43)
    parts = []
    for r in range(2):
        lo, hi = multi_gpu.shard_bounds(len(lib), 2, r)
        with _native.Searcher(0) as s:
            s.set_genome_array(genome, off)
            s.set_library(lib[lo:hi])
            s.set_pam("NGG")
            s.set_param(_native.BC_PARAM_SPACER_ID_BASE, lo)
            s.search(3)
            parts.append(s.hits())
    with _native.Searcher(0) as s:
        s.set_genome_array(genome, off)
        s.set_library(lib)
        s.set_pam("NGG")
        s.search(3)
        whole = _native.canonical_sort(s.hits())
    merged = _native.canonical_sort(np.concatenate(parts))
    assert merged.tobytes() == whole.tobytes()


@pytest.mark.parametrize("path", [1, 2])
def test_search_and_gather_single_rank_group(path):
    """multi_gpu.search_and_gather through a one-rank NCCL group: the slice callbacks deliver every
    record exactly once (also across a hit-buffer overflow, which repeats the search)."""
    import socket
    import torch
    import torch.distributed as dist
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(0)
    dist.init_process_group("nccl", rank=0, world_size=1)
    try:
        dev = torch.device("cuda", 0)
        genome, off = synth.random_genome(800_000, seed=71, n_contigs=4, n_fraction=0.002)
        lib = synth.random_library(40_000, 20, seed=72)
        synth.plant(lib, genome, 0.3, 2, seed=73)
        with _native.Searcher(0) as s:
            s.set_genome_array(genome, off)
            s.set_library(lib)
            s.set_pam("NGG")
            s.set_param(_native.BC_PARAM_PATH, path)
            s.search(2)
            whole = _native.canonical_sort(s.hits())
            assert len(whole) > 5000
            calls = []
            s.set_slice_callback(lambda base, b, e: calls.append((b, e)))
            s.search(2)
            s.set_slice_callback(None)
            assert calls[0][0] == 0 and calls[-1][1] == len(whole)
            assert all(calls[i][1] == calls[i + 1][0] for i in range(len(calls) - 1))
            merged, per_rank = multi_gpu.search_and_gather(s, 2, dev, capacity=16)
            assert per_rank == [len(whole)]
            got = _native.canonical_sort(multi_gpu.records_from_tensor(merged))
            assert got.tobytes() == whole.tobytes()
            # peer-buffer merge (world of one: the owner's own region, filled through the hit sink)
            pg = multi_gpu.PeerGather(s, dev, len(whole) + 10)
            n = s.search(2)
            assert pg.finish(n) == [len(whole)]
            got = _native.canonical_sort(multi_gpu.records_from_tensor(pg.merged()))
            assert got.tobytes() == whole.tobytes()
            pg.close()
        with _native.Searcher(0) as s:
            s.set_genome_array(genome, off)
            s.set_library(lib)
            s.set_pam("NGG")
            s.set_param(_native.BC_PARAM_PATH, path)
            s.set_param(_native.BC_PARAM_HIT_CAPACITY, 1000)
            merged, per_rank = multi_gpu.search_and_gather(s, 2, dev)
            got = _native.canonical_sort(multi_gpu.records_from_tensor(merged))
            assert got.tobytes() == whole.tobytes()
    finally:
        dist.destroy_process_group()


def test_targets_script_records_match_fixture(golden_dir, plasmids):
    """the targets.py-shaped records (circular overhang, direction-aware PAM, per-gene rows,
    origin-spanning gene) reproduce every plasmid row of the reference's CN-32-zmo.tsv."""
    from barcoder_b200 import targets
    with open(os.path.join(golden_dir, "cn32_plasmid_rows.tsv")) as h:
        rows = list(csv.DictReader(h, delimiter="\t"))
    spacers = sorted({r["spacer"] for r in rows})
    final = targets.find_targets(spacers, plasmids, "NGNC", 0, "downstream")
    assert {"spacer", "locus_tag", "gene", "chr", "target", "tar_start", "tar_end", "offset", "overlap", "sp_dir",
            "tar_dir", "note"} <= set(final.columns)
    got = {}
    for r in final.itertuples():
        if isinstance(r.target, str):
            got[(r.spacer, r.chr, int(r.tar_start), r.locus_tag)] = r
    for row in rows:
        r = got.get((row["spacer"], row["chr"], int(row["tar_start"]), row["locus_tag"]))
        assert r is not None, row
        assert r.target.upper() == row["target"] and int(r.tar_end) == int(row["tar_end"])
        assert r.sp_dir == row["sp_dir"] and r.tar_dir == row["tar_dir"] and r.gene == row["gene"]
        assert int(r.offset) == int(row["offset"]) and int(r.overlap) == int(row["overlap"]), row
        if "pam" in final.columns:
            assert r.pam == row["pam"]
    # JSON serialises like targets.py --json
    import json
    recs = json.loads(final.to_json(orient="records"))
    assert len(recs) == len(final) and "note" in recs[0]
    # mismatching hits carry lower-case target bases and a diff-compatible mask
    mm = targets.find_targets(spacers[:200], plasmids, "NGNC", 2, "downstream")
    assert "mismatches" in mm.columns
    with_mm = mm[mm["mismatches"] > 0]
    assert len(with_mm) and all(sum(ch.islower() for ch in t) == m for t, m in zip(with_mm["target"], with_mm["mismatches"]))


def test_two_contexts_on_one_gpu_equal_one_context(golden_dir, cn32_spacers):
    """BowtieRunner(devices=[0, 0]): two search contexts (host thread each, slot-range sharding) must
    give the frame of the single-context run; devices="auto" + num_threads=1 uses one GPU; an
    explicit upstream PAM is tagged so that CRISPRiLibrary re-annotates with its own finder."""
    genbank = GenBankParser(os.path.join(golden_dir, "zmo_plasmids.gb"))
    spacers = cn32_spacers[-2500:] + ["ACGTACGTACGTACGTACGTACGTACGTACGT", "TTGACAGCTAGCTCAGTCCT"]
    frames = []
    for devices in (None, [0, 0], "auto"):
        PAMFinder(genbank.records, "NGNC", "downstream")
        with BowtieRunner(devices=devices, write_sam=False) as bowtie:
            bowtie.make_fasta(genbank.records)
            bowtie.make_fastq(spacers)
            bowtie.create_index()
            bowtie.align(num_mismatches=2, num_threads=1)
            assert len(bowtie.stats) == (4 if devices == [0, 0] else 2)   # two spacer lengths x contexts
            frames.append(frame_rows(PySamParser(bowtie.sam_path).ranges.df))
            hits = bowtie.hits
    assert frames[0] == frames[1] == frames[2] and len(frames[0]) > 500
    # the hit table is in `bowtie --best` order: per read, fewest mismatches first
    key = list(zip(hits["spacer_id"].tolist(), ((hits["meta"] >> 1) & 3).tolist(), hits["gpos"].tolist()))
    assert key == sorted(key)
    finder = PAMFinder(genbank.records, "NGNC", "upstream")
    with BowtieRunner(write_sam=False) as bowtie:
        bowtie.set_pam("NGNC", "upstream")
        bowtie.make_fasta(genbank.records)
        bowtie.make_fastq(spacers[-700:])
        bowtie.create_index()
        bowtie.align(num_mismatches=1)
        sam = PySamParser(bowtie.sam_path)
        df = sam.ranges.df
        assert df.attrs["pam_key"] == ("NGNC", "upstream")
        lib = CRISPRiLibrary(sam.ranges.join(genbank.ranges).df, finder)
    for r in lib.targets_df.itertuples():   # re-annotated with the finder's own (3' slice) rule
        assert r.PAM == finder.get_pam_seq(r) and r.Targeting == finder.pam_matches(r.PAM)
