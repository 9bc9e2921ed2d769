"""Script-path record layer: the JSON/TSV target records of the reference's `targets.py`,
produced from the CUDA search instead of `bowtie -k 100` + pysam.

Mirrors, with file:line into the reference's targets.py:
  * topological genome: circular records are searched as seq + seq[:100_000] and coordinates are
    folded back modulo the true length (:35-56, :380-384);
  * PAM rule: direction-aware, bounds-checked against the topological length, `N` = any character
    (:219-307); alignments whose PAM does not match become non-targeting rows (:350-352);
  * per alignment: spacer, len, target (reference bases in spacer orientation, lower-case at
    mismatches as pysam reports them), mismatches, chr, tar_start/tar_end, sp_dir, pam, coords,
    type, diff (:354-410);
  * one row per overlapping gene with locus_tag, gene, offset, overlap, tar_dir (:412-462), genes
    spanning the origin extended past the end (:99-128);
  * post-processing: de-duplication, off-target filter, sort, per-spacer site/gene/intergenic
    counts, `note`, conditional columns, Int64 casts (:607-694).

Difference kept on purpose: the reference caps bowtie at 100 alignments per read (`-k 100`, :499),
choosing which ones arbitrarily; this path reports all of them (`-a`), which is also what the
class API does (BowtieRunner.py:114).
"""
from __future__ import annotations

import argparse
import sys

import numpy as np
import pandas as pd

from . import _native, seqio
from .seqio import reverse_complement

OVERHANG = 100_000


def gene_intervals(records):
    """Gene intervals the way create_locus_map lays them out (:76-165): one interval per location
    part; a gene whose parts touch both ends of a circular record becomes ONE interval
    [start_of_last_part, len + end_of_first_part)."""
    out = {}
    for rid, rec in records.items():
        n = len(rec.seq)
        rows = []
        for f in rec.features:
            if f.type != "gene":
                continue
            tag = f.qualifiers.get("locus_tag", [None])[0]
            name = f.qualifiers.get("gene", [None])[0]
            parts = f.location.parts
            strand = f.location.strand
            if len(parts) > 1 and any(p.start == 0 or p.end == n for p in parts):
                tail = next(p for p in parts if p.end == n)
                head = next(p for p in parts if p.start == 0)
                rows.append((int(tail.start), int(head.end) + n, tag, name, strand))
            else:
                for p in parts:
                    rows.append((int(p.start), int(p.end), tag, name, strand))
        out[rid] = rows
    return out


def get_coords(tar_start, tar_end, chrom_length):
    s = tar_start % chrom_length
    e = tar_end % chrom_length if tar_end % chrom_length != 0 else chrom_length
    return f"({s}..{chrom_length}, 0..{e})" if s > e else f"{s}..{e}"


def _pam_ok(pam, extracted):
    if not extracted:
        return False
    if pam == "N" * len(pam) or not pam:
        return True
    return len(extracted) >= len(pam) and all(p == "N" or p == c for p, c in zip(pam, extracted))


def find_targets(spacers, records, pam, mismatches, pam_direction="downstream", device=0, names=None):
    """spacers: list[str]; names: optional read names (default: the spacer itself, what
    create_fake_topological_fastq yields for an unnamed FASTA).  Returns the final DataFrame of
    targets.py (before printing)."""
    names = list(names) if names is not None else list(spacers)
    ids = list(records)
    true_len = {rid: len(records[rid].seq) for rid in ids}
    topo = []
    for rid in ids:
        s = str(records[rid].seq)
        circular = records[rid].annotations.get("topology") == "circular"
        topo.append(s + s[:OVERHANG] if circular else s)
    genes = gene_intervals(records)
    pam = pam or ""
    k = int(mismatches)

    by_len = {}
    for i, s in enumerate(spacers):
        by_len.setdefault(len(s), []).append(i)
    parts = []
    with _native.Searcher(device) as srch:
        srch.set_genome(topo)
        off = np.asarray(srch.contig_offsets, dtype=np.int64)
        # the device evaluates the same rule for full, unambiguous PAMs; the rest is resolved below
        device_pam = pam if pam and pam.isalpha() and len(pam) <= 8 else ""
        srch.set_pam(device_pam, pam_direction)
        for L, idx in sorted(by_len.items()):
            if L < 1 or L > 32:
                raise ValueError(f"spacer length {L} is outside 1..32")
            idx = np.asarray(idx, dtype=np.int64)
            srch.set_library([spacers[i].upper() for i in idx])
            srch.search(k)
            h = srch.hits()
            h["spacer_id"] = idx[h["spacer_id"]].astype(np.uint32)
            parts.append(h)
    hits = np.concatenate(parts) if parts else np.zeros(0, dtype=_native.HIT_DTYPE)
    order = np.lexsort((hits["meta"] & 1, hits["gpos"], hits["spacer_id"]))
    hits = hits[order]

    rows = []
    P = len(pam)
    for h in hits:
        sid = int(h["spacer_id"])
        spacer = spacers[sid].upper()
        L = len(spacer)
        ci = int(np.searchsorted(off[1:], int(h["gpos"]), side="right"))
        rid = ids[ci]
        ref_start = int(h["gpos"]) - int(off[ci])
        ref_end = ref_start + L
        meta = int(h["meta"])
        minus = bool(meta & 1)
        base = {"name": names[sid], "spacer": spacer, "len": L}
        extracted = None
        if pam:
            right = (pam_direction == "downstream") != minus
            a = ref_end if right else ref_start - P
            if a >= 0 and a + P <= len(topo[ci]):
                extracted = topo[ci][a:a + P].upper()
                if minus:
                    extracted = reverse_complement(extracted)
            if not _pam_ok(pam, extracted):
                rows.append(base)  # alignment without a matching PAM -> non-targeting row (:350-352)
                continue
        window = topo[ci][ref_start:ref_end].upper()
        target = reverse_complement(window) if minus else window
        mask = int(h["mm_mask"])
        mm_pos = [i for i in range(L) if mask >> i & 1]
        target = "".join(ch.lower() if i in mm_pos else ch for i, ch in enumerate(target))
        n = true_len[rid]
        tar_start, tar_end = ref_start % n, ref_end % n
        if tar_end < tar_start:
            tar_start -= n
        nmm = (meta >> 1) & 3
        diff = ",".join(f"{target[i]}{i + 1}{spacer[i]}" for i in mm_pos) or None
        rec = dict(base, target=target, mismatches=nmm, chr=rid, tar_start=tar_start, tar_end=tar_end,
                   sp_dir="R" if minus else "F", pam=extracted, coords=get_coords(tar_start, tar_end, n),
                   type="mismatch" if nmm else "perfect", diff=diff)
        lo = max(tar_start, 0)
        found = [(fs, fe, tag, gname, strand) for fs, fe, tag, gname, strand in genes[rid]
                 if fs < tar_end and lo < fe]
        if not found:
            rows.append(dict(rec, locus_tag=None, offset=None, overlap=None, tar_dir=None))
            continue
        for fs, fe, tag, gname, strand in set(found):
            tdir = "F" if strand == 1 else "R" if strand == -1 else None
            offset = tar_start - fs if tdir == "F" else fe - tar_end if tdir == "R" else None
            ov_s, ov_e = max(tar_start, fs), min(tar_end, fe)
            rows.append(dict(rec, locus_tag=tag, gene=gname if gname else tag, offset=offset,
                             overlap=ov_e - ov_s if ov_s < ov_e else 0, tar_dir=tdir))
    aligned = {int(s) for s in hits["spacer_id"]} if len(hits) else set()
    for sid in range(len(spacers)):  # reads without any alignment (flag-4 SAM lines, :366-368)
        if sid not in aligned:
            rows.append({"name": names[sid], "spacer": spacers[sid].upper(), "len": len(spacers[sid])})
    if not rows:
        raise RuntimeError("No results were returned from the search. Check your input files and parameters.")
    return shape_results(pd.DataFrame(rows), true_len)


def shape_results(results, seq_lens):
    """targets.py:607-694 on the row table."""
    for col in ("target", "mismatches", "chr", "tar_start", "tar_end", "sp_dir", "pam", "coords", "type", "diff",
                "locus_tag", "gene", "offset", "overlap", "tar_dir"):
        if col not in results.columns:
            results[col] = None
    results = results.drop_duplicates()
    targeting = results[results["target"].notna()]["spacer"].unique()
    results = results[~(results["target"].isna() & results["spacer"].isin(targeting))]
    results = results.copy()
    results["min_tar"] = [ts - seq_lens[c] if (pd.notna(ts) and ts > te) else ts
                          for ts, te, c in zip(results["tar_start"], results["tar_end"], results["chr"])]
    results = results.sort_values(by=["chr", "min_tar", "spacer"])
    spacers_seen = results[["name", "spacer"]].drop_duplicates().groupby("spacer").size()
    results = results.drop("name", axis=1).drop_duplicates()
    has_t = results["target"].notnull()
    results.loc[has_t, "site"] = results.loc[has_t, "chr"].astype(str) + "_" + results.loc[has_t, "coords"].astype(str)
    site_counts = results.groupby("spacer")["site"].nunique()
    gene_counts = results.loc[results["locus_tag"].notnull(), "spacer"].value_counts()
    intergenic = results.loc[results["locus_tag"].isnull() & has_t, "spacer"].value_counts()
    note = pd.DataFrame({"count": spacers_seen, "sites": site_counts, "genes": gene_counts,
                         "intergenic": intergenic}).fillna(0).astype(int)

    def make_note(r):
        if r["sites"] <= 0:
            return "non-targeting"
        p = [f"{r['sites']} {'site' if r['sites'] == 1 else 'sites'}"]
        if r["genes"] > 0:
            p.append(f"{r['genes']} {'gene' if r['genes'] == 1 else 'genes'}")
        if r["intergenic"] > 0:
            p.append(f"{r['intergenic']} intergenic")
        return ", ".join(p)

    note["note"] = note.apply(make_note, axis=1)
    results = results.merge(note, left_on="spacer", right_index=True, how="left")
    cols = ["spacer", "locus_tag", "gene", "chr"]
    if not (results["count"] == 1).all():
        cols.append("count")
    if not (results["pam"].isnull().all() or results["pam"].nunique() == 1):
        cols.append("pam")
    if not (results["mismatches"] == 0).all():
        cols.append("mismatches")
    cols += ["target", "tar_start", "tar_end", "offset", "overlap", "sp_dir", "tar_dir", "note"]
    final = results.reindex(columns=cols)
    for col in ("mismatches", "offset", "overlap", "tar_start", "tar_end"):
        if col in final.columns:
            final[col] = final[col].astype("Int64")
    return final.reset_index(drop=True)


def main(argv=None):
    ap = argparse.ArgumentParser(description="Map barcodes to a circular genome (CUDA search)")
    ap.add_argument("sgrna_file", help="FASTA of spacers")
    ap.add_argument("genome_file", help="GenBank genome")
    ap.add_argument("pam")
    ap.add_argument("mismatches", type=int)
    ap.add_argument("--pam_direction", choices=["upstream", "downstream"], default="downstream")
    ap.add_argument("--json", action="store_true", default=False)
    args = ap.parse_args(argv)
    recs = list(seqio.read_fasta(args.sgrna_file))
    records = seqio.genbank_to_dict(args.genome_file)
    final = find_targets([str(r.seq) for r in recs], records, args.pam, args.mismatches, args.pam_direction,
                         names=[r.id for r in recs])
    if args.json:
        print(final.to_json(orient="records", indent=4))
    else:
        final.to_csv(sys.stdout, sep="\t", index=False, na_rep="None")


if __name__ == "__main__":
    main()
