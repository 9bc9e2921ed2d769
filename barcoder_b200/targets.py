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


def search_hits(spacers, topo, pam, mismatches, pam_direction="downstream", device=0):
    """The CUDA search over the topological contigs -> (hit records sorted by (spacer, position,
    strand) on the device, contig offsets)."""
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
            srch.sort_hits("canonical")
            h = srch.hits()
            h["spacer_id"] = idx[h["spacer_id"]].astype(np.uint32)
            parts.append(h)
    hits = np.concatenate(parts) if parts else np.zeros(0, dtype=_native.HIT_DTYPE)
    if len(parts) > 1:
        hits = hits[np.lexsort((hits["meta"] & 1, hits["gpos"], hits["spacer_id"]))]
    return hits, off


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
    hits, off = search_hits(spacers, topo, pam, mismatches, pam_direction, device)
    rows = build_rows(hits, off, spacers, names, ids, topo, true_len, genes, pam, pam_direction)
    if not len(rows):
        raise RuntimeError("No results were returned from the search. Check your input files and parameters.")
    return shape_results(rows, true_len)


_COMP = np.zeros(256, dtype=np.uint8)
for _a, _b in zip(b"ACGTNacgtn", b"TGCANtgcan"):
    _COMP[_a] = _b
for _c in range(256):
    if _COMP[_c] == 0:
        _COMP[_c] = _c
_UPPER = np.arange(256, dtype=np.uint8)
_UPPER[ord("a"):ord("z") + 1] -= 32
_LOWER = np.arange(256, dtype=np.uint8)
_LOWER[ord("A"):ord("Z") + 1] += 32


def _strings(arr):
    """uint8 [n, w] -> object array of Python str (one C-level view + decode, no per-row loop)."""
    n, w = arr.shape
    if n == 0 or w == 0:
        return np.full(n, "", dtype=object)
    return np.char.decode(np.ascontiguousarray(arr).view(f"S{w}").ravel(), "ascii").astype(object)


def build_rows(hits, off, spacers, names, ids, topo, true_len, genes, pam, pam_direction):
    """parse_sam_output (targets.py:310-464) on the hit table, as column operations: PAM extraction
    and test, target string (lower case at mismatches), coordinates folded modulo the true length,
    diff, one row per overlapping gene.  Returns the row DataFrame shape_results expects.
    build_rows_loop below is the literal per-alignment restatement the tests compare against."""
    from .ranges import overlap_pairs
    P = len(pam)
    n_sp = len(spacers)
    sp_upper = np.asarray([s.upper() for s in spacers], dtype=object)
    sp_len = np.fromiter((len(s) for s in spacers), dtype=np.int64, count=n_sp)
    names_arr = np.asarray(names, dtype=object)
    topo_u8 = [np.frombuffer(t.encode("ascii"), dtype=np.uint8) for t in topo]
    topo_len = np.asarray([len(t) for t in topo], dtype=np.int64)
    n = len(hits)
    sid = hits["spacer_id"].astype(np.int64)
    gpos = hits["gpos"].astype(np.int64)
    ci = np.searchsorted(off[1:], gpos, side="right")
    L = sp_len[sid]
    ref_start = gpos - off[ci]
    ref_end = ref_start + L
    minus = (hits["meta"] & 1) != 0
    nmm = ((hits["meta"] >> 1) & 3).astype(np.int64)
    cat = np.concatenate(topo_u8) if topo_u8 else np.zeros(0, np.uint8)   # contigs back to back, like gpos
    base = off[ci]

    # ---- PAM: direction-aware slice, bounds-checked against the topological contig (:227-307)
    extracted = np.full(n, None, dtype=object)
    pam_ok = np.ones(n, dtype=bool)
    if P:
        right = (pam_direction == "downstream") != minus
        a = np.where(right, ref_end, ref_start - P)
        inb = (a >= 0) & (a + P <= topo_len[ci])
        idx = np.where(inb, base + a, 0)[:, None] + np.arange(P)[None, :]
        pam_arr = _UPPER[cat[idx]] if n else np.zeros((0, P), np.uint8)
        pam_arr = np.where(minus[:, None], _COMP[pam_arr[:, ::-1]], pam_arr)
        ex = _strings(pam_arr)
        extracted[inb] = ex[inb]
        if pam == "N" * P:
            pam_ok = inb & (np.asarray([bool(e) for e in ex]) if n else np.zeros(0, bool))
        else:
            want = np.frombuffer(pam.encode("ascii"), dtype=np.uint8)
            pam_ok = inb & ((pam_arr == want[None, :]) | (want[None, :] == ord("N"))).all(axis=1)
    frames = []
    base_cols = pd.DataFrame({"name": names_arr[sid], "spacer": sp_upper[sid], "len": L})
    if (~pam_ok).any():   # alignment without a matching PAM -> non-targeting row (:350-352)
        frames.append(base_cols[~pam_ok])
    keep = np.nonzero(pam_ok)[0]

    # ---- alignments with a PAM: target, coordinates, diff (:354-410)
    if len(keep):
        k_sid, k_ci, k_L = sid[keep], ci[keep], L[keep]
        k_minus, k_start, k_end = minus[keep], ref_start[keep], ref_end[keep]
        mask = hits["mm_mask"][keep].astype(np.int64)
        target = np.empty(len(keep), dtype=object)
        diff = np.full(len(keep), None, dtype=object)
        for Lv in np.unique(k_L):           # one pass per spacer length (usually one)
            g = np.nonzero(k_L == Lv)[0]
            idx = (base[keep][g] + k_start[g])[:, None] + np.arange(Lv)[None, :]
            win = _UPPER[cat[idx]]
            win = np.where(k_minus[g][:, None], _COMP[win[:, ::-1]], win)
            mm = ((mask[g][:, None] >> np.arange(Lv)[None, :]) & 1).astype(bool)
            tgt = np.where(mm, _LOWER[win], win)
            target[g] = _strings(tgt)
            has = np.nonzero(mm.any(axis=1))[0]
            if len(has):                     # "{target_nt}{1-based position}{spacer_nt}", comma separated (:184-190)
                sp_arr = np.frombuffer("".join(sp_upper[k_sid[g][has]]).encode("ascii"), np.uint8).reshape(-1, Lv)
                rr, cc = np.nonzero(mm[has])
                piece = np.char.add(np.char.add(_strings(tgt[has][rr, cc][:, None]).astype(str), (cc + 1).astype(str)),
                                    _strings(sp_arr[rr, cc][:, None]).astype(str))
                starts = np.concatenate([[0], np.nonzero(np.diff(rr))[0] + 1])
                joined = [",".join(piece[s0:s1]) for s0, s1 in zip(starts, list(starts[1:]) + [len(rr)])]
                diff[g[has]] = joined
        rid = np.asarray(ids, dtype=object)[k_ci]
        tl = np.asarray([true_len[r] for r in ids], dtype=np.int64)[k_ci]
        tar_start, tar_end = k_start % tl, k_end % tl
        tar_start = np.where(tar_end < tar_start, tar_start - tl, tar_start)
        s_mod = tar_start % tl
        e_mod = np.where(tar_end % tl != 0, tar_end % tl, tl)
        coords = np.where(s_mod > e_mod,
                          np.char.add(np.char.add(np.char.add("(", s_mod.astype(str)), np.char.add("..", tl.astype(str))),
                                      np.char.add(np.char.add(", 0..", e_mod.astype(str)), ")")),
                          np.char.add(np.char.add(s_mod.astype(str), ".."), e_mod.astype(str))).astype(object)
        k_nmm = nmm[keep]
        rec = pd.DataFrame({
            "name": names_arr[k_sid], "spacer": sp_upper[k_sid], "len": k_L, "target": target, "mismatches": k_nmm,
            "chr": rid, "tar_start": tar_start, "tar_end": tar_end, "sp_dir": np.where(k_minus, "R", "F"),
            "pam": extracted[keep] if P else None, "coords": coords,
            "type": np.where(k_nmm > 0, "mismatch", "perfect"), "diff": diff})
        # ---- one row per overlapping gene (:412-462); identical gene tuples collapse (set(found))
        gtab = []
        for c, r in enumerate(ids):
            for fs, fe, tag, gname, strand in set(genes[r]):
                gtab.append((c, fs, fe, tag, gname, strand))
        li = np.zeros(0, dtype=np.int64)
        if gtab:
            gdf = pd.DataFrame({"Chromosome": [g[0] for g in gtab], "Start": [g[1] for g in gtab], "End": [g[2] for g in gtab]})
            hdf = pd.DataFrame({"Chromosome": k_ci, "Start": np.maximum(tar_start, 0), "End": tar_end})
            li, ri = overlap_pairs(hdf, gdf)
        hit_has_gene = np.zeros(len(keep), dtype=bool)
        hit_has_gene[li] = True
        if (~hit_has_gene).any():
            nog = rec[~hit_has_gene].copy()
            for col in ("locus_tag", "offset", "overlap", "tar_dir"):
                nog[col] = None
            frames.append(nog)
        if len(li):
            gj = rec.iloc[li].reset_index(drop=True)
            fs = np.asarray([gtab[j][1] for j in ri], dtype=np.int64)
            fe = np.asarray([gtab[j][2] for j in ri], dtype=np.int64)
            tag = np.asarray([gtab[j][3] for j in ri], dtype=object)
            gname = np.asarray([gtab[j][4] for j in ri], dtype=object)
            strand = np.asarray([gtab[j][5] if gtab[j][5] in (1, -1) else 0 for j in ri], dtype=np.int64)
            ts, te = tar_start[li], tar_end[li]
            tdir = np.where(strand == 1, "F", np.where(strand == -1, "R", None)).astype(object)
            offset = np.where(strand == 1, ts - fs, fe - te).astype(object)
            offset[strand == 0] = None
            ov_s, ov_e = np.maximum(ts, fs), np.minimum(te, fe)
            gj["locus_tag"] = tag
            gj["gene"] = np.where(pd.isna(gname) | (gname == ""), tag, gname)
            gj["gene"] = [g if g else t for g, t in zip(gname, tag)]
            gj["offset"] = offset
            gj["overlap"] = np.where(ov_s < ov_e, ov_e - ov_s, 0)
            gj["tar_dir"] = tdir
            frames.append(gj)
    aligned = np.zeros(n_sp, dtype=bool)
    aligned[sid] = True
    missing = np.nonzero(~aligned)[0]
    if len(missing):   # reads without any alignment (flag-4 SAM lines, :366-368)
        frames.append(pd.DataFrame({"name": names_arr[missing], "spacer": sp_upper[missing], "len": sp_len[missing]}))
    if not frames:
        return pd.DataFrame()
    return pd.concat(frames, ignore_index=True)


def build_rows_loop(hits, off, spacers, names, ids, topo, true_len, genes, pam, pam_direction):
    """Literal per-alignment restatement of parse_sam_output (targets.py:310-464): what build_rows
    must reproduce (tests/test_host_api.py compares the two)."""
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
    return pd.DataFrame(rows)


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
