#!/usr/bin/env python
"""Regenerates tests/golden/ from the reference checkout (run in the build container only;
/root/reference does not exist on the GPU box, which uses the committed outputs).

Inputs (read-only, data files - no reference source code is copied):
  /root/reference/GCA_003054575.1/CP02371{6,7,8,9}.dna   four Zymomonas ZM4 plasmids (SnapGene)
  /root/reference/Example_Libraries/CN-32-zmo.tsv        the reference's only output fixture

Outputs:
  zmo_plasmids.gb         the four plasmids re-serialised as GenBank (ids CP02371x.1, circular,
                          source + gene features) - exercises the GenBank reader too
  cn32_spacers.txt        the 9,503 unique `spacer` values of the fixture, first-seen order
  cn32_plasmid_rows.tsv   the 772 fixture rows whose chr is one of the four plasmids (G1/G3)
  g2_known_answer.tsv     SURVEY.md section 8c G2: every <=2-mismatch hit of those spacers on the
                          plasmids (linear, both strands) as canonical tuples
  g2_known_answer.sha256  sha256 of that tuple list; must equal the survey's value
"""
import csv
import hashlib
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from barcoder_b200.seqio import read_snapgene, write_genbank  # noqa: E402
from oracle import oracle  # noqa: E402

REF = "/root/reference"
SURVEY_SHA = "5dbbc9c4be09b08acddf816ffebfa751eff16374e50da6054f2854c03aa493da"
PLASMIDS = ["CP023716", "CP023717", "CP023718", "CP023719"]


def main():
    recs = []
    for acc in PLASMIDS:
        r = read_snapgene(f"{REF}/GCA_003054575.1/{acc}.dna", record_id=f"{acc}.1")
        r.features = [f for f in r.features if f.type in ("source", "gene")]
        for f in r.features:  # keep the qualifiers the path reads
            f.qualifiers = {k: v for k, v in f.qualifiers.items() if k in ("locus_tag", "gene", "organism")}
        recs.append(r)
    write_genbank(recs, os.path.join(HERE, "zmo_plasmids.gb"))

    with open(f"{REF}/Example_Libraries/CN-32-zmo.tsv") as h:
        rows = list(csv.reader(h, delimiter="\t"))
    header, rows = rows[0], rows[1:]
    col = {c: i for i, c in enumerate(header)}
    spacers, seen = [], set()
    for r in rows:
        s = r[col["spacer"]]
        if s not in seen:
            seen.add(s)
            spacers.append(s)
    with open(os.path.join(HERE, "cn32_spacers.txt"), "w") as h:
        h.write("\n".join(spacers) + "\n")
    ids = {r.id for r in recs}
    with open(os.path.join(HERE, "cn32_plasmid_rows.tsv"), "w") as h:
        h.write("\t".join(header) + "\n")
        for r in rows:
            if r[col["chr"]] in ids:
                h.write("\t".join(r) + "\n")

    # G2 known answer via the exhaustive oracle, PAM = 4 nt 3' of the protospacer (NGNC fixture).
    contigs = [str(r.seq) for r in recs]
    hits = oracle.search(contigs, spacers, 2, pam="NGNC", direction="downstream", mode="brute")
    genome, off = oracle.concat_genome(contigs)
    tuples = []
    for h in hits:
        ci = int((off[1:] <= h["gpos"]).sum())
        meta = int(h["meta"])
        pam = ""
        if meta & oracle.META_PAM_FULL:
            pam = "".join("ACGT"[(meta >> (16 + 2 * i)) & 3] for i in range(4))
        tuples.append((spacers[h["spacer_id"]], recs[ci].id, int(h["gpos"] - off[ci]),
                       "-" if meta & 1 else "+", (meta >> 1) & 3, pam))
    tuples.sort()
    text = "\n".join("\t".join(str(x) for x in t) for t in tuples)
    sha = hashlib.sha256(text.encode()).hexdigest()
    with open(os.path.join(HERE, "g2_known_answer.tsv"), "w") as h:
        h.write(text + "\n")
    with open(os.path.join(HERE, "g2_known_answer.sha256"), "w") as h:
        h.write(sha + "\n")
    by = [sum(1 for t in tuples if t[4] == m) for m in range(3)]
    print(f"{len(tuples)} hits {by}; sha256 {sha}; survey match: {sha == SURVEY_SHA}")
    print("first:", tuples[0])
    return 0 if sha == SURVEY_SHA else 1


if __name__ == "__main__":
    sys.exit(main())
