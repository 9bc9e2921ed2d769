"""Host-side genome input for the search path.

Mirrors the public surface of the reference's GenBankParser module (GenBankParser.py:10-123):
`GenBankReader(filename).records`, and `GenBankParser(filename)` with `.records`, `.organisms`,
`.seq_lens`, `.topologies`, `.num_genes`, `.overhangs`, `.ranges`, `.make_fasta()` and
`.find_gene_name_for_locus()`.  Parsing is done by the Biopython-free reader in seqio.py; the
per-record summaries are computed once at construction instead of through cached properties.
"""
import pandas as pd

from .Logger import Logger
from .ranges import PyRanges
from . import seqio

_STRAND_SYMBOL = {1: "+", -1: "-"}
CIRCULAR_OVERHANG = 100_000  # targets.py:43, GenBankParser.py:63


class GenBankReader:
    """`records`: dict id -> SeqRecord, what SeqIO.to_dict(SeqIO.parse(h, "genbank")) returns."""

    def __init__(self, filename):
        self.filename = filename
        self._records = None

    @property
    def records(self):
        if self._records is None:
            self._records = seqio.genbank_to_dict(self.filename)
        return self._records


def feature_intervals(records, types=("source", "gene")):
    """Rows of the feature table the hit frame is joined with: one row per location part of every
    feature whose type is in `types` (GenBankParser.py:67-103).  Coordinates are 0-based
    half-open; Strand is '+', '-' or '.'."""
    cols = {"Chromosome": [], "Start": [], "End": [], "Strand": [], "Locus_Tag": [], "Gene": [], "Type": []}
    for rid, rec in records.items():
        for feat in rec.features:
            if feat.type not in types:
                continue
            tag = feat.qualifiers.get("locus_tag", [None])[0]
            gene = feat.qualifiers.get("gene", [None])[0]
            for part in feat.location.parts:
                cols["Chromosome"].append(rid)
                cols["Start"].append(int(part.start))
                cols["End"].append(int(part.end))
                cols["Strand"].append(_STRAND_SYMBOL.get(part.strand, "."))
                cols["Locus_Tag"].append(tag)
                cols["Gene"].append(gene)
                cols["Type"].append(feat.type)
    return pd.DataFrame(cols)


class GenBankParser(Logger):
    def __init__(self, filename):
        super().__init__()
        self.reader = GenBankReader(filename)
        self.records = self.reader.records
        recs = self.records
        self.organisms = {rid: r.annotations.get("organism") for rid, r in recs.items()}
        self.seq_lens = {rid: len(r.seq) for rid, r in recs.items()}
        self.topologies = {rid: r.annotations.get("topology") for rid, r in recs.items()}
        self.num_genes = {rid: sum(f.type == "gene" for f in r.features) for rid, r in recs.items()}
        self.overhangs = {rid: CIRCULAR_OVERHANG if t == "circular" else 0 for rid, t in self.topologies.items()}
        self._ranges = None
        self.info("Found the following records:")
        self.json(self.organisms)

    @property
    def ranges(self):
        if self._ranges is None:
            self._ranges = PyRanges(feature_intervals(self.records))
        return self._ranges

    def make_fasta(self, filename):
        with open(filename, "w") as handle:
            seqio.write_fasta(self.records.values(), handle)

    def find_gene_name_for_locus(self, locus_tag):
        """Gene name of the `gene` feature carrying this locus tag (the tag itself when the feature
        has no /gene qualifier); None when no feature matches."""
        for rec in self.records.values():
            for feat in rec.features:
                if feat.type == "gene" and feat.qualifiers.get("locus_tag", [None])[0] == locus_tag:
                    return feat.qualifiers.get("gene", [locus_tag])[0]
        return None
