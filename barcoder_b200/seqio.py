"""Biopython-free sequence records for the search path.

The reference builds its inputs with Biopython (`SeqIO.parse(handle, "genbank")`,
GenBankParser.py:14-18; `Seq(...).reverse_complement()`, PySamParser.py:8-9,
PAMProcessor.py:10-14).  Biopython is not part of this build, so this module
provides the small subset of that object model the path touches:

* ``Seq``            - immutable string wrapper with slicing and reverse_complement()
* ``SeqRecord``      - id / name / description / seq / annotations / features
* ``SeqFeature``     - type / location / qualifiers (qualifier values are lists)
* ``SimpleLocation`` / ``CompoundLocation`` - 0-based half-open, strand +1/-1/None
* ``read_genbank``   - GenBank flat-file reader (LOCUS topology, VERSION, ORGANISM,
                       FEATURES with complement()/join()/order(), ORIGIN)
* ``write_genbank``  - writer used for synthetic genomes and the golden plasmid fixture
* ``read_snapgene``  - SnapGene .dna reader (the only form the Zymomonas plasmids
                       survive in)
* ``read_fasta`` / ``write_fasta``
"""
from __future__ import annotations

import re
import struct
from html import unescape

_COMP = str.maketrans(
    "ACGTUMRWSYKVHDBNacgtumrwsykvhdbn", "TGCAAKYWSRMBDHVNtgcaakywsrmbdhvn"
)


def reverse_complement(s: str) -> str:
    return s.translate(_COMP)[::-1]


class Seq:
    """String-backed sequence; mirrors the Bio.Seq.Seq calls the reference makes."""

    __slots__ = ("_data",)

    def __init__(self, data=""):
        self._data = str(data)

    def __str__(self):
        return self._data

    def __repr__(self):
        d = self._data
        return f"Seq({d!r})" if len(d) <= 60 else f"Seq({d[:54]!r}...{d[-3:]!r})"

    def __len__(self):
        return len(self._data)

    def __getitem__(self, item):
        if isinstance(item, slice):
            return Seq(self._data[item])
        return self._data[item]

    def __add__(self, other):
        return Seq(self._data + str(other))

    def __radd__(self, other):
        return Seq(str(other) + self._data)

    def __eq__(self, other):
        return self._data == str(other)

    def __hash__(self):
        return hash(self._data)

    def __iter__(self):
        return iter(self._data)

    def __contains__(self, item):
        return str(item) in self._data

    def upper(self):
        return Seq(self._data.upper())

    def lower(self):
        return Seq(self._data.lower())

    def find(self, sub, *a):
        return self._data.find(str(sub), *a)

    def count(self, sub):
        return self._data.count(str(sub))

    def complement(self):
        return Seq(self._data.translate(_COMP))

    def reverse_complement(self):
        return Seq(reverse_complement(self._data))


class SimpleLocation:
    __slots__ = ("start", "end", "strand")

    def __init__(self, start, end, strand=None):
        self.start = int(start)
        self.end = int(end)
        self.strand = strand

    @property
    def parts(self):
        return [self]

    def __len__(self):
        return self.end - self.start

    def __repr__(self):
        s = {1: "(+)", -1: "(-)"}.get(self.strand, "")
        return f"[{self.start}:{self.end}]{s}"


FeatureLocation = SimpleLocation


class CompoundLocation:
    def __init__(self, parts, operator="join"):
        self.parts = list(parts)
        self.operator = operator

    @property
    def strand(self):
        strands = {p.strand for p in self.parts}
        return strands.pop() if len(strands) == 1 else None

    @property
    def start(self):
        return min(p.start for p in self.parts)

    @property
    def end(self):
        return max(p.end for p in self.parts)

    def __repr__(self):
        return f"{self.operator}{{{', '.join(map(repr, self.parts))}}}"


class SeqFeature:
    def __init__(self, location=None, type="", qualifiers=None):
        self.location = location
        self.type = type
        self.qualifiers = qualifiers if qualifiers is not None else {}

    def __repr__(self):
        return f"SeqFeature({self.type}, {self.location!r})"


class SeqRecord:
    def __init__(self, seq, id="<unknown id>", name="<unknown name>",
                 description="<unknown description>", annotations=None, features=None):
        self.seq = seq if isinstance(seq, Seq) else Seq(seq)
        self.id = id
        self.name = name
        self.description = description
        self.annotations = annotations if annotations is not None else {}
        self.features = features if features is not None else []
        self.letter_annotations = {}

    def __len__(self):
        return len(self.seq)

    def __repr__(self):
        return f"SeqRecord(id={self.id!r}, len={len(self.seq)}, features={len(self.features)})"


# --------------------------------------------------------------------------- GenBank

_LOC_TOKEN = re.compile(r"[<>]?(\d+)(?:\.\.[<>]?(\d+)|\^[<>]?(\d+))?")


def _parse_location(text: str, strand=1):
    """GenBank location string -> SimpleLocation/CompoundLocation (0-based half-open).

    Follows Biopython's conventions: ``complement(join(a..b,c..d))`` yields parts in
    reverse order with strand -1 (Bio.SeqFeature semantics the reference relies on
    in GenBankParser.py:76-86 and targets.py:99-116)."""
    text = text.strip()
    if text.startswith("complement(") and text.endswith(")"):
        inner = _parse_location(text[11:-1], -strand if strand else strand)
        if isinstance(inner, CompoundLocation):
            inner.parts = inner.parts[::-1]
        return inner
    for op in ("join", "order"):
        if text.startswith(op + "(") and text.endswith(")"):
            body = text[len(op) + 1:-1]
            parts, depth, cur = [], 0, ""
            for ch in body:
                if ch == "(":
                    depth += 1
                elif ch == ")":
                    depth -= 1
                if ch == "," and depth == 0:
                    parts.append(cur)
                    cur = ""
                else:
                    cur += ch
            if cur:
                parts.append(cur)
            locs = []
            for p in parts:
                sub = _parse_location(p, strand)
                locs.extend(sub.parts)
            return CompoundLocation(locs, op) if len(locs) > 1 else locs[0]
    if ":" in text:  # remote reference "ACC.1:1..10" - keep coordinates only
        text = text.split(":", 1)[1]
    m = _LOC_TOKEN.fullmatch(text)
    if not m:
        raise ValueError(f"Unsupported GenBank location: {text!r}")
    a = int(m.group(1))
    b = int(m.group(2) or m.group(3) or m.group(1))
    return SimpleLocation(a - 1, b, strand)


def _finish_feature(ftype, loc_text, qual_lines):
    quals: dict[str, list] = {}
    key, val = None, None

    def flush():
        if key is None:
            return
        v = val
        if v is not None and len(v) >= 2 and v[0] == '"' and v[-1] == '"':
            v = v[1:-1].replace('""', '"')
        quals.setdefault(key, []).append("" if v is None else v)

    for line in qual_lines:
        if line.startswith("/") and (val is None or not _open_quote(val)):
            flush()
            if "=" in line:
                key, val = line[1:].split("=", 1)
            else:
                key, val = line[1:], None
        else:
            if val is None:
                val = line
            elif key == "translation":
                val += line
            else:
                val += " " + line
    flush()
    return SeqFeature(_parse_location(loc_text), ftype, quals)


def _open_quote(v: str) -> bool:
    return v.startswith('"') and (len(v) == 1 or not v.endswith('"') or v.count('"') % 2 == 1)


def read_genbank(handle_or_path):
    """Yield SeqRecord objects from a GenBank flat file (one per LOCUS ... //)."""
    if isinstance(handle_or_path, (str, bytes)):
        import gzip
        opener = gzip.open if str(handle_or_path).endswith(".gz") else open
        with opener(handle_or_path, "rt") as h:
            yield from _read_genbank(h)
    else:
        yield from _read_genbank(handle_or_path)


def _read_genbank(handle):
    rec = None
    section = None
    feats = []
    cur = None  # (type, loc_text, [qual lines])
    seq_chunks = []
    last_key = None
    for raw in handle:
        line = raw.rstrip("\n")
        if line.startswith("LOCUS"):
            toks = line.split()
            rec = SeqRecord(Seq(""), id=toks[1], name=toks[1], description="")
            rec.annotations["topology"] = (
                "circular" if "circular" in toks else "linear" if "linear" in toks else None
            )
            if rec.annotations["topology"] is None:
                del rec.annotations["topology"]
            section, feats, cur, seq_chunks, last_key = "header", [], None, [], "LOCUS"
            continue
        if rec is None:
            continue
        if line.startswith("//"):
            if cur:
                feats.append(_finish_feature(*cur))
            rec.features = feats
            # Biopython upper-cases GenBank sequence.
            rec.seq = Seq("".join(seq_chunks).upper())
            if rec.description.endswith("."):
                rec.description = rec.description[:-1]
            yield rec
            rec = None
            continue
        if section == "header":
            if line.startswith("FEATURES"):
                section = "features"
                continue
            if line.startswith("ORIGIN"):
                section = "origin"
                continue
            key = line[:12].strip()
            val = line[12:].strip()
            if key:
                last_key = key
                if key == "DEFINITION":
                    rec.description = val
                elif key == "ACCESSION":
                    rec.annotations["accessions"] = val.split()
                elif key == "VERSION":
                    if val:
                        rec.id = val.split()[0]
                elif key == "ORGANISM":
                    rec.annotations["organism"] = val
                elif key == "SOURCE":
                    rec.annotations["source"] = val
            elif last_key == "DEFINITION":
                rec.description += " " + val
            continue
        if section == "features":
            if line.startswith("ORIGIN"):
                section = "origin"
                continue
            if line.startswith("CONTIG"):
                # CONTIG join(...) may wrap over several indented lines: everything up to ORIGIN / //
                # belongs to it (Biopython keeps it as an annotation and does not parse it as features)
                if cur:
                    feats.append(_finish_feature(*cur))
                    cur = None
                section = "contig"
                continue
            if line and not line.startswith(" "):
                continue
            key = line[5:21].strip()
            body = line[21:].strip()
            if key:
                if cur:
                    feats.append(_finish_feature(*cur))
                cur = [key, body, []]
            elif cur is not None:
                if not cur[2] and not body.startswith("/"):
                    cur[1] += body  # location continuation
                else:
                    cur[2].append(body)
            continue
        if section == "contig":
            if line.startswith("ORIGIN"):
                section = "origin"
            continue
        if section == "origin":
            seq_chunks.append("".join(line.split()[1:]))


def genbank_to_dict(path):
    """Equivalent of ``SeqIO.to_dict(SeqIO.parse(handle, "genbank"))`` (GenBankParser.py:17-18)."""
    out = {}
    for r in read_genbank(path):
        if r.id in out:
            raise ValueError(f"Duplicate key '{r.id}'")
        out[r.id] = r
    return out


def _format_location(loc):
    def one(p):
        return f"{p.start + 1}..{p.end}"

    if isinstance(loc, CompoundLocation):
        strand = loc.strand
        parts = loc.parts[::-1] if strand == -1 else loc.parts
        body = f"{loc.operator}({','.join(one(p) for p in parts)})"
    else:
        strand = loc.strand
        body = one(loc)
    return f"complement({body})" if strand == -1 else body


def write_genbank(records, path):
    """Minimal GenBank writer: enough for read_genbank (and Biopython) to round-trip
    id, topology, organism, description, source/gene features and sequence."""
    with open(path, "w") as h:
        for r in records:
            acc = r.id.split(".")[0]
            topo = r.annotations.get("topology", "linear") or "linear"
            h.write(f"LOCUS       {acc:<16} {len(r.seq):>11} bp    DNA     {topo:<8} BCT 01-JAN-2000\n")
            h.write(f"DEFINITION  {r.description or acc}.\n")
            h.write(f"ACCESSION   {acc}\n")
            h.write(f"VERSION     {r.id}\n")
            org = r.annotations.get("organism")
            if org:
                h.write(f"SOURCE      {org}\n  ORGANISM  {org}\n")
            h.write("FEATURES             Location/Qualifiers\n")
            for f in r.features:
                h.write(f"     {f.type:<16}{_format_location(f.location)}\n")
                for k, vals in f.qualifiers.items():
                    for v in vals:
                        h.write(f'                     /{k}="{v}"\n')
            h.write("ORIGIN\n")
            s = str(r.seq).lower()
            for i in range(0, len(s), 60):
                chunk = s[i:i + 60]
                h.write(f"{i + 1:>9} " + " ".join(chunk[j:j + 10] for j in range(0, len(chunk), 10)) + "\n")
            h.write("//\n")


# --------------------------------------------------------------------------- SnapGene

def _strip_html(s: str) -> str:
    s = unescape(s)
    s = re.sub(r"<br\s*/?>", " ", s)
    return re.sub(r"<[^>]+>", "", s)


def read_snapgene(path, record_id=None):
    """SnapGene .dna -> SeqRecord.  Packets are ``[type u8][len u32 BE][data]``;
    type 0 = flags byte (bit0 circular) + sequence, type 10 = features XML
    (SURVEY.md section 8c "Plasmid inputs")."""
    import xml.etree.ElementTree as ET

    data = open(path, "rb").read()
    i, seq, topo, feats, notes = 0, "", "linear", [], {}
    xml_features = None
    while i + 5 <= len(data):
        t = data[i]
        (ln,) = struct.unpack(">I", data[i + 1:i + 5])
        body = data[i + 5:i + 5 + ln]
        i += 5 + ln
        if t == 9 and not body.startswith(b"SnapGene"):
            raise ValueError("not a SnapGene file")
        elif t == 0:
            topo = "circular" if body[0] & 1 else "linear"
            seq = body[1:].decode("ascii")
        elif t == 10:
            xml_features = body.decode("utf-8")
        elif t == 6:
            txt = body.decode("utf-8", "replace")
            for tag in ("AccessionNumber", "Organism", "Description"):
                m = re.search(rf"<{tag}>(.*?)</{tag}>", txt, re.S)
                if m:
                    notes[tag] = _strip_html(m.group(1)).strip()
    n = len(seq)
    if xml_features:
        root = ET.fromstring(xml_features)
        for fe in root.iter("Feature"):
            ftype = fe.get("type")
            direc = fe.get("directionality")
            strand = {"1": 1, "2": -1}.get(direc, 1 if ftype == "source" else None)
            parts = []
            for seg in fe.iter("Segment"):
                a, b = (int(x) for x in seg.get("range").split("-"))
                if a <= b:
                    parts.append(SimpleLocation(a - 1, b, strand))
                else:  # origin-spanning segment on a circular molecule: a..n + 1..b
                    parts.append(SimpleLocation(a - 1, n, strand))
                    parts.append(SimpleLocation(0, b, strand))
            if not parts:
                continue
            if strand == -1:
                parts = parts[::-1]
            loc = parts[0] if len(parts) == 1 else CompoundLocation(parts)
            quals = {}
            for q in fe.findall("Q"):
                vals = []
                for v in q.findall("V"):
                    if v.get("text") is not None:
                        vals.append(_strip_html(v.get("text")))
                    elif v.get("predef") is not None:
                        vals.append(v.get("predef"))
                    elif v.get("int") is not None:
                        vals.append(v.get("int"))
                if vals:
                    quals[q.get("name")] = vals
            feats.append(SeqFeature(loc, ftype, quals))
    acc = notes.get("AccessionNumber") or record_id or "unknown"
    rid = record_id or acc
    rec = SeqRecord(Seq(seq.upper()), id=rid, name=acc,
                    description=notes.get("Description", "").rstrip("."))
    rec.annotations["topology"] = topo
    if "Organism" in notes:
        rec.annotations["organism"] = notes["Organism"]
    rec.features = feats
    return rec


# --------------------------------------------------------------------------- FASTA

def read_fasta(path):
    import gzip
    opener = gzip.open if str(path).endswith(".gz") else open
    rid, desc, chunks = None, "", []
    with opener(path, "rt") as h:
        for line in h:
            line = line.rstrip("\n")
            if line.startswith(">"):
                if rid is not None:
                    yield SeqRecord(Seq("".join(chunks)), id=rid, name=rid, description=desc)
                desc = line[1:]
                rid = desc.split()[0] if desc.split() else ""
                chunks = []
            elif line:
                chunks.append(line.strip())
    if rid is not None:
        yield SeqRecord(Seq("".join(chunks)), id=rid, name=rid, description=desc)


def write_fasta(records, handle, width=60):
    n = 0
    for r in records:
        desc = r.description if r.description and r.description != "<unknown description>" else ""
        title = r.id if not desc or desc.startswith(r.id) and False else f"{r.id} {desc}".strip()
        handle.write(f">{title}\n")
        s = str(r.seq)
        for i in range(0, len(s), width):
            handle.write(s[i:i + width] + "\n")
        n += 1
    return n
