#!/usr/bin/env python
"""profiles/traffic.json from an `ncu --set full` capture of one bench.py step.

    python tools/make_traffic.py gpurun_out/prof.ncu-rep cfg4:key10:c15:path3:n1
    python tools/make_traffic.py profiles/r2/full_cfg4_raw.csv cfg4:key10:c15:path3:n1   # the raw-page CSV export works too

Per kernel of the compact join: dram__bytes_read.sum + dram__bytes_write.sum of ONE launch (the
first captured), keyed the way bench.py looks it up.  The file is stamped with a hash of the kernel
sources (bench.source_hash): bench.py ignores it when the sources have changed since the capture."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

NAMES = {"k_cverify": "k_cverify", "k_cfinish": "k_cfinish", "k_cbin<false>": "k_cbin", "k_cbin<(bool)0>": "k_cbin", "k_cbin<0>": "k_cbin",
         "k_cplace_bulk<false": "k_cplace", "k_cplace_bulk<(bool)0": "k_cplace", "k_cplace_bulk<0": "k_cplace",
         "k_scan_probe": "verify"}
# bench.py's "k_ccount" stage = bin count + slot count of the window side (the slot count kernel also runs once for the index)
COUNT = ("k_ccount<false>", "k_ccount<(bool)0>", "k_ccount<0>", "k_cbincount<false>", "k_cbincount<(bool)0>", "k_cbincount<0>")
INDEX = ("k_ccount<true>", "k_ccount<(bool)1>", "k_ccount<1>", "k_cbin<true>", "k_cbin<(bool)1>", "k_cbin<1>", "k_cplace_bulk<true",
         "k_cplace_bulk<(bool)1", "k_cplace_bulk<1", "k_cbincount<true>", "k_cbincount<(bool)1>", "k_cbincount<1>")


def unit_scale(u):
    return {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(u, 1.0)


def main():
    rep, key = sys.argv[1], sys.argv[2]
    if rep.endswith(".csv"):
        raw = open(rep).read()
    else:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ir, iw, ik = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("Kernel Name")
    out, index_bytes, count_bytes, slotcounts = {}, 0.0, 0.0, []
    for r in rows[2:]:
        name = r[ik]
        b = float(r[ir]) * unit_scale(units[ir]) + float(r[iw]) * unit_scale(units[iw])
        if any(name.startswith("void " + p) or name.startswith(p) for p in INDEX):
            index_bytes += b
            continue
        if any(name.startswith("void " + p) or name.startswith(p) for p in COUNT):
            count_bytes += b
            continue
        if name.startswith("k_cslotcount"):
            slotcounts.append(b)  # first launch = index build, second = window side
            continue
        for pat, short in NAMES.items():
            if (name.startswith("void " + pat) or name.startswith(pat)) and short not in out:
                out[short] = int(b)
    if slotcounts:
        if len(slotcounts) > 1:
            index_bytes += slotcounts[0]
        count_bytes += slotcounts[-1]
    if count_bytes:
        out["k_ccount"] = int(count_bytes)
    if index_bytes:
        out["index_build"] = int(index_bytes)
    path = os.path.join(ROOT, "profiles", "traffic.json")
    doc = {"_comment": "DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) from ncu --set full captures of "
                       "`python bench.py` steps; source_hash = bench.source_hash() of the kernel sources they were taken from",
           "source_hash": bench.source_hash(), "entries": {}}
    if os.path.exists(path):
        try:
            old = json.load(open(path))
            if old.get("source_hash") == doc["source_hash"]:
                doc["entries"] = old.get("entries", {})
        except ValueError:
            pass
    doc["entries"][key] = out
    doc["entries"][key]["_report"] = os.path.basename(rep)
    with open(path, "w") as h:
        json.dump(doc, h, indent=1)
    print(json.dumps(doc["entries"][key], indent=1))


if __name__ == "__main__":
    main()
