#!/usr/bin/env python
"""Markdown tables for profiles/r2_summary.md from the evidence files in profiles/r2/ (bench JSON lines, the ncu launch
list and the raw-page CSV export of the full-set capture).

    python tools/r2_report.py [suffix]        # suffix of the evidence set, default "final"
"""
import collections
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
D = os.path.join(ROOT, "profiles", "r2")


def jline(name):
    path = os.path.join(D, name)
    if not os.path.exists(path):
        return None
    for line in open(path):
        if line.startswith("{"):
            return json.loads(line)
    return None


def launches(name):
    path = os.path.join(D, name)
    if not os.path.exists(path):
        return None
    rows = list(csv.reader(open(path)))
    h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hd = rows[h]
    kn, mv = hd.index("Kernel Name"), hd.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[h + 1:]:
        if len(r) <= mv:
            continue
        agg.setdefault(r[kn].split("(")[0].replace("void ", ""), []).append(float(r[mv].replace(",", "")) / 1e6)
    return agg


def raw(name):
    path = os.path.join(D, name)
    if not os.path.exists(path):
        return None
    rows = list(csv.reader(open(path)))
    hdr = rows[0]
    idx = {h: i for i, h in enumerate(hdr)}
    return [{h: r[i] for h, i in idx.items()} for r in rows[2:]]


def f(x, nd=2):
    try:
        return f"{float(x):.{nd}f}"
    except (TypeError, ValueError):
        return "-"


def main():
    sfx = sys.argv[1] if len(sys.argv) > 1 else "final"
    b4 = jline(f"bench_cfg4_{sfx}.json")
    if b4:
        print(f"cfg4 1 GPU: {b4['ms_per_step']:.2f} ms/step, value {b4['value']:.3e}, e2e {b4['e2e']['ms_per_step']:.2f} ms "
              f"({b4['e2e']['value']:.3e}), hits {b4['config']['hits_per_step']}, launches {b4['gpu_launches']}")
        r = b4["roofline"]
        print(f"roofline: {r['bound']} {r['kernel']} achieved {r['achieved']:.3f} / peak {r['peak']:.3f} {r['unit']} = {r['frac']:.3f}; "
              f"share {r['share_of_step']:.2f}; pairs_vs_k+1 {r.get('pairs_vs_k_plus_1')}")
        print("stage_ms:", r["stage_ms"])
        print("e2e phases:", b4["e2e"]["host_phase_ms"])
        if b4.get("cpu_baseline"):
            c = b4["cpu_baseline"]
            print(f"cpu baseline: {c['value']:.3e} {c['unit']} on {c['cores']} cores ({c['kind']}); ratio value/cpu = {b4['value'] / c['value']:.0f}, "
                  f"e2e/cpu = {b4['e2e']['value'] / c['value']:.0f}")
    for n in (2, 4, 8):
        bn = jline(f"bench_cfg4_n{n}_{sfx}.json")
        if bn and b4:
            print(f"cfg4 {n} GPUs: {bn['ms_per_step']:.2f} ms/step, value {bn['value']:.3e}, speed-up {b4['ms_per_step'] / bn['ms_per_step']:.2f} "
                  f"(efficiency {b4['ms_per_step'] / bn['ms_per_step'] / n:.3f}), merged_ok {bn['merged_ok']}, e2e {bn['e2e']['ms_per_step']:.2f} ms")
    b5 = jline(f"bench_cfg5_{sfx}.json")
    if b5:
        r = b5["roofline"]
        print(f"cfg5 1 GPU: {b5['ms_per_step']:.2f} ms/step, value {b5['value']:.3e}; roofline {r['bound']} {r['achieved']:.1f}/{r['peak']:.1f} {r['unit']} = {r['frac']:.3f}")
    bc = jline(f"bench_class_{sfx}.json")
    if bc:
        for k, v in bc["results"].items():
            print(f"class API {k}: total {v['phase_s']['total']:.2f} s, align {v['phase_s']['align']:.3f} s (device {v['phase_s']['align_device_ms']:.1f} ms), "
                  f"{v['guides']} guides, {v['alignments']} alignments")
    la = launches(f"launches_cfg4_{sfx}.csv")
    if la:
        print("\n| kernel | launches | ms per launch (last step) |\n|---|---:|---:|")
        tot = 0.0
        for k, a in la.items():
            if a[-1] >= 0.02 and not k.startswith(("k_popc", "k_verify_atom", "k_gather", "k_pack")):
                print(f"| {k} | {len(a)} | {a[-1]:.3f} |")
                tot += a[-1]
        print(f"| sum | | {tot:.2f} |")
    for name in (f"full_cfg4_raw_{sfx}.csv", f"full_cfg5_raw_{sfx}.csv"):
        rw = raw(name)
        if not rw:
            continue
        print(f"\n{name}\n| kernel | ms | DRAM r GB | DRAM w GB | DRAM % | L2 % | l1tex % | issue % | ALU % | XU % | warps % | regs | long sb | short sb |\n"
              "|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
        for r in rw:
            print("| " + " | ".join([
                r["Kernel Name"].split("(")[0].replace("void ", ""), f(r.get("gpu__time_duration.sum"), 3),
                f(r.get("dram__bytes_read.sum")), f(r.get("dram__bytes_write.sum")),
                f(r.get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), 1),
                f(r.get("lts__throughput.avg.pct_of_peak_sustained_elapsed"), 1),
                f(r.get("l1tex__throughput.avg.pct_of_peak_sustained_elapsed"), 1),
                f(r.get("smsp__issue_active.avg.pct_of_peak_sustained_active"), 1),
                f(r.get("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"), 1),
                f(r.get("sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active"), 1),
                f(r.get("sm__warps_active.avg.pct_of_peak_sustained_active"), 1),
                r.get("launch__registers_per_thread", "-"),
                f(r.get("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio")),
                f(r.get("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"))]) + " |")


if __name__ == "__main__":
    main()
