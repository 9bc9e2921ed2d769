#!/usr/bin/env python
"""Offline generator of barcoder_b200/csrc/bc_designs.inc (seed covering designs).

For a spacer length L and a mismatch budget k, a *seed design* is a family of position masks
(each the key of one seed combination) such that every set of k mismatch positions is avoided by
at least one mask.  The pigeonhole scheme "b blocks, every (b-k)-subset" is one such family; for
k >= 2 smaller families with longer keys exist (covering designs, e.g. L=20, k=3: 15 masks of 10 nt
against the 20 masks of 9..11 nt of b=6).  tools/design_search.c finds them by simulated annealing;
this script drives it over a grid of (L, k, key length), keeps the smallest family found for each
cell and writes the table the library compiles in.  Every row is re-verified here (exhaustively)
and again by tests/test_designs.py.

    gcc -O2 -o /tmp/design_search tools/design_search.c -lm
    python tools/make_designs.py [--jobs 8] [--budget 40]
"""
import argparse
import itertools
import json
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOOL = "/tmp/design_search"
MAX_RUNS = 6
MAX_MASKS = 40


def covers(L, k, masks):
    for T in itertools.combinations(range(L), k):
        t = sum(1 << p for p in T)
        if not any(m & t == 0 for m in masks):
            return False
    return True


def runs(m):
    return bin(m & ~(m << 1)).count("1")


def block_scheme(L, k, b, cap=12):
    bs = [j * L // b for j in range(b + 1)]
    out = []
    for sub in itertools.combinations(range(b), b - k):
        budget, m = cap, 0
        for j in sub:
            ln = min(bs[j + 1] - bs[j], budget)
            m |= ((1 << ln) - 1) << bs[j]
            budget -= ln
            if not budget:
                break
        out.append(m)
    return out


def search(L, k, s, C, seed, iters, timeout):
    try:
        r = subprocess.run([TOOL, str(L), str(k), str(s), str(C), str(MAX_RUNS), str(seed), str(iters)],
                           capture_output=True, text=True, timeout=timeout)
    except subprocess.TimeoutExpired:
        return None
    if r.returncode != 0:
        return None
    f = r.stdout.split()
    masks = [int(x, 16) for x in f[4:]]
    assert len(masks) == C and all(bin(m).count("1") == s and runs(m) <= MAX_RUNS for m in masks)
    assert covers(L, k, masks)
    return masks


def best_for(L, k, s, budget):
    """Smallest family found for (L, k, s): walk C downwards from a feasible start."""
    # feasible start: the block scheme with the smallest b whose shortest key is >= s
    start = MAX_MASKS
    iters = 12_000_000 if L <= 24 else 5_000_000
    best = None
    C = start
    fails = 0
    while C >= 2:
        got = None
        for seed in (1, 2, 3):
            got = search(L, k, s, C, seed, iters, budget)
            if got:
                break
        if not got:
            fails += 1
            if best is None and C == start:
                return None
            break
        best = got
        C -= 1 if C <= 24 else 2
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--jobs", type=int, default=os.cpu_count() or 4)
    ap.add_argument("--budget", type=float, default=40.0, help="seconds per annealing run")
    ap.add_argument("--json", default=os.path.join(ROOT, "tools", "designs.json"))
    ap.add_argument("--render-only", action="store_true")
    args = ap.parse_args()
    cells = []
    for L in range(14, 33):
        for s in (8, 9, 10, 11, 12):
            if s < L - 3:
                cells.append((L, 3, s))
        for s in (10, 11, 12):
            if s < L - 2:
                cells.append((L, 2, s))
    table = {}
    if os.path.exists(args.json):
        with open(args.json) as h:
            table = {tuple(map(int, key.split(","))): v for key, v in json.load(h).items()}
    if not args.render_only:
        todo = [c for c in cells if c not in table]
        with ThreadPoolExecutor(args.jobs) as ex:
            for cell, masks in zip(todo, ex.map(lambda c: best_for(*c, args.budget), todo)):
                if masks:
                    table[cell] = masks
                    print(cell, len(masks), file=sys.stderr, flush=True)
                    with open(args.json, "w") as h:
                        json.dump({",".join(map(str, key)): v for key, v in sorted(table.items())}, h, indent=0)
    render(table)


def render(table):
    """Keep a design only if no block scheme has as few masks with keys at least as long."""
    rows = []
    for (L, k, s), masks in sorted(table.items()):
        assert covers(L, k, masks)
        dominated = False
        for b in range(k + 1, min(k + 5, 9, L + 1)):
            bm = block_scheme(L, k, b)
            if len(bm) <= len(masks) and min(bin(m).count("1") for m in bm) >= s:
                dominated = True
        if not dominated:
            rows.append((L, k, s, masks))
    out = ["// bc_designs.inc - generated by tools/make_designs.py (simulated annealing, tools/design_search.c);",
           "// do not edit.  One row per seed covering design: every set of k positions of an L-mer is avoided",
           "// by at least one mask; all masks of a row have key_nt bits set in at most 6 runs.",
           "// tests/test_designs.py re-verifies every row exhaustively.",
           "// {L, k, key_nt, n_masks, first index into bc_design_masks}"]
    flat = []
    out.append("static const BcDesignRow bc_design_rows[] = {")
    for L, k, s, masks in rows:
        out.append(f"    {{{L}, {k}, {s}, {len(masks)}, {len(flat)}}},")
        flat += masks
    out.append("};")
    out.append("static const uint32_t bc_design_masks[] = {")
    for i in range(0, len(flat), 8):
        out.append("    " + ", ".join(f"0x{m:08x}u" for m in flat[i:i + 8]) + ",")
    out.append("};")
    path = os.path.join(ROOT, "barcoder_b200", "csrc", "bc_designs.inc")
    with open(path, "w") as h:
        h.write("\n".join(out) + "\n")
    print(f"{len(rows)} designs, {len(flat)} masks -> {path}", file=sys.stderr)


if __name__ == "__main__":
    main()
