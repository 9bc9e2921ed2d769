#!/usr/bin/env python
"""A/B runner for variant builds of the library (tools/build_variant.sh): the workload is generated once,
every variant runs in its own process against the same arrays.

    python tools/ab_bench.py [--config cfg4] [--steps 3] name[:ENV=VAL[,ENV=VAL]] ...

name = 'cur' (barcoder_b200/libbarcoder_b200.so) or X for bench_kernels/var_X.so; environment overrides after ':'.
Prints one line per variant: step ms, per-stage ms, hits.  Measurement helper, not product code."""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(args):
    import numpy as np
    import torch
    import bench
    from barcoder_b200 import _native
    cfg = bench.CONFIGS[args.config]
    genome = np.load(args.data + "_g.npy")
    off = np.load(args.data + "_o.npy")
    lib = np.load(args.data + "_l.npy")
    n, L = lib.shape
    dev = torch.device("cuda", 0)
    d_genome = torch.from_numpy(genome).to(dev)
    d_lib = torch.from_numpy(lib.reshape(-1)).to(dev)
    s = _native.Searcher(0)
    s.set_pam(cfg["pam"], "downstream", iupac=cfg["iupac"])
    for key, val in (("path", args.path), ("key_nt", args.key_nt)):
        if val:
            s.set_param({"path": _native.BC_PARAM_PATH, "key_nt": _native.BC_PARAM_KEY_NT}[key], val)
    s.set_genome_device(d_genome.data_ptr(), off)
    s.set_library_device(d_lib.data_ptr(), n, L)
    k = cfg["k"]
    for _ in range(2):
        s.build_index(k)
        s.search(k)
    torch.cuda.synchronize()
    keys = ["ms_build_index", "ms_search", "ms_scan_kernel", "ms_genome_bucket", "ms_win_count", "ms_win_bin", "ms_win_place",
            "ms_finish"]
    acc = {key: 0.0 for key in keys}
    t0 = time.time()
    nh = 0
    for _ in range(args.steps):
        s.build_index(k)
        nh = s.search(k)
        st = s.stats()
        for key in keys:
            acc[key] += st[key]
    wall = (time.time() - t0) / args.steps * 1e3
    out = {key[3:]: round(v / args.steps, 3) for key, v in acc.items()}
    out["verify"] = round(out["scan_kernel"] - out["finish"], 3)
    print(json.dumps({"variant": args.child, "step_ms": round(out["build_index"] + out["search"], 3), "wall_ms": round(wall, 2),
                      "hits": int(nh), "combos": st["combos"], "key_nt": st["key_nt"], "path": st["path"], **out}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("variants", nargs="*")
    ap.add_argument("--config", default="cfg4")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--path", type=int, default=0)
    ap.add_argument("--key-nt", type=int, default=0)
    ap.add_argument("--child", default=None)
    ap.add_argument("--data", default="/dev/shm/bc_ab")
    args = ap.parse_args()
    if args.child:
        child(args)
        return
    import numpy as np
    import bench
    genome, off, lib = bench.make_workload(bench.CONFIGS[args.config], 0, args.scale)
    np.save(args.data + "_g.npy", genome)
    np.save(args.data + "_o.npy", off)
    np.save(args.data + "_l.npy", lib)
    for spec in args.variants:
        name, _, envs = spec.partition(":")
        env = dict(os.environ)
        if name != "cur":
            env["BARCODER_B200_LIB"] = os.path.join(ROOT, "bench_kernels", f"var_{name}.so")
        for kv in filter(None, envs.split(",")):
            key, _, val = kv.partition("=")
            env[key] = val
        cmd = [sys.executable, os.path.abspath(__file__), "--child", spec, "--config", args.config, "--steps", str(args.steps),
               "--data", args.data, "--path", str(args.path), "--key-nt", str(args.key_nt)]
        r = subprocess.run(cmd, env=env, capture_output=True, text=True)
        sys.stdout.write(r.stdout if r.returncode == 0 else f"{spec}: FAILED rc={r.returncode}\n{r.stderr[-2000:]}\n")
        sys.stdout.flush()
    for suffix in ("_g.npy", "_o.npy", "_l.npy"):
        try:
            os.remove(args.data + suffix)
        except OSError:
            pass


if __name__ == "__main__":
    main()
