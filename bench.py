#!/usr/bin/env python
"""bench.py - guides*Mbp/s of the spacer->genome mismatch search on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # the CPU path timed on host cores

Workload (config.workload): BASELINE.json configs[3] - ONE 10^7-spacer 20-mer library (1 % planted)
against a 100 Mbp synthetic genome, <= 3 mismatches, PAM NGG downstream.  At N > 1 the job is the
same library (STRONG scaling): every GPU holds the packed genome and library, and the seed
directory is cut into N slot ranges (BC_PARAM_SLOT_PART) - a GPU indexes, sorts and verifies only
the (window, entry) pairs whose seed key falls in its range, so index build, window sort and
verification all shrink with N; the hit records are merged on rank 0 through a peer-memory sink
(copy engines over NVLink while the search runs).  `--shard library` runs the round-1 weak-scaling
job instead (10^7 spacers PER GPU), `--shard genome` cuts the genome into ranges (cfg 5).

A step = seed-index build + genome scan (window sort + verification, PAM fused) + hit merge, with
the ASCII inputs already resident in HBM (`value`).  `e2e` is the same metric through the host
C-ABI calls: pinned host ASCII buffers -> H2D -> pack -> index -> scan -> D2H of the hit records
(streamed into a pinned buffer through bc_set_hit_sink while the scan runs).
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # name: genome bp, contigs, N fraction, genome seed, spacers per GPU, L, library seed, planted, k, pam, iupac
    "cfg4": dict(G=100_000_000, contigs=1, nfrac=0.0, gseed=4, n=10_000_000, L=20, lseed=40, planted=0.01,
                 k=3, pam="NGG", iupac=False,
                 name="cfg4: 10M 20-mers x 100 Mbp synthetic, k<=3, NGG downstream (BASELINE.json configs[3])"),
    "cfg3": dict(G=4_641_652, contigs=1, nfrac=0.0, gseed=1, n=None, L=20, lseed=0, planted=0.0, k=3, pam="NGG",
                 iupac=False, name="cfg3: all NGG 20-mers of a 4.64 Mbp synthetic genome vs itself, k<=3"),
    "cfg5": dict(G=3_000_000_000, contigs=24, nfrac=0.001, gseed=5, n=1_000_000, L=32, lseed=50, planted=0.01,
                 k=2, pam="NNGRRT", iupac=True,
                 name="cfg5: 1M 32-mers x 3 Gbp synthetic (24 contigs, 0.1% N), k<=2, NNGRRT downstream"),
    "tiny": dict(G=2_000_000, contigs=3, nfrac=0.001, gseed=7, n=200_000, L=20, lseed=70, planted=0.01, k=3,
                 pam="NGG", iupac=False, name="tiny: 200k 20-mers x 2 Mbp (debug)"),
}


def make_workload(cfg, rank, scale=1.0):
    from barcoder_b200 import synth
    G = max(1000, int(cfg["G"] * scale))
    genome, off = synth.random_genome(G, seed=cfg["gseed"], n_contigs=cfg["contigs"], n_fraction=cfg["nfrac"])
    if cfg["n"] is None:
        lib = synth.enumerate_pam_guides(genome, off, cfg["L"], cfg["pam"])
    else:
        n = max(100, int(cfg["n"] * scale))
        lib = synth.random_library(n, cfg["L"], seed=cfg["lseed"] + rank)
        synth.plant(lib, genome, cfg["planted"], cfg["k"], seed=cfg["lseed"] + 1000 + rank)
    return genome, off, np.ascontiguousarray(lib)


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons for one GPU during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.rows:
            if ts < t0 - 0.05 or ts > t1 + 0.15:
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0]))
                smax = float(f[1])
            except (ValueError, IndexError):
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax,
                "samples": len(sm), "reasons": sorted(reasons)}


def int_peak(device):
    import ctypes
    so = os.path.join(ROOT, "bench_kernels", "libbc_ubench.so")
    if not os.path.exists(so):
        return None
    lib = ctypes.CDLL(so)
    popc, atom, sms = ctypes.c_double(), ctypes.c_double(), ctypes.c_int()
    rc = lib.ub_int_peak(int(device), ctypes.byref(popc), ctypes.byref(atom), ctypes.byref(sms))
    if rc != 0:
        return None
    return {"popc_per_s": popc.value, "verify_atom_per_s": atom.value, "sm_count": sms.value}


def gather_peak(device, table_mb):
    """Divergent 4-byte loads per second an SM array can retire from an L2-resident table (bench_kernels/int_peak.cu):
    the bound of the probe kernel, whose directory probes and bucket reads touch one 128-byte line per lane."""
    import ctypes
    so = os.path.join(ROOT, "bench_kernels", "libbc_ubench.so")
    if not os.path.exists(so):
        return None
    lib = ctypes.CDLL(so)
    if not hasattr(lib, "ub_gather_peak"):
        return None
    rate = ctypes.c_double()
    if lib.ub_gather_peak(int(device), int(table_mb), ctypes.byref(rate)) != 0:
        return None
    return rate.value


def bind_to_gpu_numa_node(index):
    """Multi-GPU runs: pin this rank's host threads to the CPU cores next to its GPU (NVML CPU affinity) BEFORE the
    pinned host buffers are allocated, so that first touch places them on that NUMA node.  Round 1's e2e arm lost 25 %
    at 8 GPUs with all eight ranks streaming their hit records into pinned memory of one node."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return f"{len(cpus)} cores near GPU {index}"
    except Exception as exc:  # no NVML, no permission: run unbound
        return f"unbound ({type(exc).__name__})"
    return "unbound"


def source_hash():
    """Hash of the kernel sources: profiles/traffic.json is only trusted for the code it was captured from."""
    import glob
    import hashlib
    h = hashlib.sha256()
    for path in sorted(glob.glob(os.path.join(ROOT, "barcoder_b200", "csrc", "*.cu*")) +
                       glob.glob(os.path.join(ROOT, "barcoder_b200", "csrc", "*.inc")) +
                       glob.glob(os.path.join(ROOT, "barcoder_b200", "csrc", "*.h"))):
        with open(path, "rb") as f:
            h.update(f.read())
    return h.hexdigest()[:16]


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as h:
            return json.load(h), "measured (MEASURED_PEAKS.json)"
    except OSError:
        return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


def cpu_sample_once(cfg, genome, lib, n_s, G_s, threads):
    from oracle import oracle
    contigs = [bytes(genome[:G_s])]
    t0 = time.time()
    hits = oracle.search(contigs, np.ascontiguousarray(lib[:n_s]), cfg["k"], pam=cfg["pam"], direction="downstream",
                         flags=oracle.PAM_FLAG_IUPAC if cfg["iupac"] else 0, threads=threads)
    dt = time.time() - t0
    return dict(value=n_s * (G_s / 1e6) / dt, seconds=dt, n=n_s, G=G_s, hits=int(len(hits)))


def cpu_calibrate(cfg, genome, lib, budget_s, threads):
    """Pick a bounded sample (first n_s spacers x first G_s bases) that keeps the oracle busy for
    roughly budget_s seconds on this box.  The library sample is large (up to 2*10^6 spacers) so
    that the seed buckets are as dense as in the full job and the oracle's cost model picks the
    same kind of scheme; the rate includes its index build, like the GPU step does."""
    n_s = min(len(lib), 2_000_000)
    G_s = min(len(genome), 1_000_000)
    best = None
    for _ in range(5):
        best = cpu_sample_once(cfg, genome, lib, n_s, G_s, threads)
        if best["seconds"] >= budget_s / 2 or G_s >= len(genome):
            break
        G_s = min(len(genome), int(G_s * min(8.0, max(2.0, budget_s / max(best["seconds"], 1e-3)))))
    return best


def cpu_result(cfg, best, threads, n_full, G_full):
    frac = best["n"] * best["G"] / (float(n_full) * G_full)
    return {"value": best["value"], "unit": "guides*Mbp/s", "cores": threads, "kind": "port",
            "sample_fraction": frac, "extrapolated_full_job_s": n_full * (G_full / 1e6) / best["value"],
            "bowtie": bowtie_probe(),
            "sample": f"first {best['n']} spacers x first {best['G']} bp of the genome ({frac:.2e} of the job, "
                      f"rate extrapolated linearly), k={cfg['k']}, {best['seconds']:.1f} s incl. index build, "
                      f"{best['hits']} hits; oracle/oracle.c: generalised-pigeonhole seeds chosen by a cost model, "
                      f"key-sorted 2-bit library, rolling window, popcount-first verification, {threads} threads "
                      f"(CPU restatement, not bowtie - bowtie 1.3.1 is not installable here)"}


def bowtie_probe():
    """The reference shells out to bowtie / bowtie-build (BowtieRunner.py:87,107).  If the binaries
    ever appear on PATH the exact reference command line can be timed; here it only records that
    they are absent."""
    import shutil
    found = {name: shutil.which(name) for name in ("bowtie", "bowtie-build")}
    return {"available": all(found.values()), "paths": found}


def bowtie_time(cfg, genome, lib, n_s, G_s, threads):
    """bowtie-build + `bowtie -S -a --nomaqround --mm --tryhard --quiet --best -p N -v k`
    (BowtieRunner.py:111-125) on the same bounded sample, if the binaries exist."""
    import shutil
    import tempfile
    if not (shutil.which("bowtie") and shutil.which("bowtie-build")):
        return None
    with tempfile.TemporaryDirectory() as d:
        fa, fq, idx, sam = (os.path.join(d, x) for x in ("g.fasta", "r.fastq", "idx", "out.sam"))
        with open(fa, "w") as h:
            h.write(">chr\n" + bytes(genome[:G_s]).decode() + "\n")
        with open(fq, "w") as h:
            for i, row in enumerate(lib[:n_s]):
                sp = bytes(row).decode()
                h.write(f"@r{i}\n{sp}\n+\n{'I' * len(sp)}\n")
        t0 = time.time()
        subprocess.run(["bowtie-build", fa, idx], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        t1 = time.time()
        subprocess.run(["bowtie", "-S", "-a", "--nomaqround", "--mm", "--tryhard", "--quiet", "--best", "-p",
                        str(threads), "-v", str(cfg["k"]), "-x", idx, fq, sam], check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        t2 = time.time()
    return {"build_s": t1 - t0, "align_s": t2 - t1, "threads": threads, "n": n_s, "G": G_s,
            "value": n_s * (G_s / 1e6) / (t2 - t0)}


def cpu_baseline_run(cfg, genome, off, lib, budget_s=20.0, threads=None):
    """Time the oracle (CPU restatement, kind 'port') on a bounded sample of the workload."""
    threads = threads or os.cpu_count() or 1
    best = cpu_calibrate(cfg, genome, lib, budget_s, threads)
    res = cpu_result(cfg, best, threads, len(lib), len(genome))
    bt = bowtie_time(cfg, genome, lib, min(best["n"], 200_000), best["G"], threads)
    if bt:
        res["bowtie_timed"] = bt
    return res


def run_reference(args, cfg):
    """--impl reference: the reference's CPU implementation of the path on the host cores.  The
    reference shells out to bowtie 1.3.1, which cannot be installed offline (and the repo is not
    pip-installable); the oracle port of the same search stands in (see oracle/oracle.c header).
    Every step searches the same bounded sample of the workload with all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    genome, off, lib = make_workload(cfg, 0, args.scale)
    threads = os.cpu_count() or 1
    cal = cpu_calibrate(cfg, genome, lib, 12.0, threads)
    vals, last = [], cal
    for i in range(args.warmup + args.steps):
        last = cpu_sample_once(cfg, genome, lib, cal["n"], cal["G"], threads)
        if i >= args.warmup:
            vals.append(last["value"])
    value = statistics.mean(vals) if vals else cal["value"]
    n_tot = len(lib)
    res = cpu_result(cfg, last, threads, len(lib), len(genome))
    line = {
        "impl": "reference", "metric": "guides*Mbp/s at <=k mismatches", "value": value, "unit": "guides*Mbp/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * last["seconds"], "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "u64 2-bit packed (XOR/popcount)", "data": "synthetic",
        "config": {"workload": cfg["name"], "k": cfg["k"], "pam": cfg["pam"], "spacers": len(lib),
                   "genome_bp": len(genome), "note": "each step = the bounded sample in cpu_baseline.sample; "
                   f"full job would take ~{n_tot * (len(genome) / 1e6) / value:.0f} s at this rate"},
        "cpu_baseline": dict(res, value=value),
        "e2e": {"value": value, "unit": "guides*Mbp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def synthetic_genbank(path, G, seed, n_genes):
    """A GenBank file of the cfg-1 shape: one circular contig of G uniform bases with n_genes random
    gene features (the real GCA_000005845.2 file is missing from the reference checkout)."""
    from barcoder_b200 import seqio, synth
    genome, _ = synth.random_genome(G, seed=seed)
    rng = synth.rng_for(seed + 100)
    rec = seqio.SeqRecord(seqio.Seq(bytes(genome).decode()), id="SYN_000001.1", name="SYN_000001", description="synthetic")
    rec.annotations["topology"] = "circular"
    rec.annotations["organism"] = "synthetic"
    feats = [seqio.SeqFeature(seqio.SimpleLocation(0, G, 1), "source", {"organism": ["synthetic"]})]
    starts = np.sort(rng.integers(0, G - 3000, size=n_genes))
    lens = rng.integers(300, 2500, size=n_genes)
    strands = rng.integers(0, 2, size=n_genes) * 2 - 1
    for i, (a, ln, sd) in enumerate(zip(starts.tolist(), lens.tolist(), strands.tolist())):
        feats.append(seqio.SeqFeature(seqio.SimpleLocation(a, a + ln, sd), "gene",
                                      {"locus_tag": [f"SYN_{i:05d}"], "gene": [f"g{i}"]}))
    rec.features = feats
    seqio.write_genbank([rec], path)


def run_class_api(args):
    """--api class: the reference's only timed flow (testing_grounds.py:16-43; design_interactive.ipynb:344-350
    logs ~64 s wall for it on E. coli, ~22 s of which in bowtie-build + bowtie + SAM parsing), through the
    drop-in classes of this repo: GenBank parse -> guide enumeration -> BarCodeLibrary -> PAMFinder ->
    BowtieRunner (make_fasta, make_fastq, create_index, align) -> PySamParser.ranges.join(genbank.ranges) ->
    CRISPRiLibrary.  cfg1: 50,000 of the genome's NGG 20-mers, <= 1 mismatch; cfg3: all of them, <= 3."""
    import tempfile
    from barcoder_b200 import (BarCodeLibrary, BowtieRunner, CRISPRiLibrary, GenBankParser, PAMFinder, PySamParser, _native,
                               synth)
    out = {}
    with tempfile.TemporaryDirectory() as d:
        gb = os.path.join(d, "synthetic.gb")
        synthetic_genbank(gb, 4_641_652, seed=1, n_genes=4300)
        for name, k, n_guides in (("cfg1", 1, 50_000), ("cfg3", 3, None)):
            ph = {}
            t_all = time.time()
            t0 = time.time()
            genbank = GenBankParser(gb)
            ph["parse_genbank"] = time.time() - t0
            t0 = time.time()
            # all distinct ACGT 20-mers 5' of an NGG on either strand (design_guides.py:22-49), on the device
            rec = next(iter(genbank.records.values()))
            with _native.Searcher(0) as srch:
                srch.set_genome([str(rec.seq)])
                rows = srch.enumerate_guides(20, "NGG")
            if n_guides:
                rows = rows[synth.rng_for(1).choice(len(rows), size=n_guides, replace=False)]
            barcodes = BarCodeLibrary()
            barcodes.load_from_list(synth.rows_to_strings(rows))
            ph["enumerate_guides"] = time.time() - t0
            pam = PAMFinder(genbank.records, "NGG", "downstream")
            with BowtieRunner(write_sam=False, write_files=False) as bowtie:
                t0 = time.time(); bowtie.make_fasta(genbank.records); ph["make_fasta"] = time.time() - t0
                t0 = time.time(); bowtie.make_fastq(barcodes.barcodes); ph["make_fastq"] = time.time() - t0
                t0 = time.time(); bowtie.create_index(); ph["create_index"] = time.time() - t0
                t0 = time.time(); bowtie.align(k, 12); ph["align"] = time.time() - t0
                ph["align_device_ms"] = sum(st["ms_search"] + st["ms_build_index"] + st["ms_sort_hits"] for st in bowtie.stats)
                t0 = time.time(); sam = PySamParser(bowtie.sam_path); ranges = sam.ranges; ph["sam_ranges"] = time.time() - t0
                t0 = time.time(); targets = ranges.join(genbank.ranges); ph["join_features"] = time.time() - t0
                n_hits = len(bowtie.hits)
            t0 = time.time()
            lib = CRISPRiLibrary(targets.df, pam)
            ph["crispri_library"] = time.time() - t0
            ph["total"] = time.time() - t_all
            out[name] = {"guides": len(barcodes.barcodes), "k": k, "alignments": int(n_hits),
                         "joined_rows": int(len(targets.df)), "unambiguous_targets": int(len(lib.unambiguous_targets)),
                         "phase_s": {k2: round(v, 4) for k2, v in ph.items()}}
    print(json.dumps({"api": "class", "flow": "testing_grounds.py:16-43", "reference_anecdote_s": {"total": 64, "search": 22,
                      "source": "design_interactive.ipynb:344-350 (E. coli, <=1 mismatch, unknown machine)"},
                      "results": out}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg4", choices=sorted(CONFIGS))
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the workload (debug only)")
    ap.add_argument("--path", type=int, default=0, help="0 auto, 1 probe, 2 join, 3 compact join")
    ap.add_argument("--blocks", type=int, default=0)
    ap.add_argument("--key-nt", type=int, default=0, help="force a seed covering design with keys of this length")
    ap.add_argument("--key-cap", type=int, default=0, help="cap the key length of block schemes")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--shard", default=None, choices=["slots", "library", "genome"],
                    help="multi-GPU partitioning: slot ranges of the seed directory (strong scaling of one "
                         "library, default), genome ranges (strong; default for cfg5, which is probe-bound) or "
                         "library shards (weak scaling: the configured library PER GPU)")
    ap.add_argument("--gate", action="store_true", help="PAM-first gating: report only PAM-adjacent hits")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end arm (kernel experiments only)")
    ap.add_argument("--verify", action="store_true", help="check a sample of the result against the oracle")
    ap.add_argument("--api", default="abi", choices=["abi", "class"],
                    help="class: time the reference's class-API flow (testing_grounds.py) end to end on cfg1 and cfg3")
    args = ap.parse_args()
    if args.api == "class":
        run_class_api(args)
        return
    cfg = CONFIGS[args.config]
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args, cfg)
        return

    import torch
    import torch.distributed as dist
    from barcoder_b200 import _native, multi_gpu

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: barcoder_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    numa_note = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)

    shard = args.shard or ("genome" if args.config == "cfg5" else "slots")
    genome, off, lib = make_workload(cfg, rank if shard == "library" else 0, args.scale)
    n, L = lib.shape
    G = len(genome)
    k = cfg["k"]

    # pinned host copies (e2e arm) and device-resident ASCII inputs (value arm)
    h_genome = torch.from_numpy(genome).pin_memory()
    h_lib = torch.from_numpy(lib.reshape(-1)).pin_memory()
    d_genome = h_genome.to(device, non_blocking=True)
    d_lib = h_lib.to(device, non_blocking=True)
    torch.cuda.synchronize()

    s = _native.Searcher(local_rank)
    s.set_pam(cfg["pam"], "downstream", iupac=cfg["iupac"], gate=args.gate)
    if shard == "library":   # weak scaling: a different library on every rank, global spacer ids
        s.set_param(_native.BC_PARAM_SPACER_ID_BASE, rank * n)
    elif shard == "genome":  # every rank holds the whole library and scans its 1/world slice of the genome
        s.set_param(_native.BC_PARAM_SCAN_PART, rank | (world << 16))
    else:                    # every rank holds genome + library and owns 1/world of the seed directory
        s.set_param(_native.BC_PARAM_SLOT_PART, rank | (world << 16))
    if args.path:
        s.set_param(_native.BC_PARAM_PATH, args.path)
    if args.blocks:
        s.set_param(_native.BC_PARAM_BLOCKS, args.blocks)
    if args.key_nt:
        s.set_param(_native.BC_PARAM_KEY_NT, args.key_nt)
    if args.key_cap:
        s.set_param(_native.BC_PARAM_KEY_CAP, args.key_cap)
    s.set_genome_device(d_genome.data_ptr(), off)
    s.set_library_device(d_lib.data_ptr(), n, L)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    peer = {"g": None}

    def step_resident():
        s.build_index(k)
        nh = s.search(k)
        if world > 1:
            if peer["g"] is None:
                # first (warm-up) step: size the per-rank regions of rank 0's peer buffer from the
                # largest shard; from now on the records reach rank 0 while the search runs
                m = torch.tensor([nh], dtype=torch.int64, device=device)
                dist.all_reduce(m, op=dist.ReduceOp.MAX)
                peer["g"] = multi_gpu.PeerGather(s, device, int(m.item() * 1.05) + 1024)
                nh = s.search(k)
            counts = peer["g"].finish(nh)
            return sum(counts)
        return nh

    h_hits = {"buf": None}

    e2e_phase = {"set_genome": 0.0, "set_library": 0.0, "build_index": 0.0, "search": 0.0, "copy_hits": 0.0}

    # N > 1, one library: every rank uploads only ITS 1/N slice of the host buffers and the slices are
    # all-gathered over NVLink (NCCL), so the host->device bytes of the job stay G + n*L in total
    # instead of N times that
    share = world > 1 and shard in ("slots", "genome")
    if share:
        gpad = (G + world - 1) // world * world
        lpad = (n * L + world - 1) // world * world
        d_g_full = torch.empty(gpad, dtype=torch.uint8, device=device)
        d_l_full = torch.empty(lpad, dtype=torch.uint8, device=device)
        g_lo, g_hi = rank * (gpad // world), min(G, (rank + 1) * (gpad // world))
        l_lo, l_hi = rank * (lpad // world), min(n * L, (rank + 1) * (lpad // world))

    def step_e2e():
        t0 = time.time()
        if share:
            part = d_g_full[rank * (gpad // world):(rank + 1) * (gpad // world)]
            if g_hi > g_lo:
                part[:g_hi - g_lo].copy_(h_genome[g_lo:g_hi], non_blocking=True)      # H2D of this rank's slice
            dist.all_gather_into_tensor(d_g_full, part)
            torch.cuda.current_stream().synchronize()
            s.set_genome_device(d_g_full.data_ptr(), off)                            # pack
        else:
            s.set_genome_array(h_genome.numpy(), off)        # H2D + pack
        t1 = time.time()
        if share:
            part = d_l_full[rank * (lpad // world):(rank + 1) * (lpad // world)]
            if l_hi > l_lo:
                part[:l_hi - l_lo].copy_(h_lib[l_lo:l_hi], non_blocking=True)
            dist.all_gather_into_tensor(d_l_full, part)
            torch.cuda.current_stream().synchronize()
            s.set_library_device(d_l_full.data_ptr(), n, L)
        else:
            s.set_library(h_lib.numpy().reshape(n, L))       # H2D + pack
        t2 = time.time()
        s.build_index(k)
        t3 = time.time()
        if h_hits["buf"] is None:                              # first (untimed) call: size the pinned result buffer
            nh = s.search(k)
            h_hits["buf"] = torch.empty((int(nh * 1.05) + 1024, 4), dtype=torch.int32).pin_memory()
            t4 = time.time()
            s.hits_into(h_hits["buf"].data_ptr(), h_hits["buf"].shape[0])
            # from now on the records are streamed to the pinned buffer while the search runs
            s.set_hit_sink(h_hits["buf"].data_ptr(), h_hits["buf"].shape[0])
        else:
            nh = s.search(k)                                   # D2H of the records happens inside (hit sink)
            t4 = time.time()
        t5 = time.time()
        for key, dt in zip(e2e_phase, (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4)):
            e2e_phase[key] += dt * 1e3
        return nh, h_hits["buf"]

    for _ in range(args.warmup):
        step_resident()
    st0 = s.stats()
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.25)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.time()
    ev0.record()
    acc = {"ms_build_index": 0.0, "ms_search": 0.0, "ms_scan_kernel": 0.0, "ms_genome_bucket": 0.0,
           "ms_win_count": 0.0, "ms_win_bin": 0.0, "ms_win_place": 0.0, "ms_finish": 0.0}
    launches = 0
    total_hits = 0
    step_wall = []
    for _ in range(args.steps):
        t_s = time.time()
        total_hits = step_resident()
        step_wall.append(round((time.time() - t_s) * 1e3, 2))
        st = s.stats()
        for key in acc:
            acc[key] += st[key]
        launches += st["index_launches"] + st["scan_launches"]
    ev1.record()
    barrier()
    t_wall1 = time.time()
    clocks = sampler.stop(t_wall0, t_wall1)
    ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_per_step = ms.item() / args.steps
    guides_total = n * world if shard == "library" else n
    value = guides_total * (G / 1e6) / (ms_per_step / 1e3)
    for key in acc:
        acc[key] /= args.steps
    st = s.stats()
    merged_ok = None
    if peer["g"] is not None:
        # Order-independent checksum of the 16-byte records: every rank hashes what ITS search produced
        # (its own device buffer), the hashes are summed over the ranks, and rank 0 compares the sum with
        # the hash of what actually arrived in its peer buffer.  Untimed.
        def digest(t):
            if t.shape[0] == 0:
                return torch.zeros(2, dtype=torch.int64, device=device)
            v = t.to(torch.int64) & 0xffffffff
            h = (v[:, 0] * 0x9E3779B1) ^ (v[:, 1] * 0x85EBCA77) ^ (v[:, 2] * 0xC2B2AE3D) ^ (v[:, 3] * 0x27D4EB2F)
            h = (h ^ (h >> 29)) * 0x165667B19E3779F9
            return torch.stack([h.sum(), torch.tensor(t.shape[0], dtype=torch.int64, device=device)])
        mine = digest(multi_gpu.hits_as_tensor(s, device))
        dist.all_reduce(mine, op=dist.ReduceOp.SUM)
        if rank == 0:
            got = sum((digest(seg) for seg in peer["g"].segments()), torch.zeros(2, dtype=torch.int64, device=device))
            merged_ok = bool(torch.equal(got, mine)) and int(mine[1].item()) == int(total_hits)
        peer["g"].close()

    # ---- end-to-end arm: host buffers in, host records out
    e2e_steps = max(2, min(args.steps, 3)) if not args.no_e2e else 0
    nh_e2e = 0
    if e2e_steps:
        step_e2e()                      # untimed: allocates the pinned result buffer
        for key in e2e_phase:
            e2e_phase[key] = 0.0
    barrier()
    ev0.record()
    for _ in range(e2e_steps):
        nh_e2e, out = step_e2e()
    ev1.record()
    barrier()
    e2e_phase_ms = {key: round(v / e2e_steps, 2) for key, v in e2e_phase.items()} if e2e_steps else {}
    s.set_hit_sink(None, 0)
    e2e_steps = max(e2e_steps, 1)
    ms_e = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(ms_e, op=dist.ReduceOp.MAX)
    e2e_value = guides_total * (G / 1e6) / (ms_e.item() / e2e_steps / 1e3)
    h2d = int((g_hi - g_lo) + (l_hi - l_lo)) if share else int(G + n * L)    # per rank; the job moves G + n*L in total
    d2h = int(nh_e2e * 16)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel group (DESIGN.md sections 4 and 6)
    peaks, peak_src = measured_peaks()
    combos = st["combos"]
    E = 2.0 * n
    records = float(G) * combos / (world if shard in ("slots", "genome") else 1)   # (window, combination) records of this rank
    entries = E * combos / (world if shard == "slots" else 1)
    rec_b = 8.0 if st["path"] == 3 else 16.0
    if st["path"] == 3:
        # per-kernel split of the compact join (bc_stats; CUDA events on the library's stream)
        parts = {"k_cverify": acc["ms_scan_kernel"] - acc["ms_finish"], "k_cfinish": acc["ms_finish"],
                 "k_ccount": acc["ms_win_count"], "k_cbin": acc["ms_win_bin"], "k_cplace": acc["ms_win_place"],
                 "index_build": acc["ms_build_index"]}
        # algorithmic HBM bytes per launch: count = k_cbincount (planes) + k_cslotcount (every record read once); pass A
        # reads the planes and writes every record (8 B); pass B reads and writes every record; verify reads the
        # records' x word... (all 8 B sectors) and the index entries (8 B) once; finish writes the hits; the index
        # build moves 8 B records three times (write, count, read) and writes 12 B per entry
        alg_bytes = {"k_ccount": 3 * G / 8 + rec_b * records, "k_cbin": 3 * G / 8 + rec_b * records, "k_cplace": 2 * rec_b * records,
                     "k_cverify": rec_b * records + 8.0 * entries, "k_cfinish": 16.0 * st["hits"],
                     "index_build": (8 + 8 + 8 + 8 + 12) * entries}
        names = {k2: k2 for k2 in parts}
        names["k_ccount"] = "k_cbincount<win>+k_cslotcount"
        names["index_build"] = "k_cbincount<lib>+k_cbin<lib>+k_cslotcount+k_cplace_bulk<lib>"
        verify_key = "k_cverify"
    else:
        parts = {"verify": acc["ms_scan_kernel"], "window_sort": acc["ms_genome_bucket"],
                 "index_build": acc["ms_build_index"]}
        verify_key = "verify"
        if st["path"] == 2:
            alg_bytes = {"window_sort": 2 * (3 * G / 8) + 3 * rec_b * records,
                         "verify": rec_b * records + 12.0 * entries + 16.0 * st["hits"],
                         "index_build": (8 + 16 + 16 + 12) * entries}
            names = {"verify": "k_verify_dense+k_verify_sparse", "window_sort": "k_bucket<0>+scan+k_window_bin+k_window_place",
                     "index_build": "k_index_count+scan+k_index_scatter+k_fine_scatter"}
        else:
            alg_bytes = {"verify": 3 * G / 8 + 16.0 * st["hits"], "window_sort": 0.0,
                         "index_build": (8 + 16 + 16 + 12) * entries}
            names = {"verify": "k_scan_probe", "window_sort": "-", "index_build": "k_index_count+scan+scatter"}
    dominant = max(parts, key=parts.get)
    ipk = int_peak(local_rank)
    traffic, traffic_note = None, None
    try:  # DRAM bytes per launch from the committed ncu --set full capture of this configuration
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as h:
            tj = json.load(h)
        if tj.get("source_hash") == source_hash():
            traffic = tj.get("entries", {}).get(f"{args.config}:key{st['key_nt']}:c{combos}:path{st['path']}:n{world}", {}).get(dominant)
        else:
            traffic_note = "profiles/traffic.json was captured from different kernel sources; not used"
    except (OSError, ValueError):
        pass
    kernel_ms = parts[dominant]
    achieved = alg_bytes[dominant] / (kernel_ms / 1e3) / 1e9 if kernel_ms > 0 else 0.0
    roofline_hbm = {"bound": "hbm", "kernel": names[dominant], "achieved": achieved, "peak": peaks["hbm_gbs"],
                    "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"], "traffic": traffic, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": alg_bytes[dominant],
                    "share_of_step": kernel_ms / max(ms_per_step, 1e-9)}
    if traffic_note:
        roofline_hbm["traffic_note"] = traffic_note
    roofline_int = None
    if ipk and st["path"] >= 2 and acc["ms_scan_kernel"] > 0:
        # candidates verified per second against the measured POPC issue rate and the measured
        # verify-atom rate (2 LOP3 + POPC + compare fed from a shared-memory broadcast)
        s.set_param(_native.BC_PARAM_COUNT_CANDIDATES, 1)
        s.search(k)
        cand = s.stats()["candidates"]
        s.set_param(_native.BC_PARAM_COUNT_CANDIDATES, 0)
        rate = cand / (parts[verify_key] / 1e3)
        pairs_k1 = float(G) * E / (world if shard != "library" else 1) * (k + 1) / 4.0 ** (L // (k + 1)) if L >= k + 1 else None
        roofline_int = {"bound": "int_popc", "kernel": names[verify_key], "achieved": rate / 1e12,
                        "peak": ipk["popc_per_s"] / 1e12, "unit": "Tpairs/s (1 POPC per pair)",
                        "frac": rate / ipk["popc_per_s"], "frac_of_verify_atom": rate / ipk["verify_atom_per_s"],
                        "verify_atom_peak": ipk["verify_atom_per_s"] / 1e12, "candidates": cand,
                        "pairs_vs_k_plus_1": (cand / pairs_k1) if pairs_k1 else None,
                        "traffic": traffic if dominant == verify_key else None,
                        "share_of_step": parts[verify_key] / max(ms_per_step, 1e-9),
                        "peak_source": "bench_kernels/int_peak.cu measured in this run (POPC: 16/clk/SM)",
                        "note": "peak = one POPC per candidate pair; achieved counts the ALGORITHMIC pairs (windows x entries per "
                                "slot) - padded lanes of ragged tiles are work the kernel does but is not credited for; "
                                "pairs_vs_k_plus_1 = candidates / what the classic k+1-seed filter would verify"}
    roofline_gather = None
    if st["path"] == 1 and acc["ms_scan_kernel"] > 0:
        # probe kernel: divergent loads (directory probes + bucket entries, one 128-byte line per lane each) per second
        # against the rate measured by the gather microbenchmark on an L2-resident table of the directories' size
        s.set_param(_native.BC_PARAM_COUNT_CANDIDATES, 1)
        s.search(k)
        stc = s.stats()
        s.set_param(_native.BC_PARAM_COUNT_CANDIDATES, 0)
        gathers = float(stc["probes"] + stc["candidates"])
        table_mb = 64
        gpk = gather_peak(local_rank, table_mb)
        if gpk:
            rate = gathers / (parts[verify_key] / 1e3)
            roofline_gather = {"bound": "l1_gather", "kernel": names[verify_key], "achieved": rate / 1e9, "peak": gpk / 1e9,
                               "unit": "G divergent loads/s", "frac": rate / gpk, "probes": stc["probes"],
                               "bucket_entries_read": stc["candidates"], "traffic": traffic if dominant == verify_key else None,
                               "share_of_step": parts[verify_key] / max(ms_per_step, 1e-9),
                               "peak_source": f"bench_kernels/int_peak.cu k_gather measured in this run ({table_mb} MB table, "
                                              "8 independent loads in flight per thread)",
                               "note": "every window costs one directory probe per seed combination plus one load per entry "
                                       "of a non-empty bucket; each is its own L1TEX wavefront"}
    # `roofline` = the bound that binds the dominant stage: the POPC pipe for verification, the divergent-load rate for
    # the probe kernel, HBM otherwise
    if dominant == verify_key and roofline_int:
        roofline = dict(roofline_int, hbm_view=roofline_hbm)
    elif dominant == verify_key and roofline_gather:
        roofline = dict(roofline_gather, hbm_view=roofline_hbm)
    else:
        roofline = dict(roofline_hbm, int_view=roofline_int)
    roofline["stage_ms"] = {k2: round(v, 4) for k2, v in parts.items()}

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline_run(cfg, genome, off, lib)

    verified = None
    if args.verify:
        from oracle import oracle
        from barcoder_b200 import synth
        n_s, G_s = min(n, 3000), min(G, 3_000_000)
        ref = oracle.search([bytes(genome[:G_s])], synth.rows_to_strings(lib[:n_s]), k, pam=cfg["pam"],
                            flags=oracle.PAM_FLAG_IUPAC if cfg["iupac"] else 0)
        with _native.Searcher(local_rank) as s2:
            s2.set_genome_array(genome[:G_s], np.array([0, G_s], dtype=np.uint64))
            s2.set_library(lib[:n_s])
            s2.set_pam(cfg["pam"], "downstream", iupac=cfg["iupac"])
            s2.search(k)
            got = _native.canonical_sort(s2.hits())
        verified = bool(got.tobytes() == ref.tobytes())

    line = {
        "metric": "guides*Mbp/s at <=k mismatches", "value": value, "unit": "guides*Mbp/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak" if shard == "library" else "strong", "vs_baseline": None, "merged_ok": merged_ok,
        "dtype": "u32 bit-planes (XOR/popcount)", "data": "synthetic",
        "config": {"workload": cfg["name"], "k": k, "pam": cfg["pam"], "spacers_per_gpu": n, "genome_bp": G,
                   "L": L, "pam_gate": bool(args.gate),
                   "parallelism": {"library": f"library-shard x{world} (a {n}-spacer library per GPU), genome replicated",
                                   "genome": f"genome-range x{world}, library replicated",
                                   "slots": f"seed-directory slot-range x{world}: genome and library replicated, index / "
                                            "window sort / verification sharded by seed key"}[shard],
                   "seed_scheme": f"b={st['blocks']} blocks, {combos} combinations, key<={st['key_nt']} nt, path={st['path']}",
                   "l2": "working set (window records + index) is far larger than the 126 MB L2; no flush needed",
                   "hits_per_step": int(total_hits), "host_numa_binding": numa_note},
        "e2e": {"value": e2e_value, "unit": "guides*Mbp/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e.item() / e2e_steps, "host_phase_ms": e2e_phase_ms},
        "gpu_launches": int(launches),
        "clocks": clocks, "roofline": roofline,
        "cpu_baseline": cpu,
        "stage_ms": dict({k2: round(v, 4) for k2, v in acc.items()},
                         ms_pack_genome=round(s.stats()["ms_pack_genome"], 4),
                         ms_pack_library=round(s.stats()["ms_pack_library"], 4)),
        "step_wall_ms": step_wall,
    }
    if verified is not None:
        line["verified_vs_oracle"] = verified
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
