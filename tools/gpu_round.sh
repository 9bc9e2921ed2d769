#!/bin/bash
# One gpurun call's worth of evidence: GPU tests, bench lines, ncu launch list, ncu --set full captures.
# usage (from the repo root, on the GPU box):  bash tools/gpu_round.sh [tests] [bench] [cfg5] [class] [launches] [full]
# Everything lands in gpurun_out/ (merged back by gpurun); summaries are copied to profiles/ by hand.
mkdir -p gpurun_out
what="${*:-tests bench launches full}"
has() { [[ " $what " == *" $1 "* ]]; }
export PYTHONUNBUFFERED=1

if has tests; then
  timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
  echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
  tail -3 gpurun_out/pytest_gpu.log
fi
if has smoke; then
  timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
  echo "smoke rc=$?"; tail -4 gpurun_out/smoke.log
fi
if has bench; then
  timeout 600 python bench.py > gpurun_out/bench_cfg4.json 2> gpurun_out/bench_cfg4.err
  echo "bench rc=$?"; tail -c 600 gpurun_out/bench_cfg4.json
fi
if has ref; then
  timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
fi
if has cfg5; then
  timeout 900 python bench.py --config cfg5 --steps 3 --no-cpu-baseline > gpurun_out/bench_cfg5.json 2> gpurun_out/bench_cfg5.err
  echo "cfg5 rc=$?"; tail -c 400 gpurun_out/bench_cfg5.json
fi
if has class; then
  timeout 900 python bench.py --api class > gpurun_out/bench_class.json 2> gpurun_out/bench_class.err
  echo "class rc=$?"; tail -c 400 gpurun_out/bench_class.json
fi
if has launches; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/launches_cfg4.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e \
    > gpurun_out/launches_cfg4.out 2>&1
  echo "launches rc=$?"
fi
if has full; then
  # full-set capture of the step's kernels (first launches of each: index build + one search), exported to CSV on
  # the box; the .ncu-rep itself only travels back when it is small (gpurun_out/ is capped at 64 MiB)
  timeout 1500 ncu --set full --clock-control none --import-source on \
    -k regex:"^(k_cverify|k_cfinish|k_cbin|k_cplace_bulk|k_cbincount|k_cslotcount)$" -c ${NCU_COUNT:-10} \
    -o gpurun_out/full_cfg4 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e \
    > gpurun_out/full_cfg4.out 2>&1
  echo "full rc=$?"
  ncu -i gpurun_out/full_cfg4.ncu-rep --page raw --csv > gpurun_out/full_cfg4_raw.csv 2>/dev/null
  for kn in k_cverify k_cfinish k_cbin k_cplace_bulk k_cbincount k_cslotcount; do
    ncu -i gpurun_out/full_cfg4.ncu-rep --page source --csv --kernel-name regex:"^$kn\$" > gpurun_out/full_cfg4_src_$kn.csv 2>/dev/null
  done
  sz=$(du -m gpurun_out/full_cfg4.ncu-rep | cut -f1); echo "rep MiB: $sz"
  if [ "$sz" -gt 30 ]; then rm -f gpurun_out/full_cfg4.ncu-rep; fi
fi
if has full5; then
  timeout 1500 ncu --set full --clock-control none --import-source on -k regex:k_scan_probe -c 1 \
    -o gpurun_out/full_cfg5 -f python bench.py --config cfg5 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e \
    > gpurun_out/full_cfg5.out 2>&1
  echo "full5 rc=$?"
  ncu -i gpurun_out/full_cfg5.ncu-rep --page raw --csv > gpurun_out/full_cfg5_raw.csv 2>/dev/null
  ncu -i gpurun_out/full_cfg5.ncu-rep --page source --csv > gpurun_out/full_cfg5_src.csv 2>/dev/null
  sz=$(du -m gpurun_out/full_cfg5.ncu-rep | cut -f1); if [ "$sz" -gt 20 ]; then rm -f gpurun_out/full_cfg5.ncu-rep; fi
fi
du -sm gpurun_out
true
