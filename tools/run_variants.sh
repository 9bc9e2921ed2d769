# usage: bash tools/run_variants.sh  - benches every bench_kernels/var_<name>.so listed in $VARS (A/B builds of the library)
for v in ${VARS:-k3 f3 f4}; do
  echo "== $v"
  BARCODER_B200_LIB=bench_kernels/var_$v.so python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['stage_ms']['ms_scan_kernel'], d['stage_ms']['ms_genome_bucket'], d['config']['hits_per_step'])"
done
