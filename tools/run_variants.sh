# usage: VARS="a b" ARGS="--path 3 --key-nt 9" bash tools/run_variants.sh
# benches every bench_kernels/var_<name>.so listed in $VARS (A/B builds of the library, tools/build_variant.sh)
for v in ${VARS}; do
  echo "== $v $ARGS"
  BARCODER_B200_LIB=bench_kernels/var_$v.so python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e $ARGS 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],2), d['stage_ms'], d['config']['hits_per_step'])"
done
