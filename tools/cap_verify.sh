mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"k_cverify|k_cfinish" -c 2 -o gpurun_out/full_verify -f python tools/ab_bench.py --steps 1 cur > gpurun_out/full_verify.out 2>&1
ncu -i gpurun_out/full_verify.ncu-rep --page raw --csv > gpurun_out/full_verify_raw.csv
du -sm gpurun_out/full_verify.ncu-rep
