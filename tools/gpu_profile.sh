#!/bin/bash
# Round artefacts on the GPU box: launch list of a bench step and one --set full capture of every pipeline kernel
# (exported to CSV on the box: the report files are too large to travel).
set -x
export QB_BENCH_TABLES=49096
python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-pmp > gpurun_out/r2_launch_plain.json 2> gpurun_out/r2_launch_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-pmp > gpurun_out/r2_launch_ncu.log 2>&1
unset QB_BENCH_TABLES
python tools/gpu_perf.py 56832 > gpurun_out/r2_full_plain.log 2>&1 &&
ncu --set full --clock-control none -k 'regex:k_(series|scale|horner|rec_coeffs|pow|rho_to_z|offdiag_val|offdiag_close|coef_fx)' -c 26 \
    -o /tmp/r2_full python tools/gpu_perf.py 56832 > gpurun_out/r2_full_ncu.log 2>&1
ncu -i /tmp/r2_full.ncu-rep --page raw --csv > gpurun_out/r2_full_raw.csv 2>/dev/null
ls -la /tmp/r2_full.ncu-rep gpurun_out/r2_full_raw.csv
tail -3 gpurun_out/r2_full_ncu.log
