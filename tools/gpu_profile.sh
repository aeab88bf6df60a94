#!/usr/bin/env bash
# tools/gpu_profile.sh -- run under gpurun: bench line, ncu launch list, one ncu --set full capture.
# usage: gpurun --timeout 1500 -- bash tools/gpu_profile.sh <tag> [kernel-regex]
set -u
TAG="${1:-r01}"
KRE="${2:-k_fused}"
OUT=gpurun_out
mkdir -p $OUT
python bench.py --no-cpu-baseline --workload ${WL:-cfg2} > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
tail -c 2500 $OUT/bench_$TAG.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --workload ${WL:-cfg2} --steps 2 --warmup 3 --skip-e2e > $OUT/ncu_launches_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:$KRE -s 5 -c 1 -f -o $OUT/prof_$TAG \
    python bench.py --workload ${WL:-cfg2} --steps 2 --warmup 3 --skip-e2e > $OUT/ncu_$TAG.log 2>&1
ls -la $OUT | tail -8
