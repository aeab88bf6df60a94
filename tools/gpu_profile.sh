#!/usr/bin/env bash
# tools/gpu_profile.sh -- run under gpurun: bench line, ncu launch list, one ncu --set full capture.
# usage: gpurun --timeout 1500 -- bash tools/gpu_profile.sh <tag>
set -u
TAG="${1:-r01}"
OUT=gpurun_out
mkdir -p $OUT
python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
tail -c 3000 $OUT/bench_$TAG.json
python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref_$TAG.json 2>> $OUT/bench_$TAG.err
tail -c 1500 $OUT/bench_ref_$TAG.json
# every launch with its device time (cold-cache, serialised: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --skip-e2e > $OUT/ncu_launches_$TAG.log 2>&1
# the top kernels, full set
ncu --set full --clock-control none --import-source on -k regex:k_emit -s 5 -c 1 -f -o $OUT/prof_emit_$TAG \
    python bench.py --steps 2 --warmup 3 --skip-e2e > $OUT/ncu_emit_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_sim -s 5 -c 1 -f -o $OUT/prof_sim_$TAG \
    python bench.py --steps 2 --warmup 3 --skip-e2e > $OUT/ncu_sim_$TAG.log 2>&1
ls -la $OUT
