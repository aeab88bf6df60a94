#!/usr/bin/env bash
# tools/gpu_gvcf.sh -- run under gpurun: gVCF merger tests (first under compute-sanitizer), timing, launch list, ncu of k_gvcf_reduce
set -u
TAG="${1:-r01x}"
OUT=gpurun_out
mkdir -p $OUT
timeout 600 compute-sanitizer --error-exitcode 9 python -m pytest tests/test_gpu_gvcf.py -x -q -k "replayed or lowdepth" > $OUT/sanitizer_gvcf_$TAG.log 2>&1
echo "sanitizer rc=$?"; tail -4 $OUT/sanitizer_gvcf_$TAG.log
timeout 900 python -m pytest tests/test_gpu_gvcf.py -x -q > $OUT/pytest_gvcf_$TAG.log 2>&1; echo "gvcf rc=$?"; tail -15 $OUT/pytest_gvcf_$TAG.log
python tools/prof_gvcf.py 100 131072 2>&1 | tail -2
python tools/prof_gvcf.py 1000 8192 2>&1 | tail -2
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $OUT/launches_gvcf_$TAG.csv python tools/prof_gvcf.py 100 131072 2 > $OUT/ncu_launches_gvcf_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_gvcf_reduce -s 1 -c 1 -f -o $OUT/prof_gvcfreduce_$TAG python tools/prof_gvcf.py 100 131072 2 > $OUT/ncu_gvcfreduce_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_gvcf_key -s 1 -c 1 -f -o $OUT/prof_gvcfkey_$TAG python tools/prof_gvcf.py 100 131072 2 > $OUT/ncu_gvcfplan_$TAG.log 2>&1
ls -la $OUT | tail -6
