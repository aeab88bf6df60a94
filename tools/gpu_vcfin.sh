#!/usr/bin/env bash
# tools/gpu_vcfin.sh -- run under gpurun: input-path tests (first under compute-sanitizer), the whole gpu suite, bench, ncu of k_vcf_*.
set -u
TAG="${1:-r01q}"
OUT=gpurun_out
mkdir -p $OUT
timeout 600 compute-sanitizer --error-exitcode 9 python -m pytest tests/test_gpu_vcfin.py -x -q -k "single_records or chunking or test_reference_inputs" > $OUT/sanitizer_$TAG.log 2>&1
echo "sanitizer rc=$?"; tail -4 $OUT/sanitizer_$TAG.log
timeout 900 python -m pytest tests/test_gpu_vcfin.py -x -q > $OUT/pytest_vcfin_$TAG.log 2>&1; echo "vcfin rc=$?"; tail -15 $OUT/pytest_vcfin_$TAG.log
python tools/prof_vcfin.py 100 131072 2>&1 | tail -2
python tools/prof_vcfin.py 10000 4440 2>&1 | tail -2
python tools/prof_vcfin.py 1000 8192 2>&1 | tail -2
if [ "${FULL:-1}" = "1" ]; then
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu_$TAG.log; tail -3 $OUT/pytest_gpu_$TAG.log
python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; tail -c 3000 $OUT/bench_$TAG.json; tail -3 $OUT/bench_$TAG.err
fi
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches_vcfin_$TAG.csv python tools/prof_vcfin.py 100 131072 3 > $OUT/ncu_launches_vcfin_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_vcf_hdr -s 2 -c 1 -f -o $OUT/prof_vcfhdr_$TAG python tools/prof_vcfin.py 100 131072 3 > $OUT/ncu_vcfgt_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_vcf_hdr -s 2 -c 1 -f -o $OUT/prof_vcfhdr_$TAG python tools/prof_vcfin.py 100 131072 3 > $OUT/ncu_vcflines_$TAG.log 2>&1
ls -la $OUT | tail -12
