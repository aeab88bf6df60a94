#!/usr/bin/env bash
# tools/gpu_narrow.sh -- run under gpurun: narrow-plane tests, example driver both ways, bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_narrow.py tests/test_gpu_fused.py -x -q 2>&1 | tail -15
(cd vcfgl_b200/host && ./example_driver 9 > /tmp/wide.txt && VGL_NARROW=1 ./example_driver 9 > /tmp/narrow.txt && cmp /tmp/wide.txt /tmp/narrow.txt && echo "example_driver: narrow == wide" && head -3 /tmp/narrow.txt)
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_narrow.json 2> gpurun_out/bench_narrow.err; tail -c 1500 gpurun_out/bench_narrow.json; tail -3 gpurun_out/bench_narrow.err
