#!/usr/bin/env bash
# tools/gpu_r2b.sh -- full GPU suite + the driver's bench commands (both arms)
set -u
OUT=gpurun_out
mkdir -p $OUT
TAG="${1:-r2b}"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $OUT/pytest_$TAG.txt
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"; tail -c 600 $OUT/bench_$TAG.err
python - <<P
import json
try:
    d=json.loads(open("$OUT/bench_$TAG.json").read().strip().splitlines()[-1])
    print("value %.3g frac %.3f ms/step %.2f e2e %.3g (i32 %.3g bcf %s)"%(d["value"],d["roofline"]["frac"],d["ms_per_step"],d["e2e"]["value"],d["e2e"]["i32_planes"]["value"],d["e2e"]["bcf_records"] and "%.3g"%d["e2e"]["bcf_records"]["value"]))
    for k,v in (d.get("configs") or {}).items(): print(k,"%.3g cells/s kernel_ms %.3f frac %.3f"%(v["value"],v["kernel_ms"],v["roofline"]["frac"]), v.get("gvcf_merge",{}).get("kernel_ms"))
    print("cpu",d["cpu_baseline"]); print("clk",d["clocks"])
except Exception as e: print("parse failed",e)
P
if [ "${2:-}" = "ref" ]; then timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref_$TAG.json 2>&1; tail -c 900 $OUT/bench_ref_$TAG.json; fi
