#!/usr/bin/env bash
# tools/gpu_bench.sh <tag> [bench args] -- one bench line, parsed summary
OUT=gpurun_out; mkdir -p $OUT; TAG="$1"; shift
timeout 1500 python bench.py "$@" > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"; tail -c 800 $OUT/bench_$TAG.err
python - <<P
import json
d=json.loads(open("$OUT/bench_$TAG.json").read().strip().splitlines()[-1])
e=d["e2e"]
print("value %.3g frac %.3f ms/step %.2f"%(d["value"],d["roofline"]["frac"],d["ms_per_step"]))
print("e2e %.3g (%s...) d2h/step %.3g"%(e["value"],e["planes"][:14],e["d2h_bytes_per_step"]))
for k in ("narrow_planes","i32_planes","bcf_records"):
    if e.get(k): print("  ",k,"%.3g"%e[k]["value"],"d2h %.3g"%e[k]["d2h_bytes_per_step"])
for k,v in (d.get("configs") or {}).items(): print(k,"%.3g cells/s kernel_ms %.3f frac %.3f"%(v["value"],v["kernel_ms"],v["roofline"]["frac"]), v.get("gvcf_merge",{}).get("kernel_ms"))
print("cpu",d["cpu_baseline"]); print("clk",d["clocks"])
P
