#!/bin/bash
# BCF output path on the GPU box: parity tests, bench lines with the bcf_records leg, launch list of an e2e run
mkdir -p gpurun_out/bcf
python -m pytest tests/test_gpu_bcf.py -x -q 2>&1 | tail -3
python bench.py --no-cpu-baseline > gpurun_out/bcf/bench_cfg2.json 2> gpurun_out/bcf/bench_cfg2.err
python bench.py --no-cpu-baseline --workload cfg5 > gpurun_out/bcf/bench_cfg5.json 2> gpurun_out/bcf/bench_cfg5.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/bcf/launches_e2e.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bcf/b.log 2>&1
python - <<'PY'
import json
for w in ("cfg2", "cfg5"):
    d = json.loads(open("gpurun_out/bcf/bench_%s.json" % w).read().strip().splitlines()[-1])
    e = d["e2e"]
    print(w, "value %.3g" % d["value"], "e2e narrow %.3g" % e["value"], "i32 %.3g" % e["i32_planes"]["value"], "bcf", e["bcf_records"])
PY
grep -v "^==" gpurun_out/bcf/launches_e2e.csv | awk -F'","' 'NR>1{n[$5]++; t[$5]+=$NF+0} END{for(k in n) printf "%s x%d avg %.1f us\n", k, n[k], t[k]/n[k]/1000}'
