#!/usr/bin/env bash
# tools/gpu_numa.sh <N> -- e2e of N ranks with and without binding each rank to its GPU's NUMA node
N="$1"; OUT=gpurun_out; mkdir -p $OUT
nvidia-smi topo -m 2>&1 | head -14 | tee $OUT/topo.txt
nproc; numactl -H 2>/dev/null | head -6
for numa in 0 1; do
VGL_BENCH_NUMA=$numa timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 4 --warmup 3 --no-configs --no-cpu-baseline > $OUT/bench_numa${numa}_n$N.json 2> $OUT/bench_numa${numa}_n$N.err; echo "rc=$?"; tail -c 300 $OUT/bench_numa${numa}_n$N.err
python - <<P
import json
d=json.loads(open("$OUT/bench_numa${numa}_n$N.json").read().strip().splitlines()[-1])
e=d["e2e"]
print("numa=$numa N=$N value %.4g e2e %.4g narrow %.4g i32 %.4g bcf %.4g bound %s"%(d["value"],e["value"],e["narrow_planes"]["value"],e["i32_planes"]["value"],e["bcf_records"]["value"],d["config"].get("cpus_bound_near_gpu")))
P
done
