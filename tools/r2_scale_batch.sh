#!/bin/bash
# BASELINE config 5 at batch > 1: 70B-shape stack (8 layers), N-sharded, batch 8 and 64, on N GPUs (argument)
N=${1:-1}
for B in 8 64; do
  if [ "$N" = "1" ]; then
    timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-extras --no-cpu-baseline --workload llama3-70b --layers 8 --batch $B 2>gpurun_out/scaleb_n${N}.err | tail -1 > gpurun_out/scaleb_n${N}_b${B}.json
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 20 --warmup 5 --no-extras --no-cpu-baseline --workload llama3-70b --layers 8 --batch $B 2>gpurun_out/scaleb_n${N}.err | tail -1 > gpurun_out/scaleb_n${N}_b${B}.json
  fi
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/scaleb_n${N}_b${B}.json").read())
    print("N=${N} batch=${B}", round(d["value"],1), d["unit"], "ms", round(d["ms_per_step"],4), d["config"].get("collective"), d["config"].get("launch"))
except Exception as e:
    print("N=${N} batch=${B} failed", e); print(open("gpurun_out/scaleb_n${N}.err").read()[-1200:])
PY
done
