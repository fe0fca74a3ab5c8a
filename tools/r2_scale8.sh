#!/bin/bash
# round 2, final library: whole-job numbers at N GPUs (argument), 7B stack (headline workload) and 70B-shape stack
N=${1:-8}
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N "$@" 2>gpurun_out/scale_n${N}.err | tail -1; }
run --steps 30 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/scale_n${N}_7b.json
run --steps 30 --warmup 5 --no-extras --no-cpu-baseline --workload llama3-70b --layers 8 > gpurun_out/scale_n${N}_70b.json
python - <<PY
import json
for f in ("7b","70b"):
    try:
        d=json.loads(open(f"gpurun_out/scale_n${N}_{f}.json").read().strip().splitlines()[-1])
        print(f, d["n_gpus"], round(d["value"],1), d["unit"], "ms", round(d["ms_per_step"],4), d["config"].get("collective"), d.get("parity_checked"), d.get("clocks",{}).get("reasons"))
    except Exception as e:
        print(f, "failed", e); print(open("gpurun_out/scale_n${N}.err").read()[-1500:])
PY
