#!/bin/bash
# GPU session 1 of round 2: PDL hand-off probe, parity of the v2 GEMV, A/B of kernels and of the next-weight prefetcher
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2_gpu.txt
tools/bin/pdl_probe > gpurun_out/r2_pdl_probe.jsonl 2>&1
python -m pytest tests/test_gpu_gemv.py -x -q -m gpu > gpurun_out/r2_t1.log 2>&1; tail -3 gpurun_out/r2_t1.log
for cfg in "2 16 1" "2 16 0" "2 8 1" "2 8 0" "b 0 1" "b 0 0"; do
  set -- $cfg
  BNB_B200_GEMV_IMPL=$1 BNB_B200_GEMV_V2W=$2 BNB_B200_GEMV_NEXTPF=$3 python tools/kbench.py --only gemv > gpurun_out/r2_kbench_$1_w$2_pf$3.jsonl 2>&1
  BNB_B200_GEMV_IMPL=$1 BNB_B200_GEMV_V2W=$2 BNB_B200_GEMV_NEXTPF=$3 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_$1_w$2_pf$3.json 2> gpurun_out/r2_bench_$1_w$2_pf$3.err
done
cat gpurun_out/r2_pdl_probe.jsonl
for f in gpurun_out/r2_bench_*.json; do echo $f; python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['value'], d['roofline']['frac'], d['e2e']['value'], d.get('fused_same_input',{}).get('value'))"; done
