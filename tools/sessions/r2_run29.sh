#!/bin/bash
# GPU session 29: round-1 kernel removed, dispatcher small (<= 32) / wide
timeout 900 python -m pytest tests/test_gpu_gemm4.py -q -m gpu > gpurun_out/r29_t.log 2>&1; tail -3 gpurun_out/r29_t.log
timeout 300 python tools/gemm4_stress.py 40 2>&1 | tail -6
for g in 4; do
  BNB_B200_GEMM4_WIDE_G=$g timeout 300 python tools/kbench.py --only gemm4 > gpurun_out/r29_kbench_gemm4_g$g.jsonl 2>&1
  echo "--- shipped routing"
  python - <<PY
import json
for l in open('gpurun_out/r29_kbench_gemm4_g$g.jsonl'):
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print(d['kernel'], d['us'], d.get('TFLOPs'), d.get('bf16_frac'), d.get('speedup_vs_composition'))
PY
done
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
