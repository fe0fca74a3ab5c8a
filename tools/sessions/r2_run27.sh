#!/bin/bash
# GPU session 27: wide kernel (batch 65..256, scaled operand in TMEM) -- tests, cold-cache stress, same-box A/B against round 1's kernel
timeout 900 python -m pytest tests/test_gpu_gemm4.py -q -m gpu > gpurun_out/r27_t.log 2>&1; tail -4 gpurun_out/r27_t.log
timeout 300 python tools/gemm4_stress.py 60 2>&1 | tail -6
for w in 1 0; do
  BNB_B200_GEMM4_WIDE=$w timeout 300 python tools/kbench.py --only gemm4 > gpurun_out/r27_kbench_gemm4_wide$w.jsonl 2>&1
  echo "--- WIDE=$w"
  python - <<PY
import json
for l in open('gpurun_out/r27_kbench_gemm4_wide$w.jsonl'):
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    if '_b16' in d['kernel'] or '_b32' in d['kernel']: continue
    print(d['kernel'], d['us'], d.get('TFLOPs'), d.get('bf16_frac'), d.get('speedup_vs_composition'))
PY
done
BNB_B200_GEMM4_SMALL=0 timeout 300 python tools/kbench.py --only gemm4 2>&1 | python -c "
import sys,json
print('--- SMALL=0 (wide kernel at every batch)')
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print(d['kernel'], d['us'], d.get('speedup_vs_composition'))"
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
