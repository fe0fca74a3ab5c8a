#!/bin/bash
# GPU session 25: table on a 64 KB window boundary (PRMT-only lookup address); KB A/B again
timeout 900 python -m pytest tests/test_gpu_gemm4.py -q -m gpu > gpurun_out/r25_t.log 2>&1; tail -4 gpurun_out/r25_t.log
timeout 300 python tools/gemm4_stress.py 60 2>&1 | tail -6
for kb in 2 1; do
  BNB_B200_GEMM4_SMALL_KB=$kb timeout 300 python tools/kbench.py --only gemm4 > gpurun_out/r25_kbench_gemm4_kb$kb.jsonl 2>&1
  echo "--- KB=$kb"
  python - <<PY
import json
for l in open('gpurun_out/r25_kbench_gemm4_kb$kb.jsonl'):
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    if '_b128' in d['kernel'] or '_b256' in d['kernel']: continue
    print(d['kernel'], d['us'], d.get('hbm_frac'), d.get('speedup_vs_composition'))
PY
done
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
