#!/bin/bash
# GPU session 40: stream-K (one equal range of rounds per SM, last-arriver fixup) in the small-batch 4-bit GEMM
timeout 900 python -m pytest tests/test_gpu_gemm4.py -x -q -m gpu > gpurun_out/r40_t.log 2>&1; tail -4 gpurun_out/r40_t.log
timeout 300 python tools/gemm4_stress.py 60 2>&1 | tail -6 | cut -c1-120
timeout 300 python tools/kbench.py --only gemm4 > gpurun_out/r40_kbench_gemm4.jsonl 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/r40_kbench_gemm4.jsonl'):
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print(d['kernel'], d['us'], d.get('hbm_frac'), d.get('speedup_vs_composition'))
PY
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
