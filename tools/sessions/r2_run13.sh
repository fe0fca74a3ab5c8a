#!/bin/bash
# GPU session 13: small-batch fused 4-bit GEMM (per-block scaling out of TMEM)
timeout 600 python -m pytest tests/test_gpu_gemm4.py -x -q -m gpu > gpurun_out/r13_t.log 2>&1; tail -15 gpurun_out/r13_t.log
timeout 300 python tools/kbench.py --only gemm4 > gpurun_out/r13_kbench_gemm4.jsonl 2>&1
BNB_B200_GEMM4_SMALL=0 timeout 300 python tools/kbench.py --only gemm4 > gpurun_out/r13_kbench_gemm4_old.jsonl 2>&1
python - <<'PY'
import json
for f in ('gpurun_out/r13_kbench_gemm4.jsonl','gpurun_out/r13_kbench_gemm4_old.jsonl'):
    print(f)
    for l in open(f):
        try: d=json.loads(l)
        except Exception: print(l.strip()[:200]); continue
        print(d['kernel'], d['us'], d.get('hbm_frac'), d.get('bf16_frac'), d.get('speedup_vs_composition'))
PY
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
