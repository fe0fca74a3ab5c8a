#!/bin/bash
# GPU session 8: is the L2 the lever?  round-1 kernel with weights resident in L2 (2 rotating buffers) vs streamed; hint stats
BNB_B200_GEMV_PROBE=1 BNB_B200_GEMV_IMPL=b BNB_B200_GEMV_NEXTPF=0 timeout 300 python tools/kbench.py --only gemv > gpurun_out/r8_kbench_bc_stream.jsonl 2>&1
KBENCH_NBUF=2 BNB_B200_GEMV_PROBE=1 BNB_B200_GEMV_IMPL=b BNB_B200_GEMV_NEXTPF=0 timeout 300 python tools/kbench.py --only gemv > gpurun_out/r8_kbench_bc_l2.jsonl 2>&1
BNB_B200_GEMV_PROBE=1 BNB_B200_GEMV_IMPL=b BNB_B200_GEMV_NEXTPF=1 timeout 300 python tools/kbench.py --only gemv > gpurun_out/r8_kbench_bc_pf.jsonl 2>&1
timeout 900 python -m pytest tests/test_gpu_int8.py tests/test_gpu_gemm4.py -x -q -m gpu > gpurun_out/r8_t.log 2>&1; tail -2 gpurun_out/r8_t.log
python tools/kbench.py --only int8,gemm4 > gpurun_out/r8_kbench_int8_gemm4.jsonl 2>&1
for f in gpurun_out/r8_kbench_bc_*.jsonl; do echo "== $f"; python - "$f" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    try: d=json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    print(d['kernel'][24:], d['us'], d.get('hbm_frac'), d.get('cta_us'), d.get('hint_stats'), {k: round(v,2) for k,v in d.get('phase_us',{}).items()})
PY
done
python - <<'PY'
import json
for l in open('gpurun_out/r8_kbench_int8_gemm4.jsonl'):
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print(d['kernel'], d['us'], d.get('TOPS'), d.get('hbm_frac'), d.get('bf16_frac'), d.get('speedup_vs_composition'))
PY
