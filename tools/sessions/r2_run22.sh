#!/bin/bash
# GPU session 22: GEMV kbench of the final library; ncu launch list of the bench step (GEMV kernels only)
timeout 300 python tools/kbench.py --only gemv > gpurun_out/r22_kbench_gemv.jsonl 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/r22_kbench_gemv.jsonl'):
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print(d['kernel'], d['us'], d.get('GBps'), d.get('hbm_frac'))
PY
timeout 600 python bench.py --steps 3 --warmup 3 --no-extras > gpurun_out/r22_bench_short.json 2>/dev/null && timeout 900 ncu --kernel-name-base demangled -k regex:k_gemv4 --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/r22_launches.csv python bench.py --steps 3 --warmup 3 --no-extras > gpurun_out/r22_ncu.log 2>&1; tail -1 gpurun_out/r22_ncu.log | cut -c1-200; wc -l gpurun_out/r22_launches.csv
