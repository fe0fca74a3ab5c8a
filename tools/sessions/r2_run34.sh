#!/bin/bash
# GPU session 34: tiled int8 layout kernels -- parity, then kbench int8
timeout 600 python -m pytest tests/test_gpu_int8.py -x -q -m gpu > gpurun_out/r34_t.log 2>&1; tail -4 gpurun_out/r34_t.log
timeout 300 python tools/kbench.py --only int8 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print(d['kernel'], d['us'], d.get('hbm_frac'), d.get('TOPS'))"
