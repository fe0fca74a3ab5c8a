#!/bin/bash
# GPU session 35: one-pass row quantisation with packed half2 statistics and saturating conversions
timeout 600 python -m pytest tests/test_gpu_int8.py -x -q -m gpu > gpurun_out/r35_t.log 2>&1; tail -4 gpurun_out/r35_t.log
timeout 300 python tools/kbench.py --only int8 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print(d['kernel'], d['us'], d.get('hbm_frac'), d.get('TOPS'))"
