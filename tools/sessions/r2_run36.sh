#!/bin/bash
# GPU session 36: persistent prefetching 4-bit quantize kernel; wrappers without zero fills
timeout 900 python -m pytest tests/test_gpu_blockwise.py tests/test_gpu_int8.py -x -q -m gpu > gpurun_out/r36_t.log 2>&1; tail -3 gpurun_out/r36_t.log
timeout 300 python tools/kbench.py --only quant 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print(d['kernel'], d['us'], d.get('hbm_frac'))"
timeout 300 python tools/kbench.py --only int8 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print(d['kernel'], d['us'], d.get('hbm_frac'), d.get('TOPS'))"
