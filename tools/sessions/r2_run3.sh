#!/bin/bash
# GPU session 3: TMA GEMV with a parallel producer -- sweep, then the new bench.py (extras) once
export BNB_B200_GEMV_NEXTPF=0
BNB_B200_GEMV_IMPL=T BNB_B200_GEMV_TW=16 timeout 600 python -m pytest tests/test_gpu_gemv.py -x -q -m gpu > gpurun_out/r2_t3_w16.log 2>&1; tail -2 gpurun_out/r2_t3_w16.log
for cfg in "16 1" "16 0" "24 1" "24 0" "20 0" "28 0"; do
  set -- $cfg
  BNB_B200_GEMV_IMPL=T BNB_B200_GEMV_TW=$1 BNB_B200_GEMV_XREG=$2 timeout 300 python tools/kbench.py --only gemv > gpurun_out/r3_kbench_T_w$1_x$2.jsonl 2>&1
done
for cfg in "16 1" "24 0"; do
  set -- $cfg
  BNB_B200_GEMV_IMPL=T BNB_B200_GEMV_TW=$1 BNB_B200_GEMV_XREG=$2 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r3_bench_T_w$1_x$2.json 2> gpurun_out/r3_bench_T_w$1_x$2.err
done
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r3_bench_default.json 2> gpurun_out/r3_bench_default.err
for f in gpurun_out/r3_kbench_T_*.jsonl; do echo "== $f"; python - "$f" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    try: d=json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    print(d['kernel'], d['us'], d.get('GBps'), d.get('hbm_frac'))
PY
done
for f in gpurun_out/r3_bench_*.json; do echo $f; python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['value'], d['roofline']['frac'], d['e2e']['value'], d.get('fused_same_input',{}).get('value')); print({k:d[k] for k in d if k in ('llama3_70b','igemmlt')})"; done
tail -3 gpurun_out/r3_bench_default.err
