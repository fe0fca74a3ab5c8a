#!/bin/bash
# GPU session 4: v2 (x in registers / staged absmax / ring depth by shape) vs the round-1 kernel, with phase probes
export BNB_B200_GEMV_NEXTPF=0
for w in 8 16; do
  BNB_B200_GEMV_IMPL=2 BNB_B200_GEMV_V2W=$w timeout 600 python -m pytest tests/test_gpu_gemv.py -x -q -m gpu > gpurun_out/r4_t_w$w.log 2>&1; tail -2 gpurun_out/r4_t_w$w.log
done
for cfg in "2 8 1" "2 16 1" "2 8 0" "2 16 0" "2 8 2" "2 16 2" "b 0 0"; do
  set -- $cfg
  BNB_B200_GEMV_PROBE=1 BNB_B200_GEMV_IMPL=$1 BNB_B200_GEMV_V2W=$2 BNB_B200_GEMV_XREG=$3 timeout 300 python tools/kbench.py --only gemv > gpurun_out/r4_kbench_$1_w$2_x$3.jsonl 2>&1
done
for cfg in "2 8 1" "2 16 1" "2 8 0"; do
  set -- $cfg
  BNB_B200_GEMV_IMPL=$1 BNB_B200_GEMV_V2W=$2 BNB_B200_GEMV_XREG=$3 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r4_bench_$1_w$2_x$3.json 2> gpurun_out/r4_bench_$1_w$2_x$3.err
done
for f in gpurun_out/r4_kbench_*.jsonl; do echo "== $f"; python - "$f" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    try: d=json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    print(d['kernel'][24:], d['us'], d.get('hbm_frac'), d.get('cta_us'), {k: round(v,2) for k,v in d.get('phase_us',{}).items()})
PY
done
for f in gpurun_out/r4_bench_*.json; do echo $f; python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['value'], d['roofline']['frac'], d['e2e']['value'], d.get('fused_same_input',{}).get('value'))"; done
