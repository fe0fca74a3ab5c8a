#!/bin/bash
# GPU session 7: independent accumulator chains in v2; igemmlt reference ABI with the col32 epilogue
timeout 900 python -m pytest tests/test_gpu_int8.py -x -q -m gpu > gpurun_out/r7_t_int8.log 2>&1; tail -2 gpurun_out/r7_t_int8.log
for w in 8 16; do
  BNB_B200_GEMV_IMPL=2 BNB_B200_GEMV_V2W=$w timeout 600 python -m pytest tests/test_gpu_gemv.py -x -q -m gpu > gpurun_out/r7_t_w$w.log 2>&1; tail -2 gpurun_out/r7_t_w$w.log
done
BNB_B200_GEMV_IMPL=2 BNB_B200_GEMV_V2W=8 BNB_B200_GEMV_XREG=0 timeout 600 python -m pytest tests/test_gpu_gemv.py -x -q -m gpu > gpurun_out/r7_t_w8x0.log 2>&1; tail -2 gpurun_out/r7_t_w8x0.log
run() { # impl warps xreg pf nacc tag
  BNB_B200_GEMV_PROBE=1 BNB_B200_GEMV_IMPL=$1 BNB_B200_GEMV_V2W=$2 BNB_B200_GEMV_XREG=$3 BNB_B200_GEMV_NEXTPF=$4 BNB_B200_GEMV_NACC=$5 timeout 300 python tools/kbench.py --only gemv > gpurun_out/r7_kbench_$6.jsonl 2>&1
}
run 2 8 1 0 0 w8_xreg_n2
run 2 8 1 0 1 w8_xreg_n1
run 2 8 0 0 0 w8_nox_n4
run 2 16 1 0 0 w16_xreg_n2
run 2 16 0 0 0 w16_nox_n4
run 2 8 1 1 0 w8_xreg_n2_pf
run 2 8 0 1 0 w8_nox_n4_pf
benchrun() {
  BNB_B200_GEMV_IMPL=$1 BNB_B200_GEMV_V2W=$2 BNB_B200_GEMV_XREG=$3 BNB_B200_GEMV_NEXTPF=$4 BNB_B200_GEMV_NACC=$5 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r7_bench_$6.json 2> gpurun_out/r7_bench_$6.err
}
benchrun 2 8 1 0 0 w8_xreg_n2
benchrun 2 8 0 0 0 w8_nox_n4
benchrun 2 8 1 1 0 w8_xreg_n2_pf
benchrun 2 16 1 1 0 w16_xreg_n2_pf
python tools/kbench.py --only int8 > gpurun_out/r7_kbench_int8.jsonl 2>&1
for f in gpurun_out/r7_kbench_w*.jsonl; do echo "== $f"; python - "$f" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    try: d=json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    print(d['kernel'][24:], d['us'], d.get('hbm_frac'), d.get('cta_us'), {k: round(v,2) for k,v in d.get('phase_us',{}).items()})
PY
done
for f in gpurun_out/r7_bench_*.json; do echo $f; python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['value'], d['roofline']['frac'], d['e2e']['value'], d.get('fused_same_input',{}).get('value'))"; done
grep -o '"kernel": "[^"]*", "us": [0-9.]*' gpurun_out/r7_kbench_int8.jsonl
