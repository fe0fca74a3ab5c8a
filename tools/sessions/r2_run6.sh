#!/bin/bash
# GPU session 6: absmax requested ahead of the weights; one-CTA-per-SM-per-kernel co-residency (PERSM=1); full GPU test suite
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r6_tests.log 2>&1; tail -3 gpurun_out/r6_tests.log
run() { # impl warps xreg pf persm v2max tag
  BNB_B200_GEMV_PROBE=1 BNB_B200_GEMV_IMPL=$1 BNB_B200_GEMV_V2W=$2 BNB_B200_GEMV_XREG=$3 BNB_B200_GEMV_NEXTPF=$4 BNB_B200_GEMV_PERSM=$5 BNB_B200_GEMV_V2MAX=$6 timeout 300 python tools/kbench.py --only gemv > gpurun_out/r6_kbench_$7.jsonl 2>&1
}
run 2 8 1 0 0 3 v2w8
run 2 8 1 1 0 3 v2w8_pf
run 2 16 1 0 0 3 v2w16
run 2 8 0 0 1 8 v2w8_persm
run 2 8 0 1 1 8 v2w8_persm_pf
run 2 8 2 0 1 8 v2w8_persm_xreg
run b 0 0 0 0 3 bc
benchrun() {
  BNB_B200_GEMV_IMPL=$1 BNB_B200_GEMV_V2W=$2 BNB_B200_GEMV_XREG=$3 BNB_B200_GEMV_NEXTPF=$4 BNB_B200_GEMV_PERSM=$5 BNB_B200_GEMV_V2MAX=$6 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r6_bench_$7.json 2> gpurun_out/r6_bench_$7.err
}
benchrun 2 8 1 0 0 3 v2w8
benchrun 2 8 1 1 0 3 v2w8_pf
benchrun 2 8 0 0 1 8 v2w8_persm
benchrun 2 8 0 1 1 8 v2w8_persm_pf
benchrun 2 8 2 1 1 8 v2w8_persm_xreg_pf
for f in gpurun_out/r6_kbench_*.jsonl; do echo "== $f"; python - "$f" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    try: d=json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    print(d['kernel'][24:], d['us'], d.get('hbm_frac'), d.get('cta_us'), {k: round(v,2) for k,v in d.get('phase_us',{}).items()})
PY
done
for f in gpurun_out/r6_bench_*.json; do echo $f; python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['value'], d['roofline']['frac'], d['e2e']['value'], d.get('fused_same_input',{}).get('value'))"; done
