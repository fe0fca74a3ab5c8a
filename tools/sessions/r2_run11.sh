#!/bin/bash
# GPU session 11: heavy-first tile ranges + table build between the first item's loads and the rest of the ring
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r11_tests.log 2>&1; tail -3 gpurun_out/r11_tests.log
BNB_B200_GEMV_PROBE=1 timeout 300 python tools/kbench.py --only gemv > gpurun_out/r11_kbench.jsonl 2>&1
for shp in "4096 4096" "11008 4096" "28672 8192"; do
  set -- $shp
  BNB_B200_GEMV_PROBE=2 timeout 120 python tools/gemv_trace.py $1 $2 > gpurun_out/r11_trace_$1x$2.json 2> gpurun_out/r11_trace_$1x$2.err
done
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r11_bench.json 2> gpurun_out/r11_bench.err
python - <<'PY'
import json
for l in open('gpurun_out/r11_kbench.jsonl'):
    try: d=json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    print(d['kernel'][24:], d['us'], d.get('hbm_frac'), d.get('cta_us'), {k: round(v,2) for k,v in d.get('phase_us',{}).items()})
for shp in ("4096x4096","11008x4096","28672x8192"):
    try:
        d=json.load(open(f'gpurun_out/r11_trace_{shp}.json'))
        print(shp, 'period', d['period_us'], 'handoff', d['handoff_lastexitA_to_first_prevdoneB_us'], 'A.last', d['A']['last_cta'], 'A.exit', d['A']['exit'], 'A.x_ready', d['A']['x_ready'], 'A.prev_done', d['A']['prev_done'], 'A.entry', d['A']['entry'])
    except Exception as e: print(shp, e)
d=json.loads(open('gpurun_out/r11_bench.json').read().strip().splitlines()[-1])
print('value', d['value'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'], 'fused', d.get('fused_same_input',{}).get('value'), '70b', d.get('llama3_70b',{}).get('value'), d.get('llama3_70b',{}).get('tok_per_s'))
PY
