#!/bin/bash
# GPU session 42: final library -- whole GPU suite, smoke, default bench line
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r42_tests.log 2>&1; tail -3 gpurun_out/r42_tests.log
timeout 600 python __graft_entry__.py --smoke > gpurun_out/r42_smoke.log 2>&1; tail -2 gpurun_out/r42_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r42_bench.json 2> gpurun_out/r42_bench.err; tail -c 300 gpurun_out/r42_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r42_bench.json').read().strip().splitlines()[-1])
print('value', d['value'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'], 'fused', d.get('fused_same_input',{}).get('value'))
print('70b', {k:v for k,v in d.get('llama3_70b',{}).items() if k in ('value','tok_per_s','ms_per_step','error')})
for k,v in d.get('igemmlt',{}).items():
    if isinstance(v,dict): print(k, round(v['TOPS'],1), round(v['roofline']['frac_of_nominal_int8'],3), round(v['e2e']['value'],1))
for k,v in d.get('gemm_4bit',{}).items():
    if isinstance(v,dict): print(k, round(v['us'],2), round(v['reference_composition_us'],2), round(v['TFLOPs'],1))
print('cpu', d.get('cpu_baseline',{}).get('value'), 'clocks', d.get('clocks'))
PY
