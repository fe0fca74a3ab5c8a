timeout 300 python tools/kbench.py --only gemm4 > gpurun_out/kb_tmp.jsonl 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/kb_tmp.jsonl'):
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    if '_b128' in d['kernel'] or '_b256' in d['kernel']: continue
    print(d['kernel'], d['us'], d.get('hbm_frac'), d.get('speedup_vs_composition'))
PY
