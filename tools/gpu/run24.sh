cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest28.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest28.log
rm -f gpurun_out/ab24.jsonl
python tools/ab.py --tag cmp1 --top 40 >> gpurun_out/ab24.jsonl 2>> gpurun_out/ab24.err
python - <<'PY'
import json
for l in open('gpurun_out/ab24.jsonl'):
    d=json.loads(l); k=d['kernels']
    print(d['tag'], round(d['sec_per_iter'],4), 'obj', d['objective'][-1])
    print({n:v for n,v in k.items() if 'prepare' in n})
PY
