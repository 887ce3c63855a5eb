cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q > gpurun_out/pytest24.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest24.log
rm -f gpurun_out/ab18.jsonl
python tools/ab.py --tag staged --top 12 >> gpurun_out/ab18.jsonl 2>> gpurun_out/ab18.err
python - <<'PY'
import json
for l in open('gpurun_out/ab18.jsonl'):
    d=json.loads(l); k=d['kernels']
    print(d['tag'], d['nnz'], round(d['sec_per_iter'],4), 'obj', d['objective'][-1])
    print({n:v for n,v in k.items() if 'lm_' in n})
PY
