cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q > gpurun_out/pytest31.log 2>&1; echo "pytest rc=$?"
tail -2 gpurun_out/pytest31.log
rm -f gpurun_out/trace2.txt gpurun_out/ab31.jsonl
PRIMALCR_TRACE=$PWD/gpurun_out/trace2.txt python tools/ab.py --tag final --warmup 2 --iters 1 --top 40 > gpurun_out/ab31.jsonl 2> gpurun_out/ab31.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/ab31.jsonl').read().strip().splitlines()[-1]); k=d['kernels']
print(d['tag'], round(d['sec_per_iter'],4), 'obj', d['objective'][-1], {n:v for n,v in k.items() if n.startswith('u_')})
PY
