cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q > gpurun_out/pytest27.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest27.log
rm -f gpurun_out/ab22.jsonl
run() { tag=$1; shift; env "$@" python tools/ab.py --tag $tag --top 40 $ABARGS >> gpurun_out/ab22.jsonl 2>> gpurun_out/ab22.err; }
run tileM X=1
run noM PRIMALCR_NO_TILE_M=1
python - <<'PY'
import json
for l in open('gpurun_out/ab22.jsonl'):
    d=json.loads(l); k=d['kernels']
    print(d['tag'], round(d['sec_per_iter'],4), 'obj', d['objective'][-1])
    print({n:v for n,v in k.items() if 'lm_' in n or 'prepare' in n})
PY
