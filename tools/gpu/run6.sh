cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q > gpurun_out/pytest18.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest18.log
rm -f gpurun_out/ab6.jsonl
run() { tag=$1; shift; env "$@" python tools/ab.py --tag $tag --top 30 $ABARGS >> gpurun_out/ab6.jsonl 2>> gpurun_out/ab6.err; }
ABARGS=""
run n_new X=1
run n_legacy PRIMALCR_HEAVY_LEGACY=1
ABARGS="--workload powerlaw --scale 0.2 --k 200 --warmup 1 --iters 1"
run p_new X=1
run p_legacy PRIMALCR_HEAVY_LEGACY=1
python - <<'PY'
import json
for l in open('gpurun_out/ab6.jsonl'):
    d=json.loads(l); k=d['kernels']
    print(d['tag'], d['nnz'], round(d['sec_per_iter'],4), 'obj', d['objective'][-1], d['counters'])
    print('   ', {n:v for n,v in k.items()})
PY
