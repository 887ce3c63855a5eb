cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q > gpurun_out/pytest19.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest19.log
rm -f gpurun_out/ab7.jsonl
run() { tag=$1; shift; env "$@" python tools/ab.py --tag $tag --top 30 $ABARGS >> gpurun_out/ab7.jsonl 2>> gpurun_out/ab7.err; }
ABARGS=""
run n_aux X=1
run n_noaux PRIMALCR_NO_AUX_STREAM=1
ABARGS="--workload powerlaw --scale 0.2 --k 200 --warmup 1 --iters 1"
run p_aux X=1
python - <<'PY'
import json
for l in open('gpurun_out/ab7.jsonl'):
    d=json.loads(l); k=d['kernels']
    print(d['tag'], d['nnz'], round(d['sec_per_iter'],4), 'obj', d['objective'][-1], 'kernel_ms_total', d['kernel_ms_total'])
    print('   ', {n:v for n,v in list(k.items())[:14]})
PY
