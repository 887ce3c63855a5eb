cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q > gpurun_out/pytest22.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest22.log
rm -f gpurun_out/ab12.jsonl
run() { tag=$1; shift; env "$@" python tools/ab.py --tag $tag --top 16 $ABARGS >> gpurun_out/ab12.jsonl 2>> gpurun_out/ab12.err; }
ABARGS=""
run regsort X=1
python - <<'PY'
import json
for l in open('gpurun_out/ab12.jsonl'):
    d=json.loads(l); k=d['kernels']
    print(d['tag'], d['nnz'], round(d['sec_per_iter'],4), 'obj', d['objective'][-1], 'kernel_ms_total', d['kernel_ms_total'])
    print('   ', {n:v for n,v in list(k.items())})
PY
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_e2e.json 2> gpurun_out/bench_e2e.err; grep "e2e phases" gpurun_out/bench_e2e.err
