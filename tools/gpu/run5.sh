cd $GRAFT_REPO_ROOT
rm -f gpurun_out/ab5.jsonl
run() { tag=$1; shift; env "$@" python tools/ab.py --tag $tag --warmup 1 --iters 1 --top 10 $ABARGS >> gpurun_out/ab5.jsonl 2>> gpurun_out/ab5.err; }
ABARGS="--workload powerlaw --scale 0.2 --k 200"
run p_24 PRIMALCR_UBLOCK_MB=24
run p_100 PRIMALCR_UBLOCK_MB=100
run p_300 PRIMALCR_UBLOCK_MB=300
run p_2000 PRIMALCR_UBLOCK_MB=2000
ABARGS="--workload yahoo --scale 0.4 --k 100"
run y_24 PRIMALCR_UBLOCK_MB=24
run y_100 PRIMALCR_UBLOCK_MB=100
run y_300 PRIMALCR_UBLOCK_MB=300
run y_2000 PRIMALCR_UBLOCK_MB=2000
python - <<'PY'
import json
for l in open('gpurun_out/ab5.jsonl'):
    d=json.loads(l); k=d['kernels']
    print(d['tag'], d['nnz'], round(d['sec_per_iter'],4), 'rs_items',k.get('rowsum_items'),'finalize',k.get('rowsum_finalize'),'dots',k.get('dots'), 'obj', d['objective'][-1])
PY
