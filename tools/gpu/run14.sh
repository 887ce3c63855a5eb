cd $GRAFT_REPO_ROOT
rm -f gpurun_out/ab14.jsonl
run() { tag=$1; shift; env "$@" python tools/ab.py --tag $tag --warmup 1 --iters 1 --top 10 $ABARGS >> gpurun_out/ab14.jsonl 2>> gpurun_out/ab14.err; }
ABARGS="--workload yahoo --scale 0.4 --k 100"
run y_off PRIMALCR_L2_HOT_MB=0
run y_32 PRIMALCR_L2_HOT_MB=32
run y_64 PRIMALCR_L2_HOT_MB=64
run y_96 PRIMALCR_L2_HOT_MB=96
ABARGS="--workload powerlaw --scale 0.2 --k 200"
run p_off PRIMALCR_L2_HOT_MB=0
run p_64 PRIMALCR_L2_HOT_MB=64
python - <<'PY'
import json
for l in open('gpurun_out/ab14.jsonl'):
    d=json.loads(l); k=d['kernels']
    print(d['tag'], d['nnz'], round(d['sec_per_iter'],4), 'dots',k.get('dots'),'dots_active',k.get('dots_active'),'rs_users_act',k.get('rowsum_users_active'),'rs_items',k.get('rowsum_items'), 'obj', d['objective'][-1])
PY
