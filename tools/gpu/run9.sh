cd $GRAFT_REPO_ROOT
rm -f gpurun_out/ab9.jsonl
run() { tag=$1; shift; env "$@" python tools/ab.py --tag $tag --top 8 $ABARGS >> gpurun_out/ab9.jsonl 2>> gpurun_out/ab9.err; }
ABARGS=""
run base X=1
run rs256_1 PRIMALCR_RS256=1
run rs256_2 PRIMALCR_RS256=2
run dots256 PRIMALCR_DOTS256=1
run both PRIMALCR_DOTS256=1 PRIMALCR_RS256=2
python - <<'PY'
import json
for l in open('gpurun_out/ab9.jsonl'):
    d=json.loads(l); k=d['kernels']
    print(d['tag'], round(d['sec_per_iter'],4), 'rs_items',k.get('rowsum_items'),'dots',k.get('dots'),'dots_active',k.get('dots_active'),'rs_users_act',k.get('rowsum_users_active'), 'obj', d['objective'][-1])
PY
PRIMALCR_DOTS256=1 PRIMALCR_RS256=2 python -m pytest tests -m gpu -x -q > gpurun_out/pytest20.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest20.log
