cd $GRAFT_REPO_ROOT
run() { tag=$1; shift; env "$@" python tools/ab.py --tag $tag >> gpurun_out/ab3.jsonl 2>> gpurun_out/ab3.err; }
rm -f gpurun_out/ab3.jsonl
run hot0 PRIMALCR_HOT_ROWS=0
run hot16 PRIMALCR_HOT_ROWS=16
run hot64 PRIMALCR_HOT_ROWS=64
run hot80 PRIMALCR_HOT_ROWS=80
python - <<'PY'
import json
for l in open('gpurun_out/ab3.jsonl'):
    d=json.loads(l); k=d['kernels']
    print(d['tag'], round(d['sec_per_iter'],4), 'dots',k.get('dots'),'dots_active',k.get('dots_active'),'rs_users_act',k.get('rowsum_users_active'),'rs_users',k.get('rowsum_users'), 'obj', d['objective'][-1])
PY
