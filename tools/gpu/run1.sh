set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q > gpurun_out/pytest16.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest16.log
run() { tag=$1; shift; env "$@" python tools/ab.py --tag $tag >> gpurun_out/ab1.jsonl 2>> gpurun_out/ab1.err; }
rm -f gpurun_out/ab1.jsonl
run gen1_noreuse PRIMALCR_DOTS_GEN1=1 PRIMALCR_HOT_ROWS=0 PRIMALCR_NO_REUSE=1
run gen1_reuse PRIMALCR_DOTS_GEN1=1 PRIMALCR_HOT_ROWS=0
run gen2_hot0 PRIMALCR_HOT_ROWS=0
run gen2_hot128 PRIMALCR_HOT_ROWS=128
run gen2_hot192 PRIMALCR_HOT_ROWS=192
run gen2_hot240 PRIMALCR_HOT_ROWS=240
run gen2_hot400 PRIMALCR_HOT_ROWS=400
python - <<'PY'
import json
for l in open('gpurun_out/ab1.jsonl'):
    d=json.loads(l); k=d['kernels']
    print(d['tag'], round(d['sec_per_iter'],4), 'dots',k.get('dots'),'dots_active',k.get('dots_active'),'rs_users_act',k.get('rowsum_users_active'),'rs_items',k.get('rowsum_items'), 'obj', d['objective'][-1])
PY
