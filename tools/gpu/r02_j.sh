cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
export PRIMALCR_SYNTH_SORTED_ITEMS=1
for hot in 0 48 80; do
  PRIMALCR_L2_HOT_MB=$hot timeout 400 python tools/ab.py --workload yahoo --scale 0.4 --warmup 1 --iters 2 --tag hot$hot >> gpurun_out/r02j_ab.jsonl 2>> gpurun_out/r02j_ab.err; echo "ab hot$hot rc=$?"
done
grep "persisting" gpurun_out/r02j_ab.err | head -3
python - <<'PY'
import json
for l in open('gpurun_out/r02j_ab.jsonl'):
    d=json.loads(l); print(d['tag'], round(d['sec_per_iter'],4), d['objective'][-1], {k:v[0] for k,v in d['kernels'].items() if k in ('dots','dots_active','rowsum_users_active','rowsum_items','rowsum_users')})
PY
