cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for tag in base rs5 rs6 dt4 ch512 ch128; do
  if [ "$tag" = "base" ]; then lib=""; else lib="PRIMALCR_LIB=$GRAFT_REPO_ROOT/primalcr_b200/variants/lib_$tag.so"; fi
  env $lib timeout 300 python tools/stage_bench.py --side VU --reps 5 --tag $tag >> gpurun_out/r02i_stage.jsonl 2>> gpurun_out/r02i_stage.err; echo "stage $tag rc=$?"
done
python - <<'PY'
import json
for l in open('gpurun_out/r02i_stage.jsonl'):
    d=json.loads(l); print(d['tag'], {k:v[0] for k,v in d['ms_per_launch'].items() if k in ('rowsum_items','rowsum_users','dots','rowsum_finalize')})
PY
