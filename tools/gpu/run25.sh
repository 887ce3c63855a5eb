cd $GRAFT_REPO_ROOT
rm -f gpurun_out/ab25.jsonl
run() { tag=$1; shift; env "$@" python tools/ab.py --tag $tag --top 14 $ABARGS >> gpurun_out/ab25.jsonl 2>> gpurun_out/ab25.err; }
run c256 X=1
run c128 PRIMALCR_LIB=$PWD/primalcr_b200/variants/lib_chunk128.so
run c256b X=1
run c128b PRIMALCR_LIB=$PWD/primalcr_b200/variants/lib_chunk128.so
python - <<'PY'
import json
for l in open('gpurun_out/ab25.jsonl'):
    d=json.loads(l); k=d['kernels']
    print(d['tag'], round(d['sec_per_iter'],4), {n:v[0] for n,v in k.items() if n in ('dots','dots_active','rowsum_items','rowsum_users_active','rowsum_users','rowsum_finalize')})
PY
