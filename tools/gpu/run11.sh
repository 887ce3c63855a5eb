cd $GRAFT_REPO_ROOT
rm -f gpurun_out/ab11.jsonl
run() { tag=$1; shift; env "$@" python tools/ab.py --tag $tag --top 16 $ABARGS >> gpurun_out/ab11.jsonl 2>> gpurun_out/ab11.err; }
ABARGS=""
run base PRIMALCR_VERBOSE_SETUP=1
run lm5 PRIMALCR_LIB=$PWD/primalcr_b200/variants/lib_lm5.so
run lm6 PRIMALCR_LIB=$PWD/primalcr_b200/variants/lib_lm6.so
run pp5 PRIMALCR_LIB=$PWD/primalcr_b200/variants/lib_pp5.so
run pp6 PRIMALCR_LIB=$PWD/primalcr_b200/variants/lib_pp6.so
grep "primalcr setup" gpurun_out/ab11.err | head -20
python - <<'PY'
import json
for l in open('gpurun_out/ab11.jsonl'):
    d=json.loads(l); k=d['kernels']
    print(d['tag'], round(d['sec_per_iter'],4), 'lm_hv',k.get('lm_sweep_hv'),'lm_hv_L',k.get('lm_sweep_hv_L'),'prep',k.get('tile_prepare'),'prep_L',k.get('tile_prepare_L'),'obj',k.get('lm_sweep_obj'))
PY
