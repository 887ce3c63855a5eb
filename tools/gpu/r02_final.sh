cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02final_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02final_pytest.log; grep -n "^E  " gpurun_out/r02final_pytest.log | head -5
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02final_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02final_smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02final_bench.json 2> gpurun_out/r02final_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02final_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value'],'same',d['e2e']['device_same_window'],'shim',d['e2e']['shim'].get('value'),'parity',d['parity']['ok'],d['parity']['obj_rel_err'])
print('roofline', {k:v for k,v in d['roofline'].items() if k in ('kernel','achieved','frac','traffic','dram_frac','l2')})
PY
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"tile_lm_sweep_kernel|tile_prepare_kernel|u_finalize_cg_kernel|eval_sorted_kernel|hs_merge_kernel|hs_chunk_sort_kernel" -c 10 -o gpurun_out/r02_small python tools/profile_step.py --scale 1.0 > gpurun_out/r02final_small.log 2>&1; echo "ncu small rc=$?"
