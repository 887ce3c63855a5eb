cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02h_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02h_pytest.log; grep -n "^E  " gpurun_out/r02h_pytest.log | head -5
timeout 300 python tools/eval_probe.py --workload netflix > gpurun_out/r02h_eval_netflix.json 2> gpurun_out/r02h_eval.err; echo "eval netflix rc=$?"
timeout 600 python tools/eval_probe.py --workload powerlaw --scale 0.05 --k 200 > gpurun_out/r02h_eval_powerlaw005.json 2>> gpurun_out/r02h_eval.err; echo "eval powerlaw rc=$?"
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"rowsum_kernel|dots_units_kernel|tile_lm_sweep_kernel<1, 5, 256>|tile_prepare_kernel<5, 256>" -s 6 -c 6 -o gpurun_out/r02_top python tools/profile_step.py --scale 1.0 > gpurun_out/r02h_top.log 2>&1; echo "ncu full rc=$?"
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02_ncu_iteration_dram.csv python tools/profile_step.py --scale 1.0 > gpurun_out/r02h_iter.log 2>&1; echo "ncu iter rc=$?"
python - <<'PY'
import json
for f in ('netflix','powerlaw005'):
    e=json.loads(open('gpurun_out/r02h_eval_%s.json'%f).read().strip().splitlines()[-1])
    print(f, 'iter', round(e['outer_iteration_sec'],4), {k:(round(v['sec']*1e3,2), v['result'], v['kernels']) for k,v in e.items() if isinstance(v,dict)})
PY
