cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02g_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02g_pytest.log; grep -n "^E  " gpurun_out/r02g_pytest.log | head -5
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02g_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02g_smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02g_bench.json 2> gpurun_out/r02g_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02g_bench_reference.json 2> gpurun_out/r02g_bench_reference.err; echo "bench ref rc=$?"; cut -c1-400 gpurun_out/r02g_bench_reference.json
timeout 300 python tools/eval_probe.py --workload netflix > gpurun_out/r02g_eval_netflix.json 2> gpurun_out/r02g_eval.err; echo "eval netflix rc=$?"
timeout 600 python tools/eval_probe.py --workload powerlaw --scale 0.05 --k 200 > gpurun_out/r02g_eval_powerlaw005.json 2>> gpurun_out/r02g_eval.err; echo "eval powerlaw rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02g_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value'],'shim',d['e2e']['shim'].get('value'))
ks={k["name"]:round(k["ms_per_step"],2) for k in d["roofline"]["kernels"]}
print({n:ks.get(n) for n in ("u_finalize_cg","u_cg_step","rowsum_finalize","rowsum_items","rowsum_users_active")})
for f in ('netflix','powerlaw005'):
    e=json.loads(open('gpurun_out/r02g_eval_%s.json'%f).read().strip().splitlines()[-1])
    print(f, 'iter', round(e['outer_iteration_sec'],4), {k:(round(v['sec'],4), v['result']) for k,v in e.items() if isinstance(v,dict)})
PY
