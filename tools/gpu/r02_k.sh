cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 400 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02k_bench_reference.json 2> gpurun_out/r02k_bench_reference.err; echo "ref rc=$?"; tail -2 gpurun_out/r02k_bench_reference.err | cut -c1-600
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02k_bench.json 2> gpurun_out/r02k_bench.err; echo "bench rc=$?"; grep "cpu reference" gpurun_out/r02k_bench.err | cut -c1-600
python - <<'PY'
import json
r=json.loads(open('gpurun_out/r02k_bench_reference.json').read().strip().splitlines()[-1]); print('ref value', r['value'], r['extrapolation']['linear_value'], r['extrapolation']['fit'])
d=json.loads(open('gpurun_out/r02k_bench.json').read().strip().splitlines()[-1]); print('ours', d['value'], d['e2e']['value'], d['cpu_baseline']['value'], d['cpu_baseline']['linear_value'], d['cpu_baseline']['fit'])
PY
