cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_scale.py tests/test_gpu_multi.py -m gpu -q -k "sharded" > gpurun_out/r02d_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02d_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02d_netflix2_ar.json 2> gpurun_out/r02d_netflix2_ar.err; echo "netflix ar rc=$?"
PRIMALCR_SHARDED_CG=1 timeout 600 $TR --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02d_netflix2_rs.json 2> gpurun_out/r02d_netflix2_rs.err; echo "netflix rs rc=$?"
PRIMALCR_SHARDED_CG=0 timeout 600 $TR --master-port 29513 bench.py --gpus 2 --workload yahoo --scale 0.4 --steps 3 --warmup 2 > gpurun_out/r02d_yahoo04_2_ar.json 2> gpurun_out/r02d_yahoo04_2_ar.err; echo "yahoo ar rc=$?"
timeout 600 $TR --master-port 29514 bench.py --gpus 2 --workload yahoo --scale 0.4 --steps 3 --warmup 2 > gpurun_out/r02d_yahoo04_2_rs.json 2> gpurun_out/r02d_yahoo04_2_rs.err; echo "yahoo rs rc=$?"
python - <<'PY'
import json
for f in ("netflix2_ar","netflix2_rs","yahoo04_2_ar","yahoo04_2_rs"):
    try:
        d=json.loads(open("gpurun_out/r02d_%s.json"%f).read().strip().splitlines()[-1])
        ks={k["name"]:round(k["ms_per_step"],2) for k in d["roofline"]["kernels"]}
        print(f, d["value"], d["objective"][-1], {n:ks.get(n) for n in ("nccl_allreduce","nccl_reduce_scatter","nccl_all_gather","cg_update","cg_dots2","axpby","rowsum_items","dots")})
    except Exception as ex: print(f, "ERR", ex)
PY
