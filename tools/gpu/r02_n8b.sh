cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29601 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_bench_netflix_k100_8gpu.json 2> gpurun_out/r02n8b_netflix.err; echo "netflix rc=$?"
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL,TUNING timeout 600 $TR --master-port 29602 bench.py --gpus 8 --steps 2 --warmup 3 > gpurun_out/r02n8b_netflix_dbg.json 2> gpurun_out/r02n8b_netflix_dbg.err; echo "netflix dbg rc=$?"
grep -i "nvls\|algo\|Channel\|Connected" gpurun_out/r02n8b_netflix_dbg.err | head -40 > gpurun_out/r02n8b_nccl_info.txt; wc -l gpurun_out/r02n8b_nccl_info.txt
PRIMALCR_SHARDED_CG=1 timeout 600 $TR --master-port 29603 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02n8b_netflix_rs.json 2> gpurun_out/r02n8b_netflix_rs.err; echo "netflix rs rc=$?"
timeout 900 python -m pytest tests/test_gpu_scale.py tests/test_gpu_multi.py -m gpu -q -k "sharded" > gpurun_out/r02n8b_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02n8b_pytest.log
python - <<'PY'
import json
for f in ("r02_bench_netflix_k100_8gpu","r02n8b_netflix_rs"):
    d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
    ks={k["name"]:round(k["ms_per_step"],2) for k in d["roofline"]["kernels"]}
    print(f, d["value"], {n:ks.get(n) for n in ("nccl_allreduce","nccl_reduce_scatter","nccl_all_gather","cg_update","axpby","rowsum_finalize")})
PY
head -20 gpurun_out/r02n8b_nccl_info.txt | cut -c1-200
