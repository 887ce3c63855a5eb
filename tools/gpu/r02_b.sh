cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r02b_gpu.txt
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02b_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r02b_pytest.log
PRIMALCR_VERBOSE_SETUP=1 timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r02b_bench1.json 2> gpurun_out/r02b_bench1.err; echo "bench1 rc=$?"
grep -i "shim\|host\|setup" gpurun_out/r02b_bench1.err | tail -30
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02b_bench2.json 2> gpurun_out/r02b_bench2.err; echo "bench2 rc=$?"
tail -c 400 gpurun_out/r02b_bench2.err
PRIMALCR_AR_GROUPS=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02b_bench2_g1.json 2> gpurun_out/r02b_bench2_g1.err; echo "bench2 g1 rc=$?"
