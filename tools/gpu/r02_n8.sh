cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv | sort | uniq -c > gpurun_out/r02n8_gpu.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29601 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_bench_netflix_k100_8gpu.json 2> gpurun_out/r02n8_netflix.err; echo "netflix rc=$?"
timeout 900 $TR --master-port 29602 bench.py --gpus 8 --workload yahoo --steps 5 --warmup 3 > gpurun_out/r02_bench_yahoo_k100_8gpu.json 2> gpurun_out/r02n8_yahoo.err; echo "yahoo rc=$?"
timeout 1200 $TR --master-port 29603 bench.py --gpus 8 --workload powerlaw --k 200 --steps 3 --warmup 2 > gpurun_out/r02_bench_powerlaw_k200_8gpu.json 2> gpurun_out/r02n8_powerlaw.err; echo "powerlaw rc=$?"
timeout 900 python -m pytest tests/test_gpu_scale.py tests/test_gpu_multi.py -m gpu -q -k "sharded" > gpurun_out/r02n8_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02n8_pytest.log
python tools/dump_csr.py --workload netflix --out /dev/shm/csr.bin > gpurun_out/r02n8_dump.log 2>&1
for g in 8 1; do
  PRIMALCR_GPUS=$g PRIMALCR_VERBOSE_SETUP=1 timeout 300 oracle/_ref/shim-e2e /dev/shm/csr.bin 100 5000 10 > gpurun_out/r02n8_shim_${g}gpu.log 2>&1; echo "shim $g rc=$?"; grep SHIM_E2E gpurun_out/r02n8_shim_${g}gpu.log
done
rm -f /dev/shm/csr.bin
for f in netflix yahoo powerlaw; do tail -c 300 gpurun_out/r02n8_$f.err; echo; done
