cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
# (1) the bench command itself, clean, then its ncu launch list (per-launch times only; cold-cache, serialised)
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_netflix_k100_1gpu.json 2> gpurun_out/r02cap_bench.err; echo "bench rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_ncu_launches_bench_netflix_k100.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-shim-e2e > gpurun_out/r02cap_bench_under_ncu.json 2> gpurun_out/r02cap_bench_under_ncu.err; echo "ncu launches rc=$?"
# (2) one `--set full` launch of each hot kernel at full size
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"rowsum_kernel|dots_units_kernel|tile_lm_sweep_kernel|tile_prepare_kernel" -s 6 -c 8 -o gpurun_out/r02_top python tools/profile_step.py --scale 1.0 > gpurun_out/r02cap_top.log 2>&1; echo "ncu full rc=$?"
# (3) DRAM bytes of every launch of one outer iteration
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02_ncu_iteration_dram.csv python tools/profile_step.py --scale 1.0 > gpurun_out/r02cap_iter.log 2>&1; echo "ncu iter rc=$?"
# (4) full-size CLI run: wall-time split load / solve / write through our own host CLI
python tools/dump_csr.py --workload netflix --dir /dev/shm/nf > gpurun_out/r02cap_dump.log 2>&1
( cd /dev/shm && PRIMALCR_VERBOSE_SETUP=1 /usr/bin/time -v $GRAFT_REPO_ROOT/primalcr_b200/bin/primalcr-train -s 2 -k 100 -l 5000 -t 10 -p 0 -n 16 nf nf.model ) > gpurun_out/r02_cli_full_size.log 2>&1; echo "cli rc=$?"
ls -la /dev/shm/nf /dev/shm/nf.model >> gpurun_out/r02_cli_full_size.log 2>&1
rm -rf /dev/shm/nf /dev/shm/nf.model /dev/shm/U.txt /dev/shm/V.txt
tail -20 gpurun_out/r02_cli_full_size.log
# (5) cp.async row-sum variant A/B (stage level) and the power-law scale fixture
for cfg in "base:" "cpasync:PRIMALCR_ROWSUM_CPASYNC=1"; do
  tag=${cfg%%:*}; envs=${cfg#*:}
  env $envs timeout 300 python tools/stage_bench.py --side VU --tag $tag >> gpurun_out/r02cap_stage.jsonl 2>> gpurun_out/r02cap_stage.err; echo "stage $tag rc=$?"
done
cat gpurun_out/r02cap_stage.jsonl
timeout 900 python -m pytest tests/test_gpu_scale.py -m gpu -q -k "powerlaw001" > gpurun_out/r02cap_pytest_powerlaw.log 2>&1; echo "pytest powerlaw rc=$?"; tail -5 gpurun_out/r02cap_pytest_powerlaw.log
