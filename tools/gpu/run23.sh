cd $GRAFT_REPO_ROOT
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"tile_prepare_kernel" -c 1 -o gpurun_out/r01g_prep python tools/profile_step.py --scale 1.0 > gpurun_out/r01g_prep.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/r01g_prep.ncu-rep
