cd $GRAFT_REPO_ROOT
for v in 0 1; do echo "P2=$v"; GI_DEV_P2=$v python bench/walk_sweep.py 1.0 2>&1 | head -1; done > gpurun_out/r2k_sweep.log 2>&1
cat gpurun_out/r2k_sweep.log
