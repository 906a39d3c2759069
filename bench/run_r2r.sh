cd $GRAFT_REPO_ROOT
for v in 1 2 4 8; do echo "SUBS=$v"; GI_DEV_SUBS=$v python bench/walk_sweep.py 1.0 2>&1 | head -1; done > gpurun_out/r2r_sweep.log 2>&1
cat gpurun_out/r2r_sweep.log
