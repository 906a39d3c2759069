cd $GRAFT_REPO_ROOT
python bench/hull_probe.py > gpurun_out/r2f_hull.log 2>&1
cat gpurun_out/r2f_hull.log
