cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2g_pytest.log
python bench/walk_sweep.py 1.0 > gpurun_out/r2g_sweep.log 2>&1
python bench/band_probe.py 8 1.0 > gpurun_out/r2g_band.log 2>&1
python bench/hull_probe.py > gpurun_out/r2g_hull.log 2>&1
cat gpurun_out/r2g_pytest.log gpurun_out/r2g_sweep.log gpurun_out/r2g_band.log gpurun_out/r2g_hull.log
