set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -60 > gpurun_out/r2d_pytest.log
for m in 6 7 8 9 10; do echo MINB $m; GI_WALK_MINB=$m python bench/walk_sweep.py 1.0 2>&1 | head -1; done > gpurun_out/r2d_sweep.log 2>&1
python bench/band_probe.py 8 1.0 > gpurun_out/r2d_band.log 2>&1
cat gpurun_out/r2d_pytest.log gpurun_out/r2d_sweep.log gpurun_out/r2d_band.log
