set -x
cd $GRAFT_REPO_ROOT
nvidia-smi -L
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2a_pytest.log
python bench/walk_sweep.py 1.0 > gpurun_out/r2a_sweep.log 2>&1
python bench/band_probe.py 8 1.0 > gpurun_out/r2a_band.log 2>&1
python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_c4.json 2> gpurun_out/r2a_c4.err
tail -3 gpurun_out/r2a_pytest.log; cat gpurun_out/r2a_sweep.log gpurun_out/r2a_band.log; tail -c 3000 gpurun_out/r2a_c4.json; tail -5 gpurun_out/r2a_c4.err
