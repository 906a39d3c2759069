cd $GRAFT_REPO_ROOT
python bench/walk_sweep.py 1.0 > gpurun_out/r2h_sweep.log 2>&1
python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge_cases.py tests/test_gpu_fuzz.py -x -q 2>&1 | tail -3 > gpurun_out/r2h_pytest.log
cat gpurun_out/r2h_sweep.log gpurun_out/r2h_pytest.log
