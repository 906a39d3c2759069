cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_scattered.py tests/test_gpu_parity.py tests/test_gpu_edge_cases.py tests/test_abi.py -x -q 2>&1 | tail -25 > gpurun_out/r2t_pytest.log
cat gpurun_out/r2t_pytest.log
