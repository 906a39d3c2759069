cd $GRAFT_REPO_ROOT
ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 24 --csv --log-file gpurun_out/r2p_launches.csv python bench/ncu_case.py 1.0 > gpurun_out/r2p_ncu2.log 2>&1
tail -2 gpurun_out/r2p_ncu2.log
