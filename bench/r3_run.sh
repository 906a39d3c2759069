# round-2 session-4 sweep: GPU tests, C4 bench with the plane block, ncu of the plane kernel and the hinted pack pass
set -x
cd $GRAFT_REPO_ROOT
T=${1:-s4a}
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${T}_pytest.log
python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_c4.json 2> gpurun_out/${T}_c4.err
ncu --set full --clock-control none --import-source on -k regex:locate_eval_kernel -s 2 -c 1 -o gpurun_out/${T}_locate_eval_plane -f python bench/ncu_case.py 1.0 0 1 > gpurun_out/${T}_ncu_lep.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pack_mesh_kernel -s 2 -c 1 -o gpurun_out/${T}_pack -f python bench/ncu_case.py 1.0 > gpurun_out/${T}_ncu_pack.log 2>&1
cat gpurun_out/${T}_pytest.log
tail -c 1500 gpurun_out/${T}_c4.json
