# The measurement sweep behind profiles/r02_final_* (one B200): GPU tests, smoke(), the four bench lines, the
# reference arm, ncu launch lists of the same commands, ncu --set full captures of the dominant kernels.
set -x
cd $GRAFT_REPO_ROOT
T=${1:-fin3}
python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/${T}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1
python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_c4.json 2> gpurun_out/${T}_c4.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_c4_ref.json 2> gpurun_out/${T}_c4_ref.err
for w in c2 c3 c5; do python bench.py --workload $w --steps 5 --warmup 3 > gpurun_out/${T}_$w.json 2> gpurun_out/${T}_$w.err; done
python bench/c1_latency.py > gpurun_out/${T}_c1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_c4_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${T}_ncu_c4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:locate_eval_kernel -s 2 -c 1 -o gpurun_out/${T}_locate_eval -f python bench/ncu_case.py 1.0 > gpurun_out/${T}_ncu_le.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pack_mesh_kernel -s 2 -c 1 -o gpurun_out/${T}_pack -f python bench/ncu_case.py 1.0 > gpurun_out/${T}_ncu_pack.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:bilinear_sector_kernel -s 2 -c 1 -o gpurun_out/${T}_bilinear -f python bench/phase_probe.py c5 1.0 > gpurun_out/${T}_ncu_bil.log 2>&1
cat gpurun_out/${T}_pytest.log gpurun_out/${T}_smoke.log gpurun_out/${T}_c1.log
tail -c 600 gpurun_out/${T}_c4.json
PYTHONPATH=. python bench/f32_probe.py > gpurun_out/${T}_f32.log 2>&1
tail -8 gpurun_out/${T}_f32.log
