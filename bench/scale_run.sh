# Strong-scaling lines of the literal BASELINE config at N = 2, 4, 8 (one rank per GPU) + the reference arm under torchrun.
cd $GRAFT_REPO_ROOT
export GI_CACHE=$GRAFT_REPO_ROOT/.gi_cache
T=${1:-fin2}
nvidia-smi -L | wc -l
for n in 8 4 2; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29520 + n)) bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/${T}_scale$n.json 2> gpurun_out/${T}_scale$n.err
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29540 bench.py --impl reference --gpus 8 --steps 2 --warmup 1 > gpurun_out/${T}_ref8.json 2> gpurun_out/${T}_ref8.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --workload c5 --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_c5_scale8.json 2> gpurun_out/${T}_c5_scale8.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 bench.py --workload c3 --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_c3_scale8.json 2> gpurun_out/${T}_c3_scale8.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29543 bench.py --workload c2 --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_c2_scale8.json 2> gpurun_out/${T}_c2_scale8.err
for f in scale8 scale4 scale2 ref8 c5_scale8 c3_scale8 c2_scale8; do echo "== $f"; tail -c 400 gpurun_out/${T}_$f.json; tail -2 gpurun_out/${T}_$f.err | cut -c1-300; done
