# N-GPU strong-scaling line with the tuned gather bands: bash bench/r3_scale.sh TAG N [extra bench args]
set -x
cd $GRAFT_REPO_ROOT
T=${1:-s4s}; N=${2:-8}; shift; shift
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 "$@" > gpurun_out/${T}_scale$N.json 2> gpurun_out/${T}_scale$N.err
tail -c 3000 gpurun_out/${T}_scale$N.json
tail -5 gpurun_out/${T}_scale$N.err
