# multi-GPU bench exactly as the driver launches it; usage: bash scripts/gpu_multi.sh N
N=${1:-2}
mkdir -p gpurun_out/multi
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/multi/bench_n$N.json 2> gpurun_out/multi/bench_n$N.err
echo "rc=$?"; tail -c 1500 gpurun_out/multi/bench_n$N.json; tail -5 gpurun_out/multi/bench_n$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/multi/ref_n$N.json 2> gpurun_out/multi/ref_n$N.err
echo "rc=$?"; head -c 600 gpurun_out/multi/ref_n$N.json
