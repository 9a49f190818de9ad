set -x
O=gpurun_out/r01m4; mkdir -p $O
nvidia-smi -L > $O/gpus.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 4 --steps 4 --warmup 3 > $O/bench_4gpu.log 2> $O/bench_4gpu.err; echo "bench4 rc=$?" >> $O/summary.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 4 --games 2048 --steps 2 --warmup 3 --no-cpu-baseline --midgame-ply 0 > $O/bench_4gpu_8192.log 2> $O/bench_4gpu_8192.err; echo "bench4x2048 rc=$?" >> $O/summary.txt
