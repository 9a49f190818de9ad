O=gpurun_out/r01m8; mkdir -p $O
nvidia-smi -L > $O/gpus.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 4 --warmup 3 > $O/bench_8gpu.log 2> $O/bench_8gpu.err; echo "bench8 rc=$?" >> $O/summary.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 8 --games 1024 --steps 4 --warmup 3 --no-cpu-baseline --midgame-ply 0 > $O/bench_8gpu_8192.log 2> $O/bench_8gpu_8192.err; echo "bench8x1024 rc=$?" >> $O/summary.txt
