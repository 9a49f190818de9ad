O=gpurun_out/r01f3; mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 4 --warmup 3 > $O/bench_2gpu.log 2> $O/bench_2gpu.err; echo "bench2 rc=$?" >> $O/summary.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29552 tools/multi_gpu_selfplay.py --games 64 > $O/multi.log 2> $O/multi.err; echo "multi rc=$?" >> $O/summary.txt
