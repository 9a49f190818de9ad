O=gpurun_out/r02an; mkdir -p $O
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_8gpu.json 2> $O/bench_8gpu.err; echo "bench8 rc=$?" >> $O/summary.txt
tail -c 3000 $O/bench_8gpu.json; tail -3 $O/bench_8gpu.err
