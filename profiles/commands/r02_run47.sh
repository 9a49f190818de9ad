O=gpurun_out/r02au; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_resnet.py tests/test_gpu_search.py -m gpu -x -q > $O/pytest_2gpu.log 2>&1; echo "pytest rc=$? $(tail -1 $O/pytest_2gpu.log)" >> $O/summary.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/multi_gpu_selfplay.py > $O/multi_gpu_check.json 2> $O/multi.err; echo "multi rc=$?" >> $O/summary.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_2gpu.json 2> $O/bench_2gpu.err; echo "bench2 rc=$?" >> $O/summary.txt
cat $O/summary.txt; tail -1 $O/multi_gpu_check.json | cut -c1-400
