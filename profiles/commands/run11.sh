set -x
O=gpurun_out/r01k; mkdir -p $O
nvidia-smi -L > $O/gpus.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > $O/bench2.log 2> $O/bench2.err; echo "bench2 rc=$?" >> $O/summary.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > $O/bench2_ref.log 2> $O/bench2_ref.err; echo "bench2 ref rc=$?" >> $O/summary.txt
timeout 300 python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline > $O/bench1.log 2> $O/bench1.err; echo "bench1 rc=$?" >> $O/summary.txt
