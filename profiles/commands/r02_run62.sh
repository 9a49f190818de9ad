O=gpurun_out/r02bj; mkdir -p $O
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_2gpu.json 2> $O/bench_2gpu.err; echo "bench2 rc=$?" >> $O/summary.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29552 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > $O/bench_ref_2gpu.json 2> $O/bench_ref_2gpu.err; echo "ref2 rc=$?" >> $O/summary.txt
cat $O/summary.txt; tail -c 600 $O/bench_2gpu.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r02bj/bench_2gpu.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value']); print(d['phases_ms_per_step']); print(d['per_rank'])
print(d['config4']['value'], [(p['leaf_batch_per_gpu'], round(p['value'])) for p in d['config5']['points']], d['config3_leaf_batch_256']['value'])
P
tail -c 400 $O/bench_ref_2gpu.json
