O=gpurun_out/r02bl; mkdir -p $O
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 8 --steps 20 --warmup 5 --config5-steps 5 > $O/bench_8gpu.json 2> $O/bench_8gpu.err; echo "bench8 rc=$?" >> $O/summary.txt
cat $O/summary.txt; tail -c 300 $O/bench_8gpu.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r02bl/bench_8gpu.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value']); print(d['phases_ms_per_step']); print(d['per_rank'])
print(d['config4']['value'], [(p['leaf_batch_per_gpu'], round(p['value'])) for p in d['config5']['points']], d['config3_leaf_batch_256']['value'])
P
