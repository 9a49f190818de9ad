O=gpurun_out/r02z; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_search.py tests/test_gpu_training.py -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/summary.txt
for b in 256 4096; do timeout 600 python tools/train_bench.py --batch $b --steps 30 --skip-pytorch >> $O/train_bench.jsonl 2>> $O/err.log; done
for i in 1 2; do timeout 300 python bench.py --no-extra-configs --no-cpu-baseline --midgame-ply 0 --steps 10 > $O/bench_$i.log 2>> $O/err.log; python - $O/bench_$i.log >> $O/table.txt <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(round(d['value']), round(d['ms_per_step'],2), round(d['roofline']['achieved'],1), {k[:20]:round(v,2) for k,v in d['phases_ms_per_step'].items()})
PY
done
