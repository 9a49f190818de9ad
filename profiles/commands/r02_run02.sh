O=gpurun_out/r02b; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_resnet.py tests/test_net_golden.py tests/test_gpu_search.py -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/summary.txt
B="--no-extra-configs --no-cpu-baseline --midgame-ply 0 --steps 6"
OZ_CONV_STATIC=1 timeout 300 python bench.py $B > $O/bench_static.log 2> $O/bench_static.err; echo "static rc=$?" >> $O/summary.txt
timeout 300 python bench.py $B > $O/bench_dynamic.log 2> $O/bench_dynamic.err; echo "dynamic rc=$?" >> $O/summary.txt
OZ_CONV_STATIC=1 timeout 300 python bench.py $B --pipeline 2 > $O/bench_static_p2.log 2> $O/bench_static_p2.err; echo "static p2 rc=$?" >> $O/summary.txt
timeout 300 python bench.py $B --pipeline 2 > $O/bench_dynamic_p2.log 2> $O/bench_dynamic_p2.err; echo "dynamic p2 rc=$?" >> $O/summary.txt
timeout 300 python bench.py $B --pipeline 2 --pipeline-pdl 0 > $O/bench_dynamic_p2_nopdl.log 2> $O/bench_dynamic_p2_nopdl.err; echo "dynamic p2 nopdl rc=$?" >> $O/summary.txt
timeout 300 python bench.py $B --pipeline 3 > $O/bench_dynamic_p3.log 2> $O/bench_dynamic_p3.err; echo "dynamic p3 rc=$?" >> $O/summary.txt
timeout 300 python bench.py $B --pipeline 4 > $O/bench_dynamic_p4.log 2> $O/bench_dynamic_p4.err; echo "dynamic p4 rc=$?" >> $O/summary.txt
OZ_CONV_STATIC=1 timeout 300 python bench.py $B > $O/bench_static2.log 2> $O/bench_static2.err; echo "static2 rc=$?" >> $O/summary.txt
timeout 300 python bench.py $B > $O/bench_dynamic2.log 2> $O/bench_dynamic2.err; echo "dynamic2 rc=$?" >> $O/summary.txt
for f in $O/bench_*.log; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1], round(d['value']), round(d['ms_per_step'],2), d['roofline']['achieved'], d['phases_ms_per_step'])
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done > $O/table.txt
