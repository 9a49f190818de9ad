O=gpurun_out/r02e; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_resnet.py tests/test_net_golden.py tests/test_gpu_search.py -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/summary.txt
B="--no-extra-configs --no-cpu-baseline --midgame-ply 0 --steps 6"
run() { name=$1; shift; timeout 300 env "$@" python bench.py $B $EXTRA > $O/bench_$name.log 2> $O/bench_$name.err; echo "$name rc=$?" >> $O/summary.txt; }
EXTRA="" run static OZ_CONV_STATIC=1
EXTRA="" run dynamic OZ_CONV_STATIC=0
EXTRA="--pipeline 2" run dyn_p2_prio1 OZ_PIPE_PRIO=1
EXTRA="" run static2 OZ_CONV_STATIC=1
EXTRA="" run dynamic2 OZ_CONV_STATIC=0
for f in $O/bench_*.log; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1], round(d['value']), round(d['ms_per_step'],2), round(d['roofline']['achieved'],1), {k[:24]:round(v,2) for k,v in d['phases_ms_per_step'].items()})
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done > $O/table.txt
