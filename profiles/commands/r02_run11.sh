O=gpurun_out/r02k; mkdir -p $O
B="--no-extra-configs --no-cpu-baseline --midgame-ply 0 --steps 6"
run() { name=$1; shift; timeout 300 env "$@" python bench.py $B $EXTRA > $O/bench_$name.log 2> $O/bench_$name.err; echo "$name rc=$?" >> $O/summary.txt; }
EXTRA="" run plain OZ_PIPE_PRIO=1
EXTRA="" run plain_73 OZ_CONV_CLUSTERS=73
EXTRA="--pipeline 2" run p2_wide_73 OZ_PIPE_PRIO=1 OZ_PIPE_WIDE=1 OZ_CONV_CLUSTERS=73
EXTRA="--pipeline 2" run p2_wide_72 OZ_PIPE_PRIO=1 OZ_PIPE_WIDE=1 OZ_CONV_CLUSTERS=72
EXTRA="--pipeline 3" run p3_wide_73 OZ_PIPE_PRIO=1 OZ_PIPE_WIDE=1 OZ_CONV_CLUSTERS=73
EXTRA="--pipeline 2" run p2_wide_73_prio2 OZ_PIPE_PRIO=2 OZ_PIPE_WIDE=1 OZ_CONV_CLUSTERS=73
EXTRA="--pipeline 2" run p2_narrow OZ_PIPE_PRIO=1
EXTRA="" run plain2 OZ_PIPE_PRIO=1
for f in $O/bench_*.log; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1], round(d['value']), round(d['ms_per_step'],2), round(d['roofline']['achieved'],1), {k[:24]:round(v,2) for k,v in d['phases_ms_per_step'].items()})
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done > $O/table.txt
