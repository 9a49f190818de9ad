O=gpurun_out/r02az; mkdir -p $O
run() { # games pipeline clusters
OZ_CONV_CLUSTERS=$3 timeout 300 python bench.py --channels 256 --blocks 19 --hidden 256 --sims 1600 --games $1 --pipeline $2 --steps 6 --warmup 3 --no-extra-configs --no-cpu-baseline --no-profiling --midgame-ply 0 > $O/b_$1_$2_$3.json 2> $O/b_$1_$2_$3.err
python - <<PY
import json
try:
    d=json.loads(open('$O/b_$1_$2_$3.json').read().strip().splitlines()[-1]); print('games', $1, 'pipeline', $2, 'conv clusters', $3, round(d['value']), round(d['ms_per_step'],2))
except Exception as e: print($1, $2, $3, 'fail', e)
PY
}
run 8 1 0; run 8 2 37; run 8 2 32; run 8 4 18; run 8 3 24
run 16 1 0; run 16 2 37; run 16 4 18
run 4 1 0; run 4 2 37; run 4 4 18
run 2 1 0; run 2 2 37
timeout 300 python tools/single_game_latency.py --searches 40 > $O/single.json 2>&1; tail -1 $O/single.json
