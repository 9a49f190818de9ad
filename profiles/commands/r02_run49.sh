O=gpurun_out/r02aw; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_search.py -m gpu -x -q > $O/pytest.log 2>&1; tail -2 $O/pytest.log
run() { # channels blocks hidden sims games pipeline
timeout 300 python bench.py --channels $1 --blocks $2 --hidden $3 --sims $4 --games $5 --pipeline $6 --steps 6 --warmup 3 --no-extra-configs --no-cpu-baseline --no-profiling --midgame-ply 0 > $O/b_$1_$5_$6.json 2> $O/b_$1_$5_$6.err
python - <<PY
import json
try:
    d=json.loads(open('$O/b_$1_$5_$6.json').read().strip().splitlines()[-1]); print($1, 'games', $5, 'pipeline', $6, round(d['value']), round(d['ms_per_step'],2))
except Exception as e: print($1, $5, $6, 'fail', e)
PY
}
for g in 2 4 8 16 32; do for p in 1 2 4; do run 256 19 256 1600 $g $p; done; done
for g in 2 4 8 16 32; do for p in 1 2 4; do run 128 9 128 800 $g $p; done; done
