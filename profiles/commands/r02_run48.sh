O=gpurun_out/r02av; mkdir -p $O
for g in 4 8 16 32; do for p in 1 2 4; do
timeout 300 python bench.py --channels 256 --blocks 19 --hidden 256 --sims 1600 --games $g --pipeline $p --steps 6 --warmup 3 --no-extra-configs --no-cpu-baseline --no-profiling --midgame-ply 0 > $O/b_${g}_$p.json 2> $O/b_${g}_$p.err
python - <<PY
import json
try:
    d=json.loads(open('$O/b_${g}_$p.json').read().strip().splitlines()[-1]); print($g, $p, round(d['value']), d['ms_per_step'])
except Exception as e: print($g, $p, 'fail', e)
PY
done; done
