O=gpurun_out/r02bd; mkdir -p $O
for w in 2 4 8 16; do OZ_SELECT_WARPS=$w timeout 300 python tools/single_game_latency.py --searches 40 > $O/single_w$w.json 2>&1; echo "warps $w $(tail -1 $O/single_w$w.json | cut -c90-200)"; done
for w in 2 4 8; do OZ_SELECT_WARPS=$w timeout 600 python bench.py --steps 8 --warmup 3 --no-extra-configs --no-cpu-baseline --midgame-ply 0 > $O/bench_w$w.json 2> $O/bench_w$w.err; python - <<PY
import json
d=json.loads(open('$O/bench_w$w.json').read().strip().splitlines()[-1])
print($w, d['value'], d['ms_per_step'], d['phases_ms_per_step']['select_encode'])
PY
done
