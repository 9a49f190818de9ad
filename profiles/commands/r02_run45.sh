O=gpurun_out/r02as; mkdir -p $O
for w in 0 4 8 32; do OZ_APPLY_WIDE=$w timeout 600 python bench.py --steps 10 --warmup 3 --no-extra-configs --no-cpu-baseline > $O/bench_w$w.json 2> $O/bench_w$w.err; python - <<PY
import json
d=json.loads(open('$O/bench_w$w.json').read().strip().splitlines()[-1])
print($w, d['value'], d['ms_per_step'], d['phases_ms_per_step'], d['midgame']['ms_per_step'])
PY
done
