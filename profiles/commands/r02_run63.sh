O=gpurun_out/r02bk; mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/summary.txt
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/bench_reference.json 2> $O/bench_reference.err; echo "ref rc=$?" >> $O/summary.txt
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_1gpu.json 2> $O/bench_1gpu.err; echo "bench rc=$?" >> $O/summary.txt
timeout 300 python tools/single_game_latency.py --searches 40 > $O/single.json 2>&1
timeout 300 python tools/single_game_latency.py --searches 20 --channels 256 --blocks 19 --hidden 256 --sims 1600 >> $O/single.json 2>&1
timeout 600 python tools/unmodified_self_play.py --games 3 > $O/selfplay_ctypes.json 2> $O/selfplay_ctypes.err
timeout 600 python tools/unmodified_self_play.py --games 3 --binding cpp > $O/selfplay_cpp.json 2> $O/selfplay_cpp.err
timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$? $(tail -1 $O/pytest_gpu.log)" >> $O/summary.txt
cat $O/summary.txt; grep '^{' $O/single.json | cut -c1-330; tail -1 $O/selfplay_ctypes.json | cut -c1-420; tail -1 $O/selfplay_cpp.json | cut -c1-420
python - <<'P'
import json
d=json.loads(open('gpurun_out/r02bk/bench_1gpu.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['whole_step_frac'], d['phases_ms_per_step'], d['clocks'])
print([(p['leaf_batch_per_gpu'], round(p['value'])) for p in d['config5']['points']], d['config3_leaf_batch_256']['value'], d['config4']['value'], d['cpu_baseline'])
r=json.loads(open('gpurun_out/r02bk/bench_reference.json').read().strip().splitlines()[-1]); print(r['value'], r['steps'], r['warmup'])
P
