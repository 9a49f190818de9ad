O=gpurun_out/r02be; mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/summary.txt; tail -1 $O/smoke.log | cut -c1-200
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/bench_reference.json 2> $O/bench_reference.err; echo "ref rc=$?" >> $O/summary.txt
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_1gpu.json 2> $O/bench_1gpu.err; echo "bench rc=$?" >> $O/summary.txt
timeout 600 python tools/reference_on_gpu.py > $O/reference_on_b200.json 2> $O/reference_on_b200.err; echo "refgpu rc=$?" >> $O/summary.txt
timeout 300 python tools/single_game_latency.py --searches 40 > $O/single.json 2>&1
timeout 300 python tools/single_game_latency.py --searches 20 --channels 256 --blocks 19 --hidden 256 --sims 1600 >> $O/single.json 2>&1
timeout 600 python tools/unmodified_self_play.py --games 3 > $O/selfplay_ctypes.json 2> $O/selfplay_ctypes.err
timeout 600 python tools/unmodified_self_play.py --games 3 --binding cpp > $O/selfplay_cpp.json 2> $O/selfplay_cpp.err
timeout 900 python tools/selfplay_run.py > $O/selfplay_1000.json 2> $O/selfplay.err
timeout 900 python tools/match_run.py --pairs 250 --sims 3200 > $O/match_3200.json 2> $O/match.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 1500 --csv --log-file $O/single_launches_warm.csv python tools/single_game_latency.py --searches 2 > $O/ncu.log 2>&1; echo "ncu rc=$?" >> $O/summary.txt
cat $O/summary.txt; grep '^{' $O/single.json | cut -c1-330; tail -1 $O/selfplay_ctypes.json | cut -c1-420; tail -1 $O/selfplay_1000.json | cut -c1-300; tail -1 $O/match_3200.json | cut -c1-300
