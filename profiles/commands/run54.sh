O=gpurun_out/r01k1; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/summary.txt
timeout 300 python tools/single_game_latency.py --graphs 1 > $O/single_g1.log 2> $O/single_g1.err
timeout 300 python tools/single_game_latency.py --graphs 1 --channels 256 --blocks 19 --hidden 256 --sims 1600 > $O/single_256.log 2> $O/single_256.err
for b in 32 64 128 256 512 1024; do timeout 300 python tools/nn_forward.py --boards $b --iters 200 --warmup 20 >> $O/nn_small.log 2>&1; done
for g in 2 8; do timeout 600 python bench.py --channels 256 --blocks 19 --hidden 256 --sims 1600 --games $g --steps 2 --no-cpu-baseline --midgame-ply 0 > $O/sweep256_$g.log 2> $O/sweep256_$g.err; done
timeout 600 python bench.py --steps 4 --no-cpu-baseline --midgame-ply 0 > $O/bench.log 2> $O/bench.err
