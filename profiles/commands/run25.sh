set -x
O=gpurun_out/r01y; mkdir -p $O
timeout 300 python tools/single_game_latency.py --graphs 1 > $O/single_g1.log 2> $O/single_g1.err
timeout 300 python tools/single_game_latency.py --graphs 0 > $O/single_g0.log 2> $O/single_g0.err
timeout 300 python tools/single_game_latency.py --graphs 1 --threads 1 --batch 64 > $O/single_b64.log 2> $O/single_b64.err
timeout 300 python tools/single_game_latency.py --graphs 1 --channels 256 --blocks 19 --hidden 256 --sims 1600 > $O/single_256.log 2> $O/single_256.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/single_launches.csv python tools/single_game_latency.py --graphs 0 --searches 2 > $O/ncu.log 2>&1
