O=gpurun_out/r02bc; mkdir -p $O
timeout 300 python tools/single_game_latency.py --searches 3 > $O/single_tl.log 2>&1; grep "select round" $O/single_tl.log | tail -30
