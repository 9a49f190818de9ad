O=gpurun_out/r02ba; mkdir -p $O
timeout 600 python tools/unmodified_self_play.py --games 3 > $O/selfplay_ctypes.json 2> $O/selfplay_ctypes.err; tail -1 $O/selfplay_ctypes.json; tail -2 $O/selfplay_ctypes.err
timeout 600 python tools/unmodified_self_play.py --games 3 --binding cpp > $O/selfplay_cpp.json 2> $O/selfplay_cpp.err; tail -1 $O/selfplay_cpp.json; tail -2 $O/selfplay_cpp.err
