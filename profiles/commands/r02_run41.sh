O=gpurun_out/r02ao; mkdir -p $O
timeout 120 python tools/conv_timeline.py --boards 32 > $O/tl_128_32.log 2>&1; cat $O/tl_128_32.log | head -9
timeout 120 python tools/conv_timeline.py --channels 256 --blocks 19 --hidden 256 --boards 64 > $O/tl_256_64.log 2>&1; head -8 $O/tl_256_64.log
timeout 120 python tools/conv_timeline.py --channels 256 --blocks 19 --hidden 256 --boards 256 > $O/tl_256_256.log 2>&1; head -8 $O/tl_256_256.log
