O=gpurun_out/r01p2; mkdir -p $O
timeout 300 python tools/nn_forward.py --boards 16384 --channels 256 --blocks 19 --hidden 256 --iters 3 --warmup 1 > $O/nn256_plain.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3x3 -s 44 -c 2 -o $O/prof_conv256 python tools/nn_forward.py --boards 16384 --channels 256 --blocks 19 --hidden 256 --iters 3 --warmup 1 > $O/ncu_conv256.log 2>&1; echo "ncu conv256 rc=$?" >> $O/summary.txt
