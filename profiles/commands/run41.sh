O=gpurun_out/r01p1; mkdir -p $O
timeout 300 python tools/nn_forward.py --boards 32768 --iters 3 --warmup 1 > $O/nn_plain.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3x3 -s 24 -c 4 -o $O/prof_conv_final python tools/nn_forward.py --boards 32768 --iters 3 --warmup 1 > $O/ncu_conv.log 2>&1; echo "ncu conv rc=$?" >> $O/summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:heads_kernel -s 1 -c 1 -o $O/prof_heads_final python tools/nn_forward.py --boards 32768 --iters 3 --warmup 1 > $O/ncu_heads.log 2>&1; echo "ncu heads rc=$?" >> $O/summary.txt
