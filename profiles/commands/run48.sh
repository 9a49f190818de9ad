O=gpurun_out/r01p3; mkdir -p $O
for b in 1024 2048 4096 8192 16384 32000; do
  it=$(( 8000000 / b )); [ $it -gt 4000 ] && it=4000
  timeout 300 python tools/nn_forward.py --boards $b --iters $it --warmup 20 >> $O/nn_sweep.log 2>&1
done
