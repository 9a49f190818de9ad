O=gpurun_out/r01p4; mkdir -p $O
cp othello-alphazero_b200/othello_mcts/libothello_b200.so /tmp/lib_first.so
OZ_NVCC_FLAGS="-DOZ_CONV_EVICT_NORMAL" python othello-alphazero_b200/build_ext.py --force > $O/build_normal.log 2>&1
cp othello-alphazero_b200/othello_mcts/libothello_b200.so /tmp/lib_normal.so
for rep in 1 2 3; do
  for v in first normal; do
    cp /tmp/lib_$v.so othello-alphazero_b200/othello_mcts/libothello_b200.so
    echo "== $v $rep" >> $O/ab.log
    timeout 300 python tools/nn_forward.py --boards 32000 --iters 60 --warmup 10 >> $O/ab.log 2>&1
  done
done
for v in first normal; do
  cp /tmp/lib_$v.so othello-alphazero_b200/othello_mcts/libothello_b200.so
  timeout 300 python bench.py --steps 4 --no-cpu-baseline --midgame-ply 0 > $O/bench_$v.log 2> $O/bench_$v.err
done
