set -x
O=gpurun_out/r01v; mkdir -p $O
for R in 144 128; do
  OZ_NVCC_FLAGS="-DOZ_CONV_MAXNREG=$R" python othello-alphazero_b200/build_ext.py --force > $O/build_$R.log 2>&1
  for P in 1 2; do
    timeout 300 python bench.py --pipeline $P --steps 4 --no-cpu-baseline --midgame-ply 0 > $O/bench_r${R}_p$P.log 2> $O/bench_r${R}_p$P.err; echo "bench r$R p$P rc=$?" >> $O/summary.txt
  done
done
