set -x
O=gpurun_out/r01w; mkdir -p $O
for cfg in "1 64 1" "2 128 1" "2 64 1" "2 128 0"; do
  set -- $cfg
  OZ_CONV_RESERVE_TPC=$1 OZ_NARROW_GRID=$2 timeout 300 python bench.py --pipeline 2 --pipeline-pdl $3 --steps 4 --no-cpu-baseline --midgame-ply 0 > $O/bench_res$1_g$2_pdl$3.log 2> $O/bench_res$1_g$2_pdl$3.err; echo "bench $cfg rc=$?" >> $O/summary.txt
done
OZ_CONV_RESERVE_TPC=1 timeout 300 python bench.py --pipeline 1 --steps 4 --no-cpu-baseline --midgame-ply 0 > $O/bench_res1_p1.log 2> $O/bench_res1_p1.err
