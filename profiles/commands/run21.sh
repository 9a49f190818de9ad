set -x
O=gpurun_out/r01u; mkdir -p $O
timeout 300 python bench.py --pipeline 1 --games 500 --steps 4 --no-cpu-baseline --midgame-ply 0 > $O/bench_g500_p1.log 2> $O/bench_g500_p1.err
for cfg in "2 1" "2 0"; do
  set -- $cfg
  timeout 300 python bench.py --pipeline $1 --pipeline-pdl $2 --steps 4 --no-cpu-baseline --midgame-ply 0 > $O/bench_p$1_pdl$2.log 2> $O/bench_p$1_pdl$2.err; echo "bench p$1 pdl$2 rc=$?" >> $O/summary.txt
done
