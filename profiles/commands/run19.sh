set -x
O=gpurun_out/r01s; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/summary.txt
for cfg in "1 0" "2 0" "2 1" "3 0" "4 0"; do
  set -- $cfg
  timeout 300 python bench.py --pipeline $1 --pipeline-pdl $2 --steps 4 --no-cpu-baseline --midgame-ply 0 > $O/bench_p$1_pdl$2.log 2> $O/bench_p$1_pdl$2.err; echo "bench p$1 pdl$2 rc=$?" >> $O/summary.txt
done
