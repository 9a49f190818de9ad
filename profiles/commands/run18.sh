set -x
O=gpurun_out/r01r; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/summary.txt
timeout 600 python bench.py > $O/bench.log 2> $O/bench.err; echo "bench rc=$?" >> $O/summary.txt
for g in 2 8 32 128; do
  timeout 600 python bench.py --channels 256 --blocks 19 --hidden 256 --sims 1600 --games $g --steps 2 --no-cpu-baseline --midgame-ply 0 > $O/sweep256_$g.log 2> $O/sweep256_$g.err
done
timeout 600 python bench.py --channels 256 --blocks 19 --hidden 256 --sims 1600 --games 512 --steps 2 --no-cpu-baseline --midgame-ply 0 > $O/sweep256_512.log 2> $O/sweep256_512.err
