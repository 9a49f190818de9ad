O=gpurun_out/r01q1; mkdir -p $O
for rep in 1 2; do
  timeout 300 python bench.py --steps 6 --no-cpu-baseline --midgame-ply 0 > $O/bench_prof_$rep.log 2> $O/bench_prof_$rep.err
  timeout 300 python bench.py --steps 6 --no-cpu-baseline --midgame-ply 0 --no-profiling > $O/bench_noprof_$rep.log 2> $O/bench_noprof_$rep.err
done
