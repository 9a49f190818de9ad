O=gpurun_out/r02r; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/summary.txt
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/summary.txt
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench.log 2> $O/bench.err; echo "bench rc=$?" >> $O/summary.txt
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_ref.log 2> $O/bench_ref.err; echo "bench ref rc=$?" >> $O/summary.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file $O/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --midgame-ply 0 --no-extra-configs > $O/ncu_launches.log 2>&1; echo "ncu launches rc=$?" >> $O/summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3x3_pair -s 25 -c 4 -o $O/prof_conv python tools/nn_forward.py --boards 32768 --iters 3 > $O/ncu_conv.log 2>&1; echo "ncu conv rc=$?" >> $O/summary.txt
