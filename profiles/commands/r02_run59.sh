O=gpurun_out/r02bg; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$? $(tail -1 $O/pytest_gpu.log)" >> $O/summary.txt
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/summary.txt
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_1gpu.json 2> $O/bench_1gpu.err; echo "bench rc=$?" >> $O/summary.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"select_kernel|expand_kernel|encode_kernel|apply_kernel|heads_kernel" -s 188 -c 20 -o $O/prof_tree python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra-configs --midgame-ply 0 > $O/ncu_tree.log 2>&1; echo "ncu tree rc=$?" >> $O/summary.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra-configs --midgame-ply 0 > $O/ncu_list.log 2>&1; echo "ncu list rc=$?" >> $O/summary.txt
cat $O/summary.txt; tail -c 1500 $O/bench_1gpu.json
