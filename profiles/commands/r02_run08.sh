O=gpurun_out/r02y; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_training.py -m gpu -x -q > $O/pytest_training.log 2>&1; echo "pytest training rc=$?" >> $O/summary.txt
for b in 256 1024 4096; do timeout 600 python tools/train_bench.py --batch $b --steps 30 >> $O/train_bench.jsonl 2>> $O/train_bench.err; echo "train_bench $b rc=$?" >> $O/summary.txt; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $O/train4096_launches.csv python tools/train_bench.py --batch 4096 --steps 3 --skip-pytorch --graphs 0 > $O/ncu_train4096.log 2>&1; echo "ncu train4096 rc=$?" >> $O/summary.txt
timeout 600 python tools/train_grad_check.py > $O/grad_check.log 2>&1; echo "grad_check rc=$?" >> $O/summary.txt
