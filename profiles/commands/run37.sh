O=gpurun_out/r01t7; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_training.py -x -q > $O/pytest_train.log 2>&1; echo "pytest rc=$?" >> $O/summary.txt
for b in 256 1024 4096; do timeout 300 python tools/train_bench.py --batch $b --steps 50 > $O/train_b$b.log 2> $O/train_b$b.err; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/train_launches.csv python tools/train_bench.py --batch 256 --steps 2 --graphs 0 --skip-pytorch > $O/ncu.log 2>&1
