O=gpurun_out/r01t12; mkdir -p $O
for i in 1 2; do timeout 900 python -m pytest tests/test_gpu_training.py -x -q > $O/pytest_train_$i.log 2>&1; echo "train pytest $i rc=$?" >> $O/summary.txt; done
for b in 256 1024 4096; do timeout 300 python tools/train_bench.py --batch $b --steps 100 > $O/train_b$b.log 2> $O/train_b$b.err; done
timeout 300 python tools/train_bench.py --batch 256 --steps 100 --graphs 0 --skip-pytorch > $O/train_b256_eager.log 2>&1
