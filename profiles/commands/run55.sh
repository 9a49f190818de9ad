O=gpurun_out/r01k2; mkdir -p $O
for i in 1 2 3 4 5 6; do timeout 600 python -m pytest tests/test_gpu_training.py -x -q > $O/pytest_train_$i.log 2>&1; echo "train $i rc=$?" >> $O/summary.txt; done
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "all rc=$?" >> $O/summary.txt
