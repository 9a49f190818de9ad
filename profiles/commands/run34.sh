O=gpurun_out/r01t4; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_training.py -x -q > $O/pytest_train.log 2>&1; echo "pytest rc=$?" >> $O/summary.txt
