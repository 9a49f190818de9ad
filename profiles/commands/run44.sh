O=gpurun_out/r01t11; mkdir -p $O
timeout 300 python scratch/repro_async.py > $O/repro.log 2>&1; echo "repro rc=$?" >> $O/summary.txt
timeout 600 python -m pytest tests/test_gpu_training.py -x -q -s > $O/pytest_train.log 2>&1; echo "pytest train rc=$?" >> $O/summary.txt
