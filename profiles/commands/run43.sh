O=gpurun_out/r01t10; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest all rc=$?" >> $O/summary.txt
