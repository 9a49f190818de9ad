O=gpurun_out/r02br; mkdir -p $O
timeout 140 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$? $(tail -1 $O/pytest_gpu.log)" >> $O/summary.txt
cat $O/summary.txt
