O=gpurun_out/r01f4; mkdir -p $O
for i in 1 2 3 4; do timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_$i.log 2>&1; echo "pytest $i rc=$?" >> $O/summary.txt; done
