O=gpurun_out/r02aa; mkdir -p $O
for i in 1 2 3; do timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_$i.log 2>&1; echo "pytest $i rc=$? $(tail -1 $O/pytest_gpu_$i.log)" >> $O/summary.txt; done
