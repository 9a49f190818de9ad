O=gpurun_out/r01t3; mkdir -p $O
timeout 600 python scratch/train_diag2.py > $O/diag2.log 2>&1
