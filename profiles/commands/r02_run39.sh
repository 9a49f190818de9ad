O=gpurun_out/r02am; mkdir -p $O
timeout 300 python tools/tower_check.py --iters 200 > $O/tower_check.log 2>&1; echo "tower_check rc=$?" >> $O/summary.txt; head -1 $O/tower_check.log; tail -1 $O/tower_check.log
grep '"blocks": 9' $O/tower_check.log | cut -c1-110
timeout 900 python -m pytest tests/test_gpu_resnet.py -m gpu -x -q > $O/pytest.log 2>&1; tail -2 $O/pytest.log
