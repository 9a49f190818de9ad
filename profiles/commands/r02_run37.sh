O=gpurun_out/r02ak; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_resnet.py tests/test_net_golden.py tests/test_gpu_training.py -m gpu -x -q > $O/pytest.log 2>&1; tail -2 $O/pytest.log
for b in 512 2048 16384; do timeout 200 python tools/nn_forward.py --boards $b --channels 256 --blocks 19 --hidden 256 --iters 20 > $O/nn256_$b.json 2>&1; tail -1 $O/nn256_$b.json | cut -c1-400; done
timeout 200 python tools/nn_forward.py --boards 32000 --iters 20 > $O/nn128_32000.json 2>&1; tail -1 $O/nn128_32000.json | cut -c1-400
