O=gpurun_out/r02f; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_training.py -m gpu -x -q -k wgrad > $O/pytest_wgrad.log 2>&1; echo "pytest wgrad rc=$?" >> $O/summary.txt
timeout 900 python -m pytest tests/test_gpu_training.py tests/test_gpu_statistical.py tests/test_gpu_search.py -m gpu -q > $O/pytest_rest.log 2>&1; echo "pytest rest rc=$?" >> $O/summary.txt
for b in 256 1024 4096; do timeout 600 python tools/train_bench.py --batch $b --steps 30 >> $O/train_bench.jsonl 2>> $O/train_bench.err; echo "train_bench $b rc=$?" >> $O/summary.txt; done
timeout 600 python tools/reference_on_gpu.py > $O/reference_on_gpu.json 2> $O/reference_on_gpu.err; echo "reference_on_gpu rc=$?" >> $O/summary.txt
timeout 600 python tools/match_run.py --pairs 250 --sims 3200 > $O/match_3200.json 2> $O/match_3200.err; echo "match rc=$?" >> $O/summary.txt
