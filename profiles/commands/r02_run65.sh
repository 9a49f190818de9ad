O=gpurun_out/r02bm; mkdir -p $O
OZ_HEADS_PAIRED=1 timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_paired.log 2>&1; echo "pytest paired rc=$? $(tail -1 $O/pytest_gpu_paired.log)" >> $O/summary.txt
for hp in 0 1 0 1; do
  OZ_HEADS_PAIRED=$hp timeout 300 python tools/single_game_latency.py --searches 40 2>&1 | grep '^{' | cut -c1-250 >> $O/single_hp$hp.json
  OZ_HEADS_PAIRED=$hp timeout 300 python bench.py --steps 8 --warmup 3 --no-extra-configs --no-cpu-baseline --midgame-ply 0 2> $O/bench_hp$hp.err | tail -1 >> $O/bench_hp$hp.json
done
OZ_HEADS_PAIRED=1 timeout 300 python tools/resnet_error.py > $O/resnet_error_paired.json 2> $O/resnet_error.err
cat $O/summary.txt
for hp in 0 1; do echo "== paired=$hp"; cat $O/single_hp$hp.json | cut -c1-200; python - $O/bench_hp$hp.json <<'P'
import json,sys
for line in open(sys.argv[1]):
    d=json.loads(line); ph=d['phases_ms_per_step']
    print(round(d['value']), round(d['ms_per_step'],2), 'network', round(ph['network'],3), 'conv', round(ph['of_which_conv'],3), 'heads+encode-in-network', round(ph['network']-ph['of_which_conv'],3))
P
done
tail -c 1500 $O/resnet_error_paired.json
