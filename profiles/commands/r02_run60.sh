O=gpurun_out/r02bh; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$? $(tail -1 $O/pytest_gpu.log)" >> $O/summary.txt
for hb in 4 0; do
  OZ_HEADS_BOARDS=$hb timeout 300 python tools/single_game_latency.py --searches 40 > $O/single_128_hb$hb.json 2>&1
  OZ_HEADS_BOARDS=$hb timeout 300 python tools/single_game_latency.py --searches 20 --channels 256 --blocks 19 --hidden 256 --sims 1600 > $O/single_256_hb$hb.json 2>&1
  for g in 2 8 32; do
    OZ_HEADS_BOARDS=$hb timeout 300 python bench.py --channels 256 --blocks 19 --hidden 256 --sims 1600 --games $g --steps 10 --warmup 3 --no-extra-configs --no-cpu-baseline --midgame-ply 0 > $O/c5_g${g}_hb$hb.json 2> $O/c5_g${g}_hb$hb.err
  done
  OZ_HEADS_BOARDS=$hb timeout 300 python bench.py --games 8 --steps 10 --warmup 3 --no-extra-configs --no-cpu-baseline --midgame-ply 0 > $O/c3_g8_hb$hb.json 2> $O/c3_g8_hb$hb.err
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 3000 --csv --log-file $O/c5_g8_launches.csv python bench.py --channels 256 --blocks 19 --hidden 256 --sims 1600 --games 8 --steps 1 --warmup 3 --no-extra-configs --no-cpu-baseline --midgame-ply 0 > $O/ncu_c5.log 2>&1; echo "ncu rc=$?" >> $O/summary.txt
cat $O/summary.txt
for f in $O/single_*.json; do echo $f; grep '^{' $f | cut -c1-260; done
for f in $O/c5_*.json $O/c3_*.json; do echo $f; cut -c1-160 $f; done
