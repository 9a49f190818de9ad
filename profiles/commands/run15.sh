set -x
O=gpurun_out/r01o; mkdir -p $O
timeout 900 python tools/selfplay_run.py --games 1000 --slots 1000 --samples > $O/selfplay_1000.log 2> $O/selfplay_1000.err; echo "selfplay1000 rc=$?" >> $O/summary.txt
timeout 900 python tools/selfplay_run.py --games 1000 --slots 500 > $O/selfplay_roll500.log 2> $O/selfplay_roll500.err; echo "selfplay roll rc=$?" >> $O/summary.txt
