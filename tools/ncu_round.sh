#!/bin/bash
# usage: tools/ncu_round.sh <tag> — the ncu part of profile_round.sh alone: launch list of a short C4 bench run and one `--set full` capture of
# the k_traverse launches of a warm frame (each after the same command has run without ncu).
TAG=$1
O=gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-c5"
export RTB_LANES=1   # one stream: ncu serialises launches anyway; keeps the launch list in program order
$CMD > $O/plain_$TAG.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_$TAG.csv $CMD > $O/ncu_l_$TAG.log 2>&1; echo "launch list rc=$?"
$CMD > $O/plain2_$TAG.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_traverse" -s 14 -c 7 -f -o $O/prof_$TAG $CMD > $O/ncu_f_$TAG.log 2>&1; echo "full rc=$?"
ncu -i $O/prof_$TAG.ncu-rep --page raw --csv > $O/prof_${TAG}_raw.csv 2>/dev/null
python tools/ncu_summary.py $O/prof_${TAG}_raw.csv $O/prof_${TAG}_summary.csv "ncu --set full --clock-control none --import-source on, C4: the 7 k_traverse_lbvh launches (depth 0..6) of one warm frame, RTB_LANES=1, tag $TAG" && rm -f $O/prof_${TAG}_raw.csv
rm -f $O/prof_$TAG.ncu-rep
