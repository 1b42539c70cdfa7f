#!/bin/bash
# usage: tools/profile_round.sh <tag>   — everything DESIGN.md §8 cites, on ONE GPU: the GPU test-suite, bench lines of every workload,
# the reference arm, hardware counters keyed to this build (tools/ncu_counters.py), the ncu launch list and one `--set full` capture of
# the k_traverse launches of a warm C4 frame.  Results land in gpurun_out/ (copy what is to be judged into profiles/).
TAG=$1
O=gpurun_out
python -m pytest tests -m gpu -q > $O/pytest_$TAG.log 2>&1; echo "pytest rc=$? $(tail -1 $O/pytest_$TAG.log)"
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_$TAG.log 2>&1; echo "smoke rc=$? $(tail -1 $O/smoke_$TAG.log)"
timeout 300 python tools/ncu_counters.py c4 c3 c2 > $O/counters_$TAG.log 2>&1; echo "counters rc=$?"
cp $O/ncu_counters.json profiles/ncu_counters.json 2>/dev/null   # so that the bench lines below carry them (same box, same build)
python bench.py --steps 100 --warmup 3 > $O/bench_c4_$TAG.json 2> $O/bench_c4_$TAG.err; echo "bench c4 rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref_c4_$TAG.json 2>/dev/null; echo "reference arm rc=$?"
python bench.py --workload c2 --steps 100 --warmup 3 > $O/bench_c2_$TAG.json 2>/dev/null
python bench.py --workload c2 --bvh reference --steps 100 --warmup 3 > $O/bench_c2ref_$TAG.json 2>/dev/null
python bench.py --workload c3 --steps 30 --warmup 3 > $O/bench_c3_$TAG.json 2>/dev/null
python bench.py --workload c3 --prim analytic --steps 30 --warmup 3 > $O/bench_c3analytic_$TAG.json 2>/dev/null
python bench.py --workload c5 --steps 3 --warmup 3 > $O/bench_c5_$TAG.json 2>/dev/null
for f in c4 c2 c2ref c3 c3analytic c5; do python tools/oneline.py "[$TAG] $f" < $O/bench_${f}_$TAG.json; done
python tools/gif_bench.py > $O/gif_bench_$TAG.json 2>/dev/null; echo "gif rc=$?"
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-c5"
export RTB_LANES=1   # one stream: ncu serialises launches anyway; keeps the launch list in program order
$CMD > $O/plain_$TAG.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_$TAG.csv $CMD > $O/ncu_l_$TAG.log 2>&1; echo "launch list rc=$?"
$CMD > $O/plain2_$TAG.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_traverse" -s 14 -c 7 -f -o $O/prof_$TAG $CMD > $O/ncu_f_$TAG.log 2>&1; echo "full rc=$?"
ncu -i $O/prof_$TAG.ncu-rep --page raw --csv > $O/prof_${TAG}_raw.csv 2>/dev/null
python tools/ncu_summary.py $O/prof_${TAG}_raw.csv $O/prof_${TAG}_summary.csv "ncu --set full --clock-control none --import-source on, C4: the 7 k_traverse_lbvh launches (depth 0..6) of one warm frame, RTB_LANES=1, tag $TAG" && rm -f $O/prof_${TAG}_raw.csv
