#!/bin/bash
# usage: tools/bench_round.sh <tag> — the short form of profile_round.sh after a change that leaves the kernels alone: GPU test-suite, smoke,
# counters re-keyed to this build, and the bench lines of every workload (no launch list, no --set full capture, no GIF sweep).
TAG=$1
O=gpurun_out
python -m pytest tests -m gpu -q > $O/pytest_$TAG.log 2>&1; echo "pytest rc=$? $(tail -1 $O/pytest_$TAG.log)"
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_$TAG.log 2>&1; echo "smoke rc=$? $(tail -1 $O/smoke_$TAG.log)"
timeout 300 python tools/ncu_counters.py c4 c3 c2 > $O/counters_$TAG.log 2>&1; echo "counters rc=$?"
rm -f $O/counters_c*.ncu-rep
cp $O/ncu_counters.json profiles/ncu_counters.json 2>/dev/null
python bench.py --steps 100 --warmup 3 > $O/bench_c4_$TAG.json 2> $O/bench_c4_$TAG.err; echo "bench c4 rc=$?"
python bench.py --workload c2 --steps 100 --warmup 3 > $O/bench_c2_$TAG.json 2>/dev/null
python bench.py --workload c2 --bvh reference --steps 100 --warmup 3 > $O/bench_c2ref_$TAG.json 2>/dev/null
python bench.py --workload c3 --steps 30 --warmup 3 > $O/bench_c3_$TAG.json 2>/dev/null
python bench.py --workload c3 --prim analytic --steps 30 --warmup 3 > $O/bench_c3analytic_$TAG.json 2>/dev/null
python bench.py --workload c5 --steps 3 --warmup 3 > $O/bench_c5_$TAG.json 2>/dev/null
for f in c4 c2 c2ref c3 c3analytic c5; do python tools/oneline.py "[$TAG] $f" < $O/bench_${f}_$TAG.json; done
