#!/bin/bash
# End-of-session capture on the GPU box: tests, smoke, bench line, replay / verification timing, ncu launch lists.
# Outputs -> gpurun_out/.  Every step under its own timeout.   usage: bash tools/capture.sh <tag>
tag=${1:-r2z}
out=gpurun_out
mkdir -p $out
timeout 300 python -m pytest tests -m gpu -q 2>&1 | tail -40 > $out/${tag}_pytest_gpu.txt
tail -3 $out/${tag}_pytest_gpu.txt
timeout 90 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 > $out/${tag}_smoke.txt
cat $out/${tag}_smoke.txt
timeout 240 python bench.py --steps 20 --warmup 3 > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err
tail -c 600 $out/${tag}_bench_n1.json; tail -3 $out/${tag}_bench_n1.err
timeout 90 python tools/replay_time.py 14 5 2>&1 | tail -6 > $out/${tag}_replay_time.txt
cat $out/${tag}_replay_time.txt
NCU="ncu --metrics gpu__time_duration.sum --clock-control none --csv"
timeout 60 $NCU -c 400 --log-file $out/${tag}_launches_msm.csv python tools/prof_run.py msm > /dev/null 2>&1
timeout 60 $NCU -c 400 --log-file $out/${tag}_launches_ntt.csv python tools/prof_run.py ntt > /dev/null 2>&1
timeout 60 $NCU -c 600 --log-file $out/${tag}_launches_small.csv python tools/prof_run.py small > /dev/null 2>&1
timeout 90 $NCU -c 1200 --log-file $out/${tag}_launches_ipa.csv python tools/ipa_time.py 14 1 > /dev/null 2>&1
ls -la $out | tail -12
