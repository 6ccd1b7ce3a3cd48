#!/bin/bash
# End-of-session capture on the GPU box: tests, smoke, bench line, ncu launch lists.  Outputs -> gpurun_out/.
# usage: bash tools/capture.sh <tag>
tag=${1:-r2z}
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > $out/${tag}_pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 > $out/${tag}_smoke.txt
python bench.py --steps 20 --warmup 3 > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err
NCU="ncu --metrics gpu__time_duration.sum --clock-control none --csv"
$NCU -c 400 --log-file $out/${tag}_launches_msm.csv python tools/prof_run.py msm > /dev/null 2>&1
$NCU -c 400 --log-file $out/${tag}_launches_ntt.csv python tools/prof_run.py ntt > /dev/null 2>&1
$NCU -c 600 --log-file $out/${tag}_launches_small.csv python tools/prof_run.py small > /dev/null 2>&1
$NCU -c 1200 --log-file $out/${tag}_launches_ipa.csv python tools/ipa_time.py 14 1 > /dev/null 2>&1
ls -la $out | tail -12
