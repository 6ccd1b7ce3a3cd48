#!/bin/bash
# End-of-session capture on the GPU box: tests, bench line, ncu launch lists and one full capture.  Outputs -> gpurun_out/.
# usage: bash tools/capture.sh <tag>
tag=${1:-r1s}
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > $out/${tag}_pytest_gpu.txt
python __graft_entry__.py smoke 2>&1 | tail -2 > $out/${tag}_smoke.txt
python bench.py --steps 20 --warmup 3 > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err
NCU="ncu --metrics gpu__time_duration.sum --clock-control none --csv"
$NCU -c 400 --log-file $out/${tag}_launches_msm.csv python tools/prof_run.py msm > /dev/null 2>&1
$NCU -c 600 --log-file $out/${tag}_launches_small.csv python tools/prof_run.py small > /dev/null 2>&1
$NCU -c 1200 --log-file $out/${tag}_launches_ipa.csv python tools/ipa_time.py 14 1 > /dev/null 2>&1
$NCU -c 200 --log-file $out/${tag}_launches_ecfft.csv python tools/ecfft_time.py 14 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:ecfft_stage_kernel -s 8 -c 1 -o $out/${tag}_ecfft_stage python tools/ecfft_time.py 14 > /dev/null 2>&1
ncu -i $out/${tag}_ecfft_stage.ncu-rep --page raw --csv > $out/${tag}_ecfft_stage_ncu_full_raw.csv 2>/dev/null
rm -f $out/${tag}_ecfft_stage.ncu-rep
ls -la $out | tail -15
