#!/bin/bash
# One gpurun call (1 GPU): ncu launch lists + full captures for profiles/ (round 2).  bash scripts/r2/collect.sh <tag>
TAG=${1:-r2}
O=gpurun_out
mkdir -p $O
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum -c 600 --csv --log-file $O/launches_bench_$TAG.csv python bench.py --steps 1 --warmup 3 --draws-per-step 2 --no-cpu-baseline --no-secondary > /dev/null 2>&1
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $O/launches_smc_$TAG.csv python scripts/prof_misc.py smc > /dev/null 2>&1
PROF_L=10 $NCU --set full --import-source on -k regex:k_dense_tc -s 3 -c 2 -o $O/prof_step_$TAG -f python scripts/prof_tc.py > /dev/null 2>&1
$NCU --set full --import-source on -k regex:k_smc_ -s 6 -c 3 -o $O/prof_smc_$TAG -f python scripts/prof_misc.py smc > /dev/null 2>&1
$NCU --set full --import-source on -k regex:k_acf_rfft -s 1 -c 1 -o $O/prof_acf_$TAG -f python scripts/prof_misc.py acf > /dev/null 2>&1
ls -la $O/*_$TAG.* | tail
