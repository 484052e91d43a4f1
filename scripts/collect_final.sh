#!/bin/bash
# Short refresh after the wide-lane-group change: secondary workloads, c1 / SMC ncu captures, full GPU test suite.
O=gpurun_out; TAG=${1:-r1i}; NCU="ncu --clock-control none"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python scripts/bench_all.py > $O/bench_all_$TAG.log 2>&1
$NCU --set full --import-source on -k regex:k_sep_sampler -s 1 -c 1 -o $O/prof_c1_$TAG -f python scripts/prof_misc.py c1 > /dev/null 2>&1
$NCU --set full --import-source on -k regex:k_smc_move -s 2 -c 1 -o $O/prof_smc_$TAG -f python scripts/prof_misc.py smc > /dev/null 2>&1
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $O/launches_smc_$TAG.csv python scripts/prof_misc.py smc > /dev/null 2>&1
python bench.py > $O/bench_$TAG.json 2> $O/bench_$TAG.err
cut -c1-300 $O/bench_$TAG.json; grep -c "^{" $O/bench_all_$TAG.log; ls -la $O/*$TAG*
