#!/bin/bash
# One gpurun call: bench (both arms), secondary workloads, ncu launch lists and full captures.
# Usage (from the repo root, on the GPU box):  bash scripts/collect_profiles.sh <tag>
TAG=${1:-r1}
O=gpurun_out
mkdir -p $O
python bench.py > $O/bench_$TAG.json 2> $O/bench_$TAG.err
python bench.py --impl reference --steps 3 > $O/bench_ref_$TAG.json 2>> $O/bench_$TAG.err
python scripts/bench_all.py > $O/bench_all_$TAG.log 2>&1
python scripts/time_tc.py > $O/time_tc_$TAG.log 2>&1
for d in 0 1 2 3; do BK_TC_DEBUG=$d timeout 120 python scripts/time_step.py >> $O/time_step_$TAG.log 2>&1; done
BK_TC_FUSE=0 timeout 120 python scripts/time_step.py >> $O/time_step_$TAG.log 2>&1
./scripts/micro/stream_pattern > $O/stream_pattern_$TAG.log 2>&1
timeout 300 python scripts/e2e_sweep.py > $O/e2e_sweep_$TAG.log 2>&1
python scripts/time_c3.py > $O/time_c3_$TAG.log 2>&1
python scripts/c5_full.py > $O/c5_full_$TAG.jsonl 2>&1
BK_HLR_DEBUG=2 python scripts/time_hlr_grad.py > $O/time_hlr_grad_$TAG.log 2>&1   # bit 2: gradient-only kernel through the public call
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $O/launches_bench_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $O/launches_smc_$TAG.csv python scripts/prof_misc.py smc > /dev/null 2>&1
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $O/launches_c3_$TAG.csv python scripts/prof_misc.py c3 > /dev/null 2>&1
# k_dense_tc launches with PROF_L=10: #0 GRAD (cache fill), then per draw one fused STEP launch (9 leapfrog steps) and one GRAD
PROF_L=10 $NCU --set full --import-source on -k regex:k_dense_tc -s 3 -c 2 -o $O/prof_step_$TAG -f python scripts/prof_tc.py > /dev/null 2>&1
PROF_L=10 PROF_N=3 $NCU --metrics gpu__time_duration.sum -c 60 --csv --log-file $O/launches_draw_$TAG.csv python scripts/prof_tc.py > /dev/null 2>&1
$NCU --set full --import-source on -k regex:k_acf_fft -s 1 -c 1 -o $O/prof_acf_$TAG -f python scripts/prof_misc.py acf > /dev/null 2>&1
$NCU --set full --import-source on -k regex:k_radix_scatter -s 1 -c 1 -o $O/prof_sort_$TAG -f python scripts/prof_misc.py sort > /dev/null 2>&1
$NCU --set full --import-source on -k regex:k_hlr -s 22 -c 3 -o $O/prof_hlr_$TAG -f python scripts/prof_misc.py c3 > /dev/null 2>&1
$NCU --set full --import-source on -k regex:k_ess_stream -s 1 -c 1 -o $O/prof_ess_$TAG -f python scripts/prof_misc.py ess > /dev/null 2>&1
$NCU --set full --import-source on -k regex:k_smc_move -s 2 -c 1 -o $O/prof_smc_$TAG -f python scripts/prof_misc.py smc > /dev/null 2>&1
PROF_L=10 $NCU --set full --import-source on -k regex:k_hmc_ -s 2 -c 2 -o $O/prof_rows_$TAG -f python scripts/prof_tc.py > /dev/null 2>&1
ls -la $O | tail -20
cat $O/bench_$TAG.json | cut -c1-600
