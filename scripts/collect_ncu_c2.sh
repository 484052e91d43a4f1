O=gpurun_out; TAG=r1h; NCU="ncu --clock-control none"
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 300 python -m pytest tests/test_gpu_fused_steps.py tests/test_gpu_dense_tc.py -m gpu -x -q 2>&1 | tail -2
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $O/launches_bench_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
PROF_L=10 PROF_N=3 $NCU --metrics gpu__time_duration.sum -c 60 --csv --log-file $O/launches_draw_$TAG.csv python scripts/prof_tc.py > /dev/null 2>&1
PROF_L=10 $NCU --set full --import-source on -k regex:k_dense_tc -s 3 -c 2 -o $O/prof_step_$TAG -f python scripts/prof_tc.py > /dev/null 2>&1
PROF_L=10 $NCU --set full --import-source on -k regex:k_hmc_ -s 2 -c 2 -o $O/prof_rows_$TAG -f python scripts/prof_tc.py > /dev/null 2>&1
python bench.py --no-cpu-baseline > $O/bench_$TAG.json 2> $O/bench_$TAG.err; cut -c1-200 $O/bench_$TAG.json
ls -la $O/*$TAG*
