set -x
# launch list of the bench command (shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 120 --csv --log-file gpurun_out/launches_r2_cfg2.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras --no-e2e --reps 3 > gpurun_out/bench_under_ncu.log 2>&1
# full captures for traffic + stall pictures
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 3 -c 1 -o gpurun_out/prof_cfg2_r2 python tools/profile_step.py cfg2 6 2>&1 | tail -2
ncu --set full --clock-control none -k regex:step_kernel -s 2 -c 1 -o gpurun_out/prof_cfg3_r2 python tools/profile_step.py cfg3 4 2>&1 | tail -2
ncu --set full --clock-control none -k regex:step_kernel -s 2 -c 1 -o gpurun_out/prof_cfg4_r2 python tools/profile_step.py cfg4 4 2>&1 | tail -2
python tools/profile_step.py cfg2 8
python tools/time_small_batch.py 2>&1 | tail -9
