set -x
# launch list of the bench command restricted to the kernels of a step (shares, not absolutes): one launch per step now
ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:step_kernel|multi_tensor|record_loss|[Aa]dam' -s 20 -c 150 --csv --log-file gpurun_out/launches_r2e_cfg2.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras --no-e2e --reps 3 > gpurun_out/bench_under_ncu.log 2>&1
python bench.py --steps 200 --warmup 5 > gpurun_out/bench_r2e_n1.json 2> gpurun_out/bench_r2e_n1.err; tail -c 600 gpurun_out/bench_r2e_n1.json; tail -3 gpurun_out/bench_r2e_n1.err
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_r2e_reference_arm.json 2>/dev/null; tail -c 400 gpurun_out/bench_r2e_reference_arm.json
python bench.py --steps 20 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/bench_r2e_n1_k20.json 2>/dev/null; tail -c 300 gpurun_out/bench_r2e_n1_k20.json
