# Multi-GPU validation on one box with >= 4 GPUs (the 2-GPU parity tests are tests/test_gpu_multi.py): a 4-rank data-parallel fit against the
# single-GPU fit (LL all-reduce on the README net, flag all-reduce + tile kernel on cfg5's net), the N = 4 bench line.
set -x
for prob in readme wave3d; do
  CUDA_VISIBLE_DEVICES=0 timeout 200 python tools/check_dp.py gpurun_out/dp1_$prob.json $prob
  PINN_COMM_TIMEOUT_S=20 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29711 tools/check_dp.py gpurun_out/dp4_$prob.json $prob 2>&1 | tail -3
  python - <<PY
import json, numpy as np
a = json.load(open('gpurun_out/dp1_$prob.json')); b = json.load(open('gpurun_out/dp4_$prob.json'))
la, lb = np.asarray(a['losses']), np.asarray(b['losses'])
print('$prob: world', b['world'], 'allreduce', b['allreduce'], 'tensor_core', b['tensor_core'],
      'max rel loss diff %.2e' % np.max(np.abs(la - lb) / np.abs(la)), 'params_norm rel %.2e' % (abs(a['params_norm'] - b['params_norm']) / a['params_norm']))
PY
done
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus 4 --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r2_n4.json 2> gpurun_out/bench_r2_n4.err
tail -c 2500 gpurun_out/bench_r2_n4.json; tail -3 gpurun_out/bench_r2_n4.err
