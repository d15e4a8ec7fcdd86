#!/usr/bin/env python
""" bench.py — collocation points/sec of the pydens fit step on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our arm (one process per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU path (oracle port), rank 0

A "step" is one optimizer step of `Solver.fit` on one batch: fused kernel (sample/read points, forward
jets, residual, MSE, backward) + all-reduce (N>1) + Adam + loss record.  Workload = BASELINE.json
configs[1]: 2-D Poisson, 4-layer [10,12,15,1] tanh MLP, batch 100 000 per GPU (weak scaling).

Numbers in the JSON line:
  value      points/s, device-timed (CUDA events, max over ranks) over EXACTLY K steps replayed from a
             CUDA graph; every step reads its own batch from an HBM-resident pool of distinct batches
             (pool > L2 when K >= 160), so no step sees its input warm in L2.
  e2e        the same metric through the public call `Solver.fit(niters=K, batch_size=B, sampler=...)`
             with HOST batches: per step one H2D copy of the batch from pinned memory and one D2H read of
             the loss, all inside the timed region.
  roofline   the fused kernel alone (CUDA events around K back-to-back launches): algorithmic bytes
             (4*total B/point) over time against the measured HBM peak — the path is FP32-FMA bound
             (SURVEY.md 8d), so `fp32` carries the meaningful fraction: algorithmic flops (6*C*M per
             point) over time against 148 SMs x 128 lanes x 2 flop x the measured SM clock.
  cpu_baseline  oracle/autograd_port.py (the reference algorithm on PyTorch-CPU autograd) timed on this
             box's host cores on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

import numpy as np          # noqa: E402
import torch                # noqa: E402

METRIC = 'collocation points/sec (fit step)'
WORKLOADS = {
    # name: (problem in tests/problems.py, per-GPU batch, lr)
    'cfg2': ('poisson2d', 100000, 0.005),
    'cfg3': ('ode_param', 1000000, 0.01),
    'cfg4': ('heat2d', 1000000, 0.001),
    'cfg5': ('wave3d', 500000, 0.001),
}


def describe(workload, n_gpus):
    import problems as P
    name, batch, lr = WORKLOADS[workload]
    cfg = P.PROBLEMS[name]
    return {'workload': '%s: %s, MLP %s %s, batch_size=%d per GPU' % (
        workload, name, [cfg['ndims'] + cfg['nparams']] + cfg['features'], cfg['activation'], batch),
        'global_batch': batch * n_gpus, 'optimizer': 'Adam lr=%g' % lr, 'parallelism': 'dp%d' % n_gpus}


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
              'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
              'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.FIELDS,
                                          '--format=csv,noheader,nounits', '-lms', '50'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(',')]))

    def stop(self, t0=None, t1=None):
        if self.proc is None:
            return None
        time.sleep(0.12)
        self.proc.terminate()
        rows = [r for t, r in self.rows if (t0 is None or t >= t0 - 0.05) and (t1 is None or t <= t1 + 0.1)] \
            or [r for _, r in self.rows]
        if not rows:
            return None
        try:
            sm = sorted(float(r[0]) for r in rows)
            reasons = []
            for i, nm in enumerate(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap')):
                if any(r[3 + i].lower().startswith('active') for r in rows):
                    reasons.append(nm)
            return {'sm_mhz': sm[len(sm) // 2], 'sm_max_mhz': float(rows[0][1]), 'reasons': reasons,
                    'samples': len(rows), 'power_w_max': max(float(r[2]) for r in rows)}
        except (ValueError, IndexError):
            return None


# ------------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the oracle port of the reference loop on host cores
# ------------------------------------------------------------------------------------------------
def cpu_reference(workload, n_gpus, steps, warmup, budget_s=None):
    import problems as P
    from oracle import autograd_port as ap
    name, batch, lr = WORKLOADS[workload]
    cfg = P.PROBLEMS[name]
    cores = os.cpu_count() or 1
    torch.manual_seed(0)
    prob = ap.Problem(lambda u, *xs, D, V: cfg['equation'](u, *xs, D=D, V=V), ndims=cfg['ndims'],
                      nparams=cfg['nparams'], initial_condition=cfg['ic'], boundary_condition=cfg['bc'],
                      domain=cfg['domain'], features=cfg['features'], activation=cfg['activation'],
                      variables=cfg.get('variables'))
    gbatch = batch * n_gpus
    # bound the sample so the run ends within minutes: cap points per step
    sample_batch = min(gbatch, 100000 if name != 'wave3d' else 20000)
    ranges = cfg['ranges']

    def stream(i):
        cols = [torch.rand((sample_batch, 1)) * (hi - lo) + lo for lo, hi in ranges]
        return torch.cat(cols, dim=1)
    # the reference leaves threading to PyTorch; on a many-core host the default (all cores) can be far
    # from the best setting for these small ops, so give the baseline its best thread count
    best, best_t = None, None
    for nt in sorted({cores, min(cores, 64), min(cores, 32), min(cores, 16), min(cores, 8)}, reverse=True):
        torch.set_num_threads(nt)
        ap.fit(prob, 1, sample_batch, lr=lr, point_stream=stream)
        t = time.perf_counter()
        ap.fit(prob, 2, sample_batch, lr=lr, point_stream=stream)
        t = time.perf_counter() - t
        if best_t is None or t < best_t:
            best, best_t = nt, t
    torch.set_num_threads(best)
    ap.fit(prob, warmup, sample_batch, lr=lr, point_stream=stream)
    t0 = time.perf_counter()
    done = 0
    chunk = max(1, min(steps, 5))
    while done < steps:
        n = min(chunk, steps - done)
        ap.fit(prob, n, sample_batch, lr=lr, point_stream=stream)
        done += n
        if budget_s is not None and time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return {'value': done * sample_batch / dt, 'unit': 'points/s', 'cores': best, 'kind': 'port',
            'sample': '%d steps of batch %d (of the global %d) through oracle/autograd_port.py (reference loop '
                      'on PyTorch-CPU autograd; %d threads = best of a sweep on this %d-core host)'
                      % (done, sample_batch, gbatch, best, cores),
            'ms_per_step': 1e3 * dt / done, 'steps': done}


# ------------------------------------------------------------------------------------------------
def main():
    ap_ = argparse.ArgumentParser()
    ap_.add_argument('--gpus', type=int, default=1)
    ap_.add_argument('--steps', type=int, default=200)
    ap_.add_argument('--warmup', type=int, default=5)
    ap_.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap_.add_argument('--workload', default='cfg2', choices=list(WORKLOADS))
    ap_.add_argument('--global-batch', type=int, default=0,
                     help='fix the GLOBAL batch (strong scaling); default: per-GPU batch of the workload (weak)')
    ap_.add_argument('--no-cpu-baseline', action='store_true')
    ap_.add_argument('--no-e2e', action='store_true')
    args = ap_.parse_args()
    K, W = args.steps, max(args.warmup, 3)
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    os.environ.setdefault('PYDENS_B200_PROGRESS', '0')
    if 'PYDENS_B200_NCCL_DEBUG' not in os.environ:                 # keep stdout to the one JSON line
        os.environ.pop('NCCL_DEBUG', None)
        os.environ['NCCL_DEBUG_FILE'] = '/dev/null'

    if args.impl == 'reference':
        if rank != 0:
            return
        res = cpu_reference(args.workload, args.gpus, K, W)
        line = {'impl': 'reference', 'metric': METRIC, 'value': res['value'], 'unit': 'points/s',
                'n_gpus': args.gpus, 'steps': res['steps'], 'warmup': W, 'ms_per_step': res['ms_per_step'],
                'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
                'data': 'synthetic', 'config': describe(args.workload, args.gpus),
                'cpu_baseline': {k: res[k] for k in ('value', 'unit', 'cores', 'kind', 'sample')},
                'e2e': {'value': res['value'], 'unit': 'points/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
                'gpu_launches': 0}
        print(json.dumps(line))
        return

    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device — the fused fit step has no CPU fallback '
                         '(use --impl reference for the CPU baseline)')
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    dev = torch.device('cuda', local_rank)

    import ctypes as C
    import problems as P
    from pydens_b200 import Solver, D, V, _native
    name, batch, lr = WORKLOADS[args.workload]
    cfg = P.PROBLEMS[name]
    gbatch = args.global_batch if args.global_batch else batch * world
    torch.manual_seed(0)
    solver = Solver(P.bind(name, D, lambda n, init: V(n, data=torch.Tensor([init]))), ndims=cfg['ndims'],
                    nparams=cfg['nparams'], initial_condition=cfg['ic'], boundary_condition=cfg['bc'],
                    domain=cfg['domain'], layout=cfg['layout'], features=cfg['features'],
                    activation=cfg['activation'], device=dev, backend='fused', seed=123)
    eng = solver._get_engine()
    info = eng.info
    total = cfg['ndims'] + cfg['nparams']
    from pydens_b200.engine import shard_batch
    local_n, offset = shard_batch(gbatch, world, rank)
    inv_n = 1.0 / gbatch

    clocks = ClockSampler(local_rank) if rank == 0 else None
    if clocks:
        clocks.start()

    # ---------------- value: K graph-replayed steps over an HBM-resident pool of batches ----------------
    pool_n = min(max(K, 160), 256)          # >= 160 distinct batches: the pool (>= 128 MB at cfg2) exceeds the 126 MB L2
    gen = torch.Generator(device=dev).manual_seed(1000 + rank)
    pool = torch.empty((pool_n, local_n, total), device=dev)
    for k, (lo, hi) in enumerate(cfg['ranges']):
        pool[:, :, k] = torch.rand((pool_n, local_n), generator=gen, device=dev) * (hi - lo) + lo
    solver._make_optimizer('Adam', lr, fused_hint=True)
    opt = solver.optimizer
    ring = torch.zeros(K + W + 8, device=dev)

    def step(i, pts):
        eng._step(pts, None, local_n, inv_n, offset, allreduce=world > 1)
        if world > 1 and eng.comm is None:
            dist.all_reduce(eng.out)
        opt.step()
        _native.check(eng.lib.pinn_record_loss(eng.plan, C.c_void_p(eng.out.data_ptr()), C.c_void_p(ring.data_ptr()),
                                               C.c_int64(ring.numel()), C.c_void_p(eng.step_counter.data_ptr()),
                                               eng._stream()))
    for i in range(W):
        step(i, pool[i % pool_n])
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    graphed = True
    try:
        with torch.cuda.graph(graph):
            for i in range(K):
                step(i, pool[i % pool_n])
    except Exception as exc:            # noqa: BLE001
        graphed = False
        torch.cuda.synchronize()
        sys.stderr.write('graph capture failed (%s): timing plain launches\n' % exc)
    if graphed:
        graph.replay()                  # one untimed replay (uploads the graph), then the timed one
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t_load0 = time.time()
    e0.record()
    if graphed:
        graph.replay()
    else:
        for i in range(K):
            step(i, pool[i % pool_n])
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    value = gbatch * K / (ms_total * 1e-3)
    last_loss = float(eng.out[eng.n_params].item())

    # ---------------- in-kernel sampling variant (the default `fit(sampler=None)` mode) ----------------
    graph2 = torch.cuda.CUDAGraph()
    sampled_value = None
    try:
        torch.cuda.synchronize()
        with torch.cuda.graph(graph2):
            for i in range(K):
                step(i, None)
        graph2.replay()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0.record(); graph2.replay(); e1.record()
        torch.cuda.synchronize()
        ms2 = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
        sampled_value = gbatch * K / (float(ms2.item()) * 1e-3)
    except Exception as exc:            # noqa: BLE001
        torch.cuda.synchronize()
        sys.stderr.write('sampled-variant graph failed: %s\n' % exc)

    # ---------------- roofline: the fused kernel alone ----------------
    for i in range(3):
        eng._step(pool[i % pool_n], None, local_n, inv_n, offset)
    torch.cuda.synchronize()
    e0.record()
    for i in range(K):
        eng._step(pool[i % pool_n], None, local_n, inv_n, offset)
    e1.record()
    torch.cuda.synchronize()
    kern_ms = e0.elapsed_time(e1) / K
    t_load1 = time.time()

    # ---------------- e2e: Solver.fit with host batches (pinned H2D per step, loss D2H per step) --------
    e2e = None
    if not args.no_e2e:
        host_pool = [torch.empty((gbatch, total)).pin_memory() for _ in range(min(K, 32))]
        for hp in host_pool:
            for k, (lo, hi) in enumerate(cfg['ranges']):
                hp[:, k] = torch.rand(gbatch) * (hi - lo) + lo

        class HostBatches:
            i = 0

            def sample(self, size):
                self.i += 1
                return host_pool[self.i % len(host_pool)]
        hb = HostBatches()
        solver.fit(niters=W, batch_size=gbatch, sampler=hb, lr=lr)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        solver.fit(niters=K, batch_size=gbatch, sampler=hb, lr=lr)
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {'value': gbatch * K / float(dt.item()), 'unit': 'points/s',
               'h2d_bytes_per_step': int(local_n * total * 4) * world, 'd2h_bytes_per_step': 4 * world,
               'ms_per_step': 1e3 * float(dt.item()) / K,
               'api': 'Solver.fit(niters=K, batch_size=B, sampler=<host batches in pinned memory>)'}

    clk = clocks.stop(t_load0, t_load1) if clocks else None
    del graph, graph2                    # graphs hold captured NCCL work: drop them before tearing NCCL down
    torch.cuda.synchronize()
    if rank != 0:
        _shutdown(dist, world)
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except (OSError, ValueError):
        pass
    hbm_peak = float(peaks.get('hbm_gbs', 6650.0))
    peak_src = 'measured (MEASURED_PEAKS.json)' if 'hbm_gbs' in peaks else 'fallback (B200_PROFILING.md)'
    bytes_per_launch = info.bytes_per_point * local_n
    flops_per_launch = info.flops_per_point * local_n
    ach_gbs = bytes_per_launch / (kern_ms * 1e-3) / 1e9
    sm_mhz = (clk or {}).get('sm_mhz') or float(peaks.get('sm_max_mhz', 1965.0))
    fp32_peak = info.sm_count * 128 * 2 * sm_mhz * 1e6 / 1e12
    ach_tf = flops_per_launch / (kern_ms * 1e-3) / 1e12
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, 'profiles', 'traffic.json'))).get(args.workload)
    except (OSError, ValueError):
        pass

    line = {
        'metric': METRIC, 'value': value, 'unit': 'points/s', 'n_gpus': world, 'steps': K, 'warmup': W,
        'ms_per_step': ms_total / K, 'higher_is_better': True,
        'scaling': 'strong' if args.global_batch else 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic',
        'config': dict(describe(args.workload, world), global_batch=gbatch,
                       inputs='HBM-resident pool of %d distinct batches (%.0f MB%s), one per step; '
                              'in-kernel Philox sampling variant reported as value_sampled'
                              % (pool_n, pool.numel() * 4 / 1e6, ' > L2' if pool.numel() * 4 > 126e6 else ''),
                       cuda_graph=graphed, final_loss=last_loss,
                       allreduce=('in-kernel over NVLink peer memory (pinn_step_allreduce)' if eng.comm is not None
                                  else ('NCCL' if world > 1 else 'none')),
                       kernel='step_kernel<NF=%d,NS=%d> %d threads/CTA x %d CTAs, %d B smem, %d regs, activations in %s'
                              % (info.nf, info.ns, info.threads_per_cta, min(info.sm_count, (local_n + info.threads_per_cta - 1) // info.threads_per_cta),
                                 info.smem_bytes, info.regs_per_thread, 'smem' if info.activations_in_smem else 'gmem')),
        'value_sampled': sampled_value,
        'gpu_launches': 2 * K,
        'clocks': clk,
        'roofline': {'bound': 'hbm', 'achieved': ach_gbs, 'peak': hbm_peak, 'unit': 'GB/s', 'frac': ach_gbs / hbm_peak,
                     'traffic': traffic, 'kernel': 'step_kernel', 'kernel_ms': kern_ms,
                     'share_of_step': kern_ms / (ms_total / K), 'peak_source': peak_src,
                     'note': 'path is FP32-FMA bound (flop/byte ~1e3): see fp32'},
        'fp32': {'achieved': ach_tf, 'peak': fp32_peak, 'unit': 'TFLOP/s', 'frac': ach_tf / fp32_peak,
                 'flops_per_point': int(info.flops_per_point),
                 'peak_source': '%d SMs x 128 FMA lanes x 2 x %.0f MHz (SM clock sampled under load)' % (info.sm_count, sm_mhz)},
        'e2e': e2e,
    }
    if not args.no_cpu_baseline and world == 1:
        cb = cpu_reference(args.workload, 1, 40, 2, budget_s=20.0)
        line['cpu_baseline'] = {k: cb[k] for k in ('value', 'unit', 'cores', 'kind', 'sample')}
    print(json.dumps(line), flush=True)
    _shutdown(dist, world)


def _shutdown(dist, world):
    """ Tear the process group down; never let a stuck NCCL teardown keep the job alive. """
    if world <= 1:
        return
    t = threading.Thread(target=lambda: (dist.barrier(), dist.destroy_process_group()), daemon=True)
    t.start()
    t.join(20.0)
    sys.stdout.flush()
    sys.stderr.flush()
    os._exit(0)


if __name__ == '__main__':
    main()
