""" FusedEngine — host side of the fused fit step: flat parameter/gradient buffers, the plan, the
per-step launch sequence and its CUDA-graph capture, data-parallel sharding.

Per step (replaces reference pydens/model_torch.py:427-464):
    pinn_step         sample + forward jets + residual + MSE + backward     (one CUDA kernel); with
                      torch.distributed initialised: pinn_step_allreduce, whose tail sums [grads | loss] over
                      the ranks through NVLink peer memory (NCCL all_reduce only if IPC mapping is unavailable)
    constraints       one more launch of the same kernel family per lowered constraint, added to [grads | loss]
                      (constraints that do not lower are added by autograd)
    optimizer.step()  torch (fused, capturable Adam by default) on views of the flat buffer
    pinn_record_loss  loss -> device ring, advance the device step counter
The sequence is captured once in a CUDA graph and replayed; nothing in it touches the host.
"""
import ctypes as C
import os

import numpy as np
import torch

from . import _native

_GRAPH_MIN_ITERS = 8
_HOST_GRAPH_MIN_ITERS = 8           # host-sampler mode: one capture per staging buffer
_AUTO_PERSISTENT_STEPS = 50         # optimizer steps per launch when a tiny-batch fit goes to the persistent kernel by itself
_AUTO_PERSISTENT_MIN_ITERS = 16
_RING_LEN = 1 << 15                 # device loss ring (steps between two read-backs of a long fit)


def _dist():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return dist
    return None


def capture_graph(device, body, pool=None):
    """ Capture `body()` into a CUDA graph on a side stream.  Unlike the `torch.cuda.graph` context manager this
    does not run the garbage collector and does not empty the allocator cache, which is what makes a capture
    cost ~20 ms there; a fit only pays a millisecond or two here. """
    import gc
    _reap()                                          # native objects whose teardown was postponed: now is a safe moment
    graph = torch.cuda.CUDAGraph()
    current = torch.cuda.current_stream(device)
    side = torch.cuda.Stream(device=device)
    side.wait_stream(current)
    # the collector must not run finalizers while the capture is open: an old engine's __del__ frees device memory
    # and synchronises streams, which is illegal during a (global-mode) capture and invalidates it
    gc_was_enabled = gc.isenabled()
    gc.disable()
    try:
        with torch.cuda.stream(side):
            try:
                # thread-local error mode: other threads (the NCCL watchdog polling its events when the NCCL fallback
                # all-reduce is part of the step) may keep calling into CUDA while this thread captures
                if pool is not None:
                    graph.capture_begin(pool=pool, capture_error_mode='thread_local')
                else:
                    graph.capture_begin(capture_error_mode='thread_local')
                body()
                graph.capture_end()
            except Exception:
                try:
                    graph.capture_end()
                except Exception:                       # noqa: BLE001
                    pass
                raise
    finally:
        if gc_was_enabled:
            gc.enable()
    current.wait_stream(side)
    return graph


_graveyard = []                                      # (destroy function, handle) postponed because a capture was open


def _capturing():
    try:
        return torch.cuda.is_available() and torch.cuda.is_current_stream_capturing()
    except Exception:                                # noqa: BLE001
        return False


def _destroy(fn, handle):
    """ Tear a native object down now, or — inside a stream capture, where freeing device memory is illegal — later. """
    if _capturing():
        _graveyard.append((fn, handle))
    else:
        fn(handle)


def _reap():
    if _graveyard and not _capturing():
        pending, _graveyard[:] = list(_graveyard), []
        for fn, handle in pending:
            fn(handle)


def shard_batch(batch_size, world, rank):
    """ Contiguous, near-equal split of the global batch over ranks -> (local_n, point_offset).
    The Philox counter is the GLOBAL point index, so the sampled batch does not depend on `world`. """
    base, rem = divmod(int(batch_size), int(world))
    return base + (1 if rank < rem else 0), rank * base + min(rank, rem)


class FusedEngine:
    def __init__(self, solver):
        self.lib = _native.load()
        self.solver = solver
        model, traced, chain = solver.model, solver._traced, solver._chain
        self.device = solver.device
        if self.device.type != 'cuda':
            raise RuntimeError('the fused engine needs a CUDA device (got %s); there is no CPU fallback' % self.device)
        torch.cuda.set_device(self.device)

        # ---- canonical flat layout: W_0, b_0, ..., log_scale, equation variables, everything else ----
        entries, off = [], 0
        w_off, b_off = [], []
        for lin, _, _ in chain:
            w_off.append(off); entries.append((lin.weight, off)); off += lin.weight.numel()
            b_off.append(off); entries.append((lin.bias, off)); off += lin.bias.numel()
        log_scale_off = off
        entries.append((model.log_scale, off)); off += 1
        var_offsets = {}
        for name in traced.var_names:
            p = getattr(model, name)
            var_offsets[name] = off
            entries.append((p, off)); off += p.numel()
        seen = {id(p) for p, _ in entries}
        for p in model.parameters():
            if id(p) not in seen:
                entries.append((p, off)); off += p.numel(); seen.add(id(p))
        self.n_params = (off + 3) // 4 * 4
        self.entries = entries
        self.flat = torch.zeros(self.n_params, dtype=torch.float32, device=self.device)
        self.out = torch.zeros(self.n_params + 4, dtype=torch.float32, device=self.device)
        with torch.no_grad():
            for p, o in entries:
                n = p.numel()
                self.flat[o:o + n].copy_(p.data.reshape(-1).to(torch.float32))
                p.data = self.flat[o:o + n].view(p.shape)
                p.grad = self.out[o:o + n].view(p.shape)

        widths = [model.total] + [c[0].out_features for c in chain]
        acts = [c[1] for c in chain]
        skips = [c[2] for c in chain]
        has_bc = model.boundary_condition is not None
        has_ic = model.raw_initial_condition is not None
        self._spec_args = (widths, acts, model.ndims, model.nparams, has_bc,
                           model.boundary_condition if has_bc else 0.0, has_ic, model.domain)
        self._spec_kwargs = dict(var_offsets=var_offsets, w_off=w_off, b_off=b_off, log_scale_off=log_scale_off,
                                 n_params=self.n_params, skips=skips)
        self.spec = _native.build_spec(*self._spec_args, traced, **self._spec_kwargs)
        self._constraint_plans = {}                       # num -> fused constraint launch (see _constraint_plan)
        plan = C.c_void_p()
        _native.check(self.lib.pinn_plan_create(C.byref(self.spec), self.device.index, C.byref(plan)))
        self.plan = plan
        self.info = _native.PinnPlanInfo()
        _native.check(self.lib.pinn_plan_info(self.plan, C.byref(self.info)))
        ws_bytes = self.lib.pinn_workspace_bytes(self.plan, 1)
        self.workspace = torch.zeros(ws_bytes, dtype=torch.uint8, device=self.device)
        self.step_counter = torch.zeros(1, dtype=torch.int64, device=self.device)
        self.ring = torch.zeros(_RING_LEN, dtype=torch.float32, device=self.device)
        self.steps_done = 0
        self._graphs, self._pipes, self._pinned_batches = {}, {}, {}
        self._opt_obj, self._opt_key, self._loss_host = None, None, None
        self._adam_flat = None
        self._adam_bound_to, self._adam_bound_sig = None, None
        self.seed = solver.seed

        self.comm = None
        dist = _dist()
        if dist is not None:
            seed_t = torch.tensor([self.seed], dtype=torch.int64, device=self.device)
            dist.broadcast(seed_t, 0)
            self.seed = int(seed_t.item())
            dist.broadcast(self.flat, 0)
            self._connect_peers(dist)

    def _connect_peers(self, dist):
        """ Set up the in-kernel all-reduce over NVLink peer memory (pinn_step_allreduce): every rank
        allocates an exchange buffer inside the library, the CUDA-IPC handles are gathered and mapped.
        Any failure (no peer access, more than 8 ranks, several nodes) leaves NCCL in charge. """
        if os.environ.get('PYDENS_B200_FUSED_ALLREDUCE', '1') == '0':
            return
        world, rank = dist.get_world_size(), dist.get_rank()
        ok, comm = 1, C.c_void_p()
        handle = (C.c_ubyte * 64)()
        if world > 8 or self.lib.pinn_comm_create(self.plan, rank, world, C.byref(comm), handle) != 0:
            ok = 0
        handles = [None] * world
        dist.all_gather_object(handles, bytes(handle) if ok else None)
        if ok and all(h is not None for h in handles):
            ok = int(self.lib.pinn_comm_connect(comm, b''.join(handles)) == 0)
        else:
            ok = 0
        flag = torch.tensor([ok], dtype=torch.int32, device=self.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)              # all ranks or none
        if int(flag.item()) == 1:
            self.comm = comm
        elif comm:
            self.lib.pinn_comm_destroy(comm)

    def __del__(self):
        try:
            self._drop_graphs()
            if getattr(self, 'comm', None):
                _destroy(self.lib.pinn_comm_destroy, self.comm)
                self.comm = None
            for entry in getattr(self, '_constraint_plans', {}).values():
                if entry is not None and entry.get('plan'):
                    self.lib.pinn_plan_destroy(entry['plan'])
                    entry['plan'] = None
            if getattr(self, 'plan', None):
                self.lib.pinn_plan_destroy(self.plan)
                self.plan = None
        except Exception:                                  # pragma: no cover
            pass

    def _constraint_plan(self, num):
        """ Constraint `num` as one more launch of the same kernel family (reference :451-457): a plan whose
        residual program is the traced constraint (no derivative channels), run on the constraint's own points
        with 1/n weighting; its [grads | loss] buffer is then added to the step's.  None -> autograd adds it. """
        if num not in self._constraint_plans:
            entry = None
            lowered = self.solver._lower_constraint(num)
            if lowered is not None:
                traced, pts = lowered
                spec = _native.build_spec(*self._spec_args, traced, **self._spec_kwargs)
                plan = C.c_void_p()
                rc = self.lib.pinn_plan_create(C.byref(spec), self.device.index, C.byref(plan))
                if rc == 0:
                    ws = torch.zeros(self.lib.pinn_workspace_bytes(plan, 1), dtype=torch.uint8, device=self.device)
                    entry = dict(plan=plan, spec=spec, points=pts.to(self.device).contiguous(), n=int(pts.shape[0]),
                                 out=torch.zeros(self.n_params + 4, dtype=torch.float32, device=self.device),
                                 workspace=ws)
                elif rc != _native.E_UNSUPPORTED:
                    _native.check(rc)
            self._constraint_plans[num] = entry
        return self._constraint_plans[num]

    def _constraint_step(self, entry):
        _native.check(self.lib.pinn_step(
            entry['plan'], C.c_void_p(self.flat.data_ptr()), C.c_void_p(entry['points'].data_ptr()), None,
            C.c_uint64(self.seed), None, C.c_uint64(0), C.c_uint64(0), C.c_int64(entry['n']),
            C.c_float(1.0 if self._sum_reduction() else 1.0 / entry['n']), C.c_void_p(entry['out'].data_ptr()), None,
            C.c_void_p(entry['workspace'].data_ptr()), C.c_size_t(entry['workspace'].numel()), self._stream()))
        self.out.add_(entry['out'])                       # [grads | loss] += the constraint's

    def _sum_reduction(self):
        """ criterion(reduction='sum') of the fit this engine was built for (Solver._crit_key ends with 'sum') """
        return getattr(self.solver, '_crit_key', ('mse',))[-1] == 'sum'

    def release(self):
        """ Give every parameter its own storage back (the autograd path is taking over). """
        with torch.no_grad():
            for p, _ in self.entries:
                p.data = p.data.clone()
                p.grad = None

    # ------------------------------------------------------------------------------------------
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _step(self, points, cols, n_points, inv_n, point_offset, residual=None, use_counter=True, step_value=0,
              allreduce=False):
        args = (C.c_void_p(self.flat.data_ptr()),
                C.c_void_p(points.data_ptr()) if points is not None else None,
                cols, C.c_uint64(self.seed),
                C.c_void_p(self.step_counter.data_ptr()) if use_counter else None, C.c_uint64(step_value),
                C.c_uint64(point_offset), C.c_int64(n_points), C.c_float(inv_n),
                C.c_void_p(self.out.data_ptr()),
                C.c_void_p(residual.data_ptr()) if residual is not None else None,
                C.c_void_p(self.workspace.data_ptr()), C.c_size_t(self.workspace.numel()), self._stream())
        if allreduce and self.comm is not None:
            _native.check(self.lib.pinn_step_allreduce(self.plan, self.comm, *args))
        else:
            _native.check(self.lib.pinn_step(self.plan, *args))

    def _step_adam(self, points, cols, n_points, inv_n, point_offset, adam, allreduce=False):
        """ The whole step in ONE launch (pinn_step_adam): kernel -> in-kernel all-reduce -> Adam -> loss log. """
        _native.check(self.lib.pinn_step_adam(
            self.plan, self.comm if (allreduce and self.comm is not None) else None,
            C.c_void_p(self.flat.data_ptr()), C.c_void_p(points.data_ptr()) if points is not None else None,
            cols, C.c_uint64(self.seed), C.c_void_p(self.step_counter.data_ptr()), C.c_uint64(point_offset),
            C.c_int64(n_points), C.c_float(inv_n), C.c_void_p(self.out.data_ptr()), None,
            C.c_void_p(self.workspace.data_ptr()), C.c_size_t(self.workspace.numel()), C.byref(adam), self._stream()))

    def loss_and_grads(self, points):
        pts = torch.as_tensor(points, dtype=torch.float32).to(self.device).contiguous()
        n = pts.shape[0]
        residual = torch.empty(n, dtype=torch.float32, device=self.device)
        self._step(pts, None, n, 1.0 / n, 0, residual=residual, use_counter=False)
        torch.cuda.synchronize(self.device)
        return float(self.out[self.n_params].item()), self.out[:self.n_params].clone(), residual

    def sample(self, n_points, cols=None, step=None, point_offset=0):
        """ The batch the in-kernel sampler produces (tests / replay). """
        total = self.solver.model.total
        out = torch.empty((n_points, total), dtype=torch.float32, device=self.device)
        arr = _native.make_columns(cols, total)
        _native.check(self.lib.pinn_sample(
            self.plan, arr, C.c_uint64(self.seed),
            C.c_void_p(self.step_counter.data_ptr()) if step is None else None,
            C.c_uint64(0 if step is None else step), C.c_uint64(point_offset), C.c_int64(n_points),
            C.c_void_p(out.data_ptr()), self._stream()))
        return out

    def forward(self, pts):
        pts = pts.to(self.device, torch.float32).contiguous()
        n = pts.shape[0]
        u = torch.empty(n, dtype=torch.float32, device=self.device)
        _native.check(self.lib.pinn_forward(self.plan, C.c_void_p(self.flat.data_ptr()), C.c_void_p(pts.data_ptr()),
                                            C.c_int64(n), C.c_void_p(u.data_ptr()),
                                            C.c_void_p(self.workspace.data_ptr()),
                                            C.c_size_t(self.workspace.numel()), self._stream()))
        return u

    # ------------------------------------------------------------------------------------------
    def _optimizer_for(self, optimizer, lr, kwargs):
        """ The optimizer of this fit.  The reference builds a NEW optimizer on every `fit` (model_torch.py:419-422);
        a fresh fused Adam differs from a used one only by its state, so when the request repeats the previous one
        exactly, the existing object is kept and its state is zeroed in place — which keeps the captured CUDA graphs
        (they hold pointers into that state) valid across `fit` calls.  Returns (optimizer, state_is_materialised). """
        solver = self.solver
        if optimizer is None:
            solver._make_optimizer(None, lr)
            return solver.optimizer, bool(solver.optimizer.state)
        params = [p for p in solver.model.parameters() if p.requires_grad]
        key = (optimizer, float(lr), tuple(sorted((k, repr(v)) for k, v in kwargs.items())), tuple(id(p) for p in params))
        prev = solver.optimizer
        if (prev is not None and prev is self._opt_obj and key == self._opt_key and optimizer in ('Adam', 'AdamW')
                and prev.state and not any(kwargs.get(k) for k in ('amsgrad', 'maximize'))):
            with torch.no_grad():
                bound = self._adam_flat is not None and self._adam_bound_to is prev
                if bound:                                   # the state tensors are views of three flat buffers: three launches
                    spans = []
                    for buf in self._adam_flat[:3]:
                        buf.zero_()
                        spans.append((buf.data_ptr(), buf.data_ptr() + 4 * buf.numel()))
                    for st in prev.state.values():          # ... plus whatever state lives elsewhere
                        for v in st.values():
                            if torch.is_tensor(v) and not any(a <= v.data_ptr() < e for a, e in spans):
                                v.zero_()
                else:
                    for st in prev.state.values():
                        for v in st.values():
                            if torch.is_tensor(v):
                                v.zero_()
            return prev, True
        self._drop_graphs()
        solver._make_optimizer(optimizer, lr, fused_hint=True, **kwargs)
        self._opt_obj, self._opt_key = solver.optimizer, key
        return solver.optimizer, False

    def _drop_graphs(self):
        self._graphs = {}
        for pipe in getattr(self, '_pipes', {}).values():
            _destroy(self.lib.pinn_pipe_destroy, pipe)
        self._pipes = {}

    def _pinned_losses(self, n):
        if self._loss_host is None or self._loss_host.numel() < n:
            self._loss_host = torch.zeros(max(n, 4096), dtype=torch.float32).pin_memory()
        return self._loss_host

    def _bind_adam_state(self, opt):
        """ Make the optimizer's per-parameter Adam state views of three flat buffers (moments in parameter layout,
        step counters side by side), so that the persistent multi-step kernel and torch's own `opt.step()` update
        the very same memory.  -> mask of trainable entries. """
        if self._adam_flat is None:
            self._adam_flat = (torch.zeros(self.n_params, dtype=torch.float32, device=self.device),
                               torch.zeros(self.n_params, dtype=torch.float32, device=self.device),
                               torch.zeros(len(self.entries), dtype=torch.float32, device=self.device),
                               torch.zeros(self.n_params, dtype=torch.float32, device=self.device))
        m, v, steps, mask = self._adam_flat
        in_opt = {id(q) for g in opt.param_groups for q in g['params']}
        sig = (id(opt), tuple(sorted(in_opt)))
        if self._adam_bound_to is opt and self._adam_bound_sig == sig and all(
                (st := opt.state.get(p)) is not None and 'exp_avg' in st and st['exp_avg'].data_ptr() == m.data_ptr() + 4 * o
                for p, o in self.entries if id(p) in in_opt):
            return mask                                 # same optimizer, same parameters, still bound: nothing to do
        mask.zero_()
        rebound = False
        with torch.no_grad():
            for idx, (p, o) in enumerate(self.entries):
                n = p.numel()
                if id(p) not in in_opt:
                    continue
                mask[o:o + n] = 1.0
                st = opt.state.get(p)
                mv, vv = m[o:o + n].view(p.shape), v[o:o + n].view(p.shape)
                if st and 'exp_avg' in st:
                    if st['exp_avg'].data_ptr() == mv.data_ptr():
                        continue                                    # already bound
                    mv.copy_(st['exp_avg']); vv.copy_(st['exp_avg_sq']); steps[idx] = float(st['step'])
                else:
                    mv.zero_(); vv.zero_(); steps[idx] = 0.0
                opt.state[p] = {'step': steps[idx], 'exp_avg': mv, 'exp_avg_sq': vv}
                rebound = True
        if rebound:
            self._drop_graphs()                 # captured steps still point at the previous state tensors
        self._adam_bound_to, self._adam_bound_sig = opt, sig
        return mask

    def _fit_persistent(self, niters, batch_size, sampler, opt, k):
        """ `niters` optimizer steps, `k` per launch of the persistent multi-step kernel (pinn_multi_step):
        the reference loop model_torch.py:426-464 with optimizer.step() inside the kernel. """
        solver = self.solver
        total = solver.model.total
        g = opt.param_groups[0]
        mask = self._bind_adam_state(opt)
        m, v, steps, _ = self._adam_flat
        opt_step0 = float(steps.max().item())
        cols, host_sampler = None, None
        if sampler is not None:
            dev_cols = sampler.device_columns() if hasattr(sampler, 'device_columns') else None
            if dev_cols is not None and len(dev_cols) == total and os.environ.get('PYDENS_B200_HOST_SAMPLER') != '1':
                try:
                    cols = _native.make_columns(dev_cols, total)
                except ValueError:
                    cols = None
            if cols is None:
                host_sampler = sampler
        solver.model.train()
        ring, ring_len = self.ring, self.ring.numel()
        start = self.steps_done
        vals = np.empty(niters, dtype=np.float32)
        done, drained = 0, 0
        lr = float(g['lr']) if not torch.is_tensor(g['lr']) else float(g['lr'].item())
        b1, b2 = g['betas']
        while done < niters:
            kk = min(k, niters - done, ring_len)
            if done + kk - drained > ring_len:
                torch.cuda.synchronize(self.device)
                idx = (start + np.arange(drained, done)) % ring_len
                vals[drained:done] = ring.cpu().numpy()[idx]
                drained = done
            pts = None
            if host_sampler is not None:
                host = torch.empty((kk, batch_size, total), dtype=torch.float32).pin_memory()
                for j in range(kk):
                    b = host_sampler.sample(batch_size)
                    host[j].copy_(b if isinstance(b, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(b, dtype=np.float32)))
                pts = host.to(self.device, non_blocking=True)
            _native.check(self.lib.pinn_multi_step(
                self.plan, C.c_void_p(self.flat.data_ptr()), C.c_void_p(m.data_ptr()), C.c_void_p(v.data_ptr()),
                C.c_void_p(mask.data_ptr()), C.c_void_p(steps.data_ptr()), C.c_int(steps.numel()),
                C.c_void_p(pts.data_ptr()) if pts is not None else None, cols, C.c_uint64(self.seed),
                C.c_void_p(self.step_counter.data_ptr()), C.c_int64(batch_size), C.c_int(kk),
                C.c_float(lr), C.c_float(b1), C.c_float(b2), C.c_float(g['eps']), C.c_float(g['weight_decay']),
                C.c_float(opt_step0 + done), C.c_void_p(ring.data_ptr()), C.c_int64(ring_len), self._stream()))
            done += kk
        self.steps_done = start + niters
        torch.cuda.synchronize(self.device)
        idx = (start + np.arange(drained, niters)) % ring_len
        vals[drained:niters] = ring.cpu().numpy()[idx]
        solver.losses.extend(np.array(x, dtype=np.float32) for x in vals)

    def fit(self, niters, batch_size, sampler, loss_terms, optimizer, criterion, lr, **kwargs):
        solver = self.solver
        steps_per_launch = int(kwargs.pop('steps_per_launch', 0) or 0)
        opt, opt_ready = self._optimizer_for(optimizer, lr, kwargs)
        # Tiny batches (the README example: batch_size=100, niters=1500) are launch- and latency-bound one step at a
        # time: when nothing stands in the way, the loop runs in the persistent cluster kernel without being asked to,
        # _AUTO_PERSISTENT_STEPS optimizer steps per launch (same math, same loss log; PYDENS_B200_AUTO_PERSISTENT=0
        # keeps one launch per step).  An explicit steps_per_launch takes precedence.
        requested = steps_per_launch > 0
        if (not requested and niters >= _AUTO_PERSISTENT_MIN_ITERS and 0 < batch_size <= int(self.info.small_batch_points)
                and os.environ.get('PYDENS_B200_AUTO_PERSISTENT', '1') != '0' and os.environ.get('PYDENS_B200_NO_GRAPH') != '1'):
            steps_per_launch = _AUTO_PERSISTENT_STEPS
        if steps_per_launch > 0 and niters > 0:
            g = opt.param_groups[0] if opt.param_groups else {}
            ok = (type(opt) is torch.optim.Adam and len(opt.param_groups) == 1 and not g.get('amsgrad') and not g.get('maximize')
                  and not torch.is_tensor(g.get('lr')) and not self._sum_reduction()
                  and _dist() is None and not solver._constraint_numbers(loss_terms)
                  and 0 < batch_size <= self.lib.pinn_multi_step_max_points(self.plan))
            if not ok and not requested:
                steps_per_launch = 0
            elif ok:
                if solver.optimizer is not self._opt_obj:
                    self._drop_graphs()
                    self._opt_obj, self._opt_key = solver.optimizer, None
                return self._fit_persistent(niters, batch_size, sampler, opt, steps_per_launch)
        if steps_per_launch > 0 and niters > 0:
            import warnings
            warnings.warn('pydens_b200: steps_per_launch ignored (needs Adam, one device, no constraints, a batch of at '
                          'most %d points and a network that fits the persistent kernel)'
                          % self.lib.pinn_multi_step_max_points(self.plan), stacklevel=3)
        if solver.optimizer is not self._opt_obj:          # optimizer=None with an optimizer built elsewhere
            self._drop_graphs()
            self._opt_obj, self._opt_key = solver.optimizer, None
        nums = solver._constraint_numbers(loss_terms)
        eq_term = 'equation' in loss_terms                # reference :448: the equation term is optional (:382-389)
        fused_constraints = [e for e in (self._constraint_plan(n) for n in nums) if e is not None]
        fused_nums = tuple(n for n in nums if self._constraint_plan(n) is not None)
        nums = [n for n in nums if self._constraint_plan(n) is None]     # the rest is added by autograd
        dist = _dist()
        world = dist.get_world_size() if dist is not None else 1
        rank = dist.get_rank() if dist is not None else 0
        local_n, point_offset = shard_batch(batch_size, world, rank)
        if local_n <= 0:
            raise ValueError('batch_size %d is smaller than the number of ranks %d' % (batch_size, world))
        # weight of a point's criterion value in the loss: 1 / (global batch) — the mean of model_torch.py:448 —, or 1 for
        # criteria with reduction='sum'
        inv_n = 1.0 if self._sum_reduction() else 1.0 / float(batch_size)
        total = solver.model.total
        if niters <= 0:
            return

        # every parameter's .grad must be the view of `out` (a previous zero_grad may have dropped it)
        for p, o in self.entries:
            if p.grad is None or p.grad.data_ptr() != self.out.data_ptr() + 4 * o:
                p.grad = self.out[o:o + p.numel()].view(p.shape)
        # parameters that came to life after the engine was built (a V() first met inside a constraint that does not
        # lower) own their .grad: the kernel never overwrites it, so it has to be cleared every step like the
        # reference does (model_torch.py:428)
        entry_ids = {id(p) for p, _ in self.entries}
        stray_params = [p for p in solver.model.parameters() if id(p) not in entry_ids]

        # optimizer.step() inside the step kernel (pinn_step_adam) whenever the optimizer is the plain torch Adam this
        # engine builds and nothing has to touch the gradient between the kernel and the update
        g0 = opt.param_groups[0] if opt.param_groups else {}
        fused_adam = (os.environ.get('PYDENS_B200_FUSED_ADAM', '1') != '0' and type(opt) is torch.optim.Adam
                      and len(opt.param_groups) == 1 and bool(g0.get('capturable')) and not g0.get('amsgrad')
                      and not g0.get('maximize') and not torch.is_tensor(g0.get('lr'))
                      and not fused_constraints and not nums and not stray_params and eq_term
                      and (dist is None or self.comm is not None))
        adam, adam_key = None, None
        if fused_adam:
            mask = self._bind_adam_state(opt)          # the optimizer's state tensors become views of flat buffers
            m_flat, v_flat, steps_flat, _ = self._adam_flat
            adam = _native.PinnAdam(m_flat.data_ptr(), v_flat.data_ptr(), mask.data_ptr(), steps_flat.data_ptr(),
                                    steps_flat.numel(), float(g0['lr']), float(g0['betas'][0]), float(g0['betas'][1]),
                                    float(g0['eps']), float(g0['weight_decay']), self.ring.data_ptr(), self.ring.numel())
            adam_key = (float(g0['lr']), tuple(float(b) for b in g0['betas']), float(g0['eps']), float(g0['weight_decay']))
            opt_ready = True                           # the state exists (zeros for a fresh optimizer)

        if dist is not None:
            # ranks may reach this fit seconds apart (rank 0 plotting or saving between fits): the in-kernel
            # all-reduce waits only a bounded time for its peers, so line the ranks up first
            dist.barrier()

        cols, host_sampler, cols_key = None, None, None
        if sampler is not None:
            dev_cols = sampler.device_columns() if hasattr(sampler, 'device_columns') else None
            if dev_cols is not None and len(dev_cols) == total and os.environ.get('PYDENS_B200_HOST_SAMPLER') != '1':
                try:
                    cols = _native.make_columns(dev_cols, total)
                    cols_key = bytes(cols)
                except ValueError:                    # more mixture components than the kernel takes
                    cols = None
            if cols is None:
                host_sampler = sampler
        solver.model.train()

        ring, ring_len = self.ring, self.ring.numel()
        start = self.steps_done
        loss_idx = self.n_params
        stream = self._stream()
        ring_ptr = ring.data_ptr()
        vals = np.empty(niters, dtype=np.float32)
        drained = [0]

        def drain(upto):
            """ losses of steps [drained, upto) of this fit: device ring -> vals (the ring has been synchronised) """
            if upto > drained[0]:
                idx = (start + np.arange(drained[0], upto)) % ring_len
                vals[drained[0]:upto] = ring.cpu().numpy()[idx]
                drained[0] = upto

        def one_step(points=None, full_points=None):
            if fused_adam:
                self._step_adam(points, cols, local_n, inv_n, point_offset, adam, allreduce=dist is not None)
                return
            if eq_term:
                self._step(points, cols, local_n, inv_n, point_offset, allreduce=dist is not None)
                if dist is not None and self.comm is None:
                    dist.all_reduce(self.out)          # no peer-memory path: NCCL sums [grads | loss]
            else:
                self.out.zero_()                       # loss_terms without 'equation': the constraint terms alone
            for entry in fused_constraints:            # after the all-reduce: every rank adds the same term
                self._constraint_step(entry)
            if nums:
                # constraints that stay on autograd see the step's own batch — the WHOLE global batch, identical
                # on every rank (so that every replica adds the same gradient) and with requires_grad set, as the
                # reference hands them to the constraint (model_torch.py:436, :451-457)
                if full_points is not None:
                    batch = full_points
                else:                                  # in-kernel sampler: replay the batch the kernel just drew
                    batch = self.sample(batch_size, cols, point_offset=0)
                xs = [batch[:, k:k + 1].clone().requires_grad_() for k in range(total)]
                for p in stray_params:
                    p.grad = None
                closs = solver._constraint_loss(nums, xs, criterion)
                closs.backward()
                self.out[loss_idx] += closs.detach()
            opt.step()
            _native.check(self.lib.pinn_record_loss(self.plan, C.c_void_p(self.out.data_ptr()),
                                                    C.c_void_p(ring_ptr), C.c_int64(ring_len),
                                                    C.c_void_p(self.step_counter.data_ptr()), self._stream()))

        capturable = bool(opt.param_groups and opt.param_groups[0].get('capturable', False))
        graphs_ok = (capturable or fused_adam) and not nums and os.environ.get('PYDENS_B200_NO_GRAPH') != '1'
        from .solver import _progress
        done = 0

        if host_sampler is None:
            # ---------------- points sampled in the kernel: replay graphs of `per_graph` steps ----------------
            per_graph = int(os.environ.get('PYDENS_B200_GRAPH_STEPS', '0')) or (4 if niters >= 64 else 1)
            key = ('dev', batch_size, local_n, cols_key, fused_nums, per_graph, adam_key)
            graph = self._graphs.get(key) if graphs_ok else None
            if graphs_ok and graph is None and niters >= _GRAPH_MIN_ITERS:
                if not opt_ready:
                    one_step()                        # real step 0: also materialises optimizer state
                    done = 1
                try:
                    torch.cuda.synchronize(self.device)

                    def body():
                        for _ in range(per_graph):
                            one_step()
                    # capture executes nothing, but a first launch of a freshly captured kernel sequence must not
                    # be preceded by a dry run: the capture itself is side-effect free
                    graph = capture_graph(self.device, body)
                    self._graphs[key] = graph
                except Exception as exc:              # capture unsupported here: plain launches  # noqa: BLE001
                    import warnings
                    warnings.warn('pydens_b200: CUDA-graph capture of the step failed (%s); using plain launches' % exc)
                    graph = None
                    torch.cuda.synchronize(self.device)
            if graph is not None:
                n_rep = (niters - done) // per_graph
                budget = ring_len - done
                for _ in range(n_rep):
                    if budget < per_graph:            # the ring is about to wrap: read it out first
                        torch.cuda.synchronize(self.device)
                        drain(done)
                        budget = ring_len
                    graph.replay()
                    done += per_graph
                    budget -= per_graph
            for _ in _progress(niters - done):
                if done - drained[0] >= ring_len:
                    torch.cuda.synchronize(self.device)
                    drain(done)
                one_step()
                done += 1
            self.steps_done = start + niters
            torch.cuda.synchronize(self.device)
            drain(niters)
        else:
            # ---------------- host batches: native pipeline, one call per step ----------------
            # staging depth = how many steps the host may run ahead of the GPU (absorbs host jitter; matters most
            # with several ranks, where the in-kernel all-reduce makes every rank wait for the slowest host)
            n_stage = max(2, min(8, int(os.environ.get('PYDENS_B200_STAGES', '4'))))
            pkey = (local_n, n_stage)
            pipe = self._pipes.get(pkey)
            if pipe is None:
                pipe = C.c_void_p()
                _native.check(self.lib.pinn_pipe_create(self.plan, n_stage, C.c_int64(local_n), C.byref(pipe)))
                self._pipes[pkey] = pipe
            bufs = [_RawPtr(self.lib.pinn_pipe_buffer(pipe, k)) for k in range(n_stage)]
            pinned = self._pinned_batches.get((batch_size, n_stage))
            loss_host = self._pinned_losses(niters)
            loss_ptr = loss_host.data_ptr()
            key = ('host', batch_size, local_n, n_stage, fused_nums, adam_key)
            execs = self._graphs.get(key) if graphs_ok else None
            exec_ptrs = [g.raw_cuda_graph_exec() for g in execs] if execs else None
            keep = [None] * n_stage                    # the tensors whose memory a copy in flight still reads
            off_bytes = point_offset * total * 4
            want_graphs = graphs_ok and execs is None and niters >= _HOST_GRAPH_MIN_ITERS

            for i in _progress(niters):
                k = i % n_stage
                batch = host_sampler.sample(batch_size)
                if isinstance(batch, torch.Tensor) and batch.dtype == torch.float32 and batch.is_contiguous() \
                        and batch.device.type == 'cpu':
                    src = batch                        # copied from where it lies (pinned or not)
                else:
                    if pinned is None:
                        pinned = [torch.empty((batch_size, total), dtype=torch.float32).pin_memory()
                                  for _ in range(n_stage)]
                        self._pinned_batches[(batch_size, n_stage)] = pinned
                    _native.check(self.lib.pinn_pipe_wait(pipe, k))          # its previous copy has left the buffer
                    if isinstance(batch, torch.Tensor):
                        pinned[k].copy_(batch.detach().to('cpu', torch.float32).reshape(batch_size, total))
                    else:
                        pinned[k].copy_(torch.from_numpy(np.ascontiguousarray(batch, dtype=np.float32)).reshape(batch_size, total))
                    src = pinned[k]
                keep[k] = src
                r = (start + i) % ring_len
                if exec_ptrs is not None:
                    _native.check(self.lib.pinn_pipe_step(pipe, k, C.c_void_p(src.data_ptr() + off_bytes),
                                                          C.c_void_p(exec_ptrs[k]), C.c_void_p(ring_ptr + 4 * r),
                                                          C.c_void_p(loss_ptr + 4 * i), stream))
                    continue
                _native.check(self.lib.pinn_pipe_step(pipe, k, C.c_void_p(src.data_ptr() + off_bytes), None, None, None, stream))
                full = None
                if nums:
                    full = src.to(self.device, non_blocking=True)
                    if dist is not None:
                        dist.broadcast(full, 0)
                one_step(bufs[k], full)
                _native.check(self.lib.pinn_pipe_finish(pipe, k, C.c_void_p(ring_ptr + 4 * r), C.c_void_p(loss_ptr + 4 * i), stream))
                # after the first (eager) step the optimizer state exists: capture the compute part of the step
                # once per staging buffer, so that every later step is ONE native call (copy -> graph -> loss read)
                if want_graphs:
                    want_graphs = False
                    try:
                        torch.cuda.synchronize(self.device)
                        graphs = []
                        for kk in range(n_stage):
                            graphs.append(capture_graph(self.device, lambda kk=kk: one_step(bufs[kk]),
                                                        pool=graphs[0].pool() if graphs else None))
                        self._graphs[key] = graphs
                        exec_ptrs = [g.raw_cuda_graph_exec() for g in graphs]
                    except Exception as exc:        # capture unsupported here: keep plain launches  # noqa: BLE001
                        import warnings
                        warnings.warn('pydens_b200: CUDA-graph capture of the step failed (%s); using plain launches' % exc)
                        exec_ptrs = None
                        torch.cuda.synchronize(self.device)
            self.steps_done = start + niters
            torch.cuda.synchronize(self.device)
            vals[:] = loss_host[:niters].numpy()
        if dist is not None and self.comm is None:
            # NCCL fallback: the cached graphs hold captured NCCL collectives, and NCCL's communicator teardown
            # (dist.destroy_process_group) waits for every such graph to be destroyed first — do not keep them
            self._graphs = {}
        if self.comm is not None:
            aborted = C.c_int(0)
            _native.check(self.lib.pinn_comm_status(self.comm, C.byref(aborted), 1))
            if aborted.value:
                raise RuntimeError('pydens_b200: a rank did not reach the in-kernel all-reduce within the time limit '
                                   '(PINN_COMM_TIMEOUT_S); the gradients of this fit were poisoned with NaN')
        solver.losses.extend(np.array(v, dtype=np.float32) for v in vals[:niters])


class _RawPtr:
    """ A device pointer with the one method `_step` needs from a tensor. """
    def __init__(self, ptr):
        self.ptr = int(ptr)

    def data_ptr(self):
        return self.ptr
