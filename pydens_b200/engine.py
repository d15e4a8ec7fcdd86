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
_HOST_GRAPH_MIN_ITERS = 64          # host-sampler mode: one capture per staging buffer


def _dist():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return dist
    return None


def capture_graph(device, body, pool=None):
    """ Capture `body()` into a CUDA graph on a side stream.  Unlike the `torch.cuda.graph` context manager this
    does not run the garbage collector and does not empty the allocator cache, which is what makes a capture
    cost ~20 ms there; a fit only pays a millisecond or two here. """
    graph = torch.cuda.CUDAGraph()
    current = torch.cuda.current_stream(device)
    side = torch.cuda.Stream(device=device)
    side.wait_stream(current)
    with torch.cuda.stream(side):
        try:
            if pool is not None:
                graph.capture_begin(pool=pool)
            else:
                graph.capture_begin()
            body()
            graph.capture_end()
        except Exception:
            try:
                graph.capture_end()
            except Exception:                       # noqa: BLE001
                pass
            raise
    current.wait_stream(side)
    return graph


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
        self.steps_done = 0
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
            if getattr(self, 'comm', None):
                self.lib.pinn_comm_destroy(self.comm)
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
            C.c_float(1.0 / entry['n']), C.c_void_p(entry['out'].data_ptr()), None,
            C.c_void_p(entry['workspace'].data_ptr()), C.c_size_t(entry['workspace'].numel()), self._stream()))
        self.out.add_(entry['out'])                       # [grads | loss] += the constraint's

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
    def fit(self, niters, batch_size, sampler, loss_terms, optimizer, criterion, lr, **kwargs):
        solver = self.solver
        solver._make_optimizer(optimizer, lr, fused_hint=True, **kwargs)
        opt = solver.optimizer
        nums = solver._constraint_numbers(loss_terms)
        fused_constraints = [e for e in (self._constraint_plan(n) for n in nums) if e is not None]
        nums = [n for n in nums if self._constraint_plan(n) is None]     # the rest is added by autograd
        dist = _dist()
        world = dist.get_world_size() if dist is not None else 1
        rank = dist.get_rank() if dist is not None else 0
        local_n, point_offset = shard_batch(batch_size, world, rank)
        if local_n <= 0:
            raise ValueError('batch_size %d is smaller than the number of ranks %d' % (batch_size, world))
        inv_n = 1.0 / float(batch_size)
        total = solver.model.total

        # every parameter's .grad must be the view of `out` (a previous zero_grad may have dropped it)
        for p, o in self.entries:
            if p.grad is None or p.grad.data_ptr() != self.out.data_ptr() + 4 * o:
                p.grad = self.out[o:o + p.numel()].view(p.shape)

        cols, host_sampler = None, None
        if sampler is not None:
            dev_cols = sampler.device_columns() if hasattr(sampler, 'device_columns') else None
            if dev_cols is not None and len(dev_cols) == total and os.environ.get('PYDENS_B200_HOST_SAMPLER') != '1':
                try:
                    cols = _native.make_columns(dev_cols, total)
                except ValueError:                    # more mixture components than the kernel takes
                    cols = None
            if cols is None:
                host_sampler = sampler
        solver.model.train()

        ring = torch.zeros(max(niters, 1), dtype=torch.float32, device=self.device)
        start = self.steps_done
        loss_idx = self.n_params
        zero_host = None

        if host_sampler is not None:
            # staging depth = how many steps the host may run ahead of the GPU (absorbs host jitter; matters most
            # with several ranks, where the in-kernel all-reduce makes every rank wait for the slowest host)
            n_stage = max(2, int(os.environ.get('PYDENS_B200_STAGES', '4')))
            pinned = [torch.empty((batch_size, total), dtype=torch.float32).pin_memory() for _ in range(n_stage)]
            events = [torch.cuda.Event() for _ in range(n_stage)]          # H2D of buffer k has completed
            free_ev = [torch.cuda.Event() for _ in range(n_stage)]         # compute no longer reads dev_pts[k]
            copy_stream = torch.cuda.Stream(device=self.device)
            d2h_stream = torch.cuda.Stream(device=self.device)
            dev_pts = [torch.empty((local_n, total), dtype=torch.float32, device=self.device) for _ in range(n_stage)]
            zero_host = torch.zeros(max(niters, 1), dtype=torch.float32).pin_memory()

        def one_step(i, points=None):
            self._step(points, cols, local_n, inv_n, point_offset, allreduce=dist is not None)
            if dist is not None and self.comm is None:
                dist.all_reduce(self.out)              # no peer-memory path: NCCL sums [grads | loss]
            for entry in fused_constraints:            # after the all-reduce: every rank adds the same term
                self._constraint_step(entry)
            if nums:
                xs = solver._sample_host(sampler if host_sampler is not None else None, batch_size) \
                    if points is None else [points[:, k:k + 1] for k in range(total)]
                closs = solver._constraint_loss(nums, xs, criterion)
                closs.backward()
                self.out[loss_idx] += closs.detach()
            opt.step()
            _native.check(self.lib.pinn_record_loss(self.plan, C.c_void_p(self.out.data_ptr()),
                                                    C.c_void_p(ring.data_ptr()), C.c_int64(ring.numel()),
                                                    C.c_void_p(self.step_counter.data_ptr()), self._stream()))

        capturable = bool(opt.param_groups and opt.param_groups[0].get('capturable', False))
        use_graph = (host_sampler is None and not nums and capturable and niters >= _GRAPH_MIN_ITERS
                     and os.environ.get('PYDENS_B200_NO_GRAPH') != '1')
        done = 0
        if use_graph:
            one_step(0)                               # real step 0: also materialises optimizer state
            done = 1
            # several steps per graph when there are many: one replay = G optimizer steps, which matters in
            # the launch-bound small-batch regime (README example: batch 100 x 1500 iterations)
            per_graph = int(os.environ.get('PYDENS_B200_GRAPH_STEPS', '0')) or (4 if niters - 1 >= 64 else 1)
            graph = None
            try:
                torch.cuda.synchronize(self.device)

                def body():
                    for k in range(per_graph):
                        one_step(1 + k)
                graph = capture_graph(self.device, body)
            except Exception:                         # capture unsupported here: plain launches
                graph = None
                torch.cuda.synchronize(self.device)
            if graph is not None:
                for _ in range((niters - 1) // per_graph):
                    graph.replay()
                done = 1 + ((niters - 1) // per_graph) * per_graph
                del graph
        if done < niters:
            from .solver import _progress
            stage_graphs = None
            for i in _progress(niters - done):
                if host_sampler is not None:
                    k = i % len(pinned)
                    events[k].synchronize()
                    batch = host_sampler.sample(batch_size)
                    if isinstance(batch, torch.Tensor):
                        src = batch if batch.dtype == torch.float32 else batch.float()
                    else:
                        pinned[k].copy_(torch.from_numpy(np.ascontiguousarray(batch, dtype=np.float32)))
                        src = pinned[k]
                    # the batch travels on its own stream so that the copy of step i+1 overlaps step i
                    cur = torch.cuda.current_stream(self.device)
                    with torch.cuda.stream(copy_stream):
                        copy_stream.wait_event(free_ev[k])
                        dev_pts[k].copy_(src[point_offset:point_offset + local_n], non_blocking=True)
                        events[k].record(copy_stream)
                    cur.wait_event(events[k])
                    if stage_graphs is not None:
                        stage_graphs[k].replay()
                    else:
                        one_step(done + i, dev_pts[k])
                    free_ev[k].record(cur)
                    # the step's loss goes to the host every step (as in the reference, :464) — from the device
                    # ring and on its own stream, so that the read-back never sits between two steps
                    with torch.cuda.stream(d2h_stream):
                        d2h_stream.wait_event(free_ev[k])
                        r = (start + done + i) % ring.numel()
                        zero_host[done + i:done + i + 1].copy_(ring[r:r + 1], non_blocking=True)
                    # after the first (eager) step the optimizer state exists: capture the compute part of the
                    # step once per staging buffer, so that every later step is copy -> replay -> loss read
                    if (i == 0 and stage_graphs is None and capturable and not nums and niters - done >= _HOST_GRAPH_MIN_ITERS
                            and os.environ.get('PYDENS_B200_NO_GRAPH') != '1'):
                        try:
                            torch.cuda.synchronize(self.device)
                            graphs = []
                            for kk in range(len(pinned)):
                                graphs.append(capture_graph(self.device, lambda kk=kk: one_step(0, dev_pts[kk]),
                                                            pool=graphs[0].pool() if graphs else None))
                            stage_graphs = graphs
                        except Exception:               # capture unsupported here: keep plain launches
                            stage_graphs = None
                            torch.cuda.synchronize(self.device)
                else:
                    one_step(done + i)
            stage_graphs = None
        self.steps_done = start + niters
        torch.cuda.synchronize(self.device)
        if host_sampler is not None:
            vals = zero_host.numpy().copy()
        else:
            idx = (start + np.arange(niters)) % max(niters, 1)
            vals = ring.cpu().numpy()[idx]
        solver.losses.extend(np.array(v, dtype=np.float32) for v in vals[:niters])
