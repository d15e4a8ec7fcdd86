""" TEST INFRASTRUCTURE ONLY — CPU restatement of the pydens fit-step hot path on PyTorch autograd.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline / `--impl reference` legs may
import this module.  It is the checker (and the CPU thing that gets timed), never the product.

What it restates (reference = analysiscenter/pydens, `pydens/model_torch.py`):

    network      :164-172  batchflow Block == chain of nn.Linear / activation (layouts 'fa…f')
    ansatz       :107-128  boundary factor, time gate with trainable log_scale, initial condition
    D token      :174-178  autograd.grad(y.sum(), x, retain_graph=True, create_graph=True)[0]
    V token      :180-188  named trainable scalars
    loss         :448      MSELoss(residual, zeros)
    backward     :460      loss.backward()
    step         :419-422, :461  torch.optim.<name>(trainable params, lr)  +  optimizer.step()
    sampling     :431      one torch.rand((B, 1)) per column, in column order

The arithmetic is the reference's: the same ATen ops in the same order, so in fp32 it reproduces the
reference bit-for-bit on the same weights and points (pinned by tests/test_oracle.py against
tests/golden/*.npz, which oracle/make_golden.py produced from the unmodified reference).  With
dtype=torch.float64 it is the high-precision yardstick for the CUDA kernel.

Parameter order of every flat vector here (the engine's layout, include/pinn_b200.h):
    W_0 (row-major [out,in]), b_0, W_1, b_1, …, log_scale, V_0, V_1, …   then zero padding to 4.
"""
import math

import numpy as np
import torch


# ------------------------------------------------------------------------------------------------
# tokens
# ------------------------------------------------------------------------------------------------
def D(y, x):
    """ Differentiation token (model_torch.py:174-178). """
    return torch.autograd.grad(y.sum(), x, retain_graph=True, create_graph=True)[0]


class Problem:
    """ One PDE problem: network, ansatz configuration, equation; owns its parameters.

    equation(u, *xs, D=..., V=...) -> residual.  `V(name, init)` returns the named trainable scalar.
    """

    def __init__(self, equation, ndims, nparams=0, initial_condition=None, boundary_condition=None,
                 domain=(0.0, 1.0), features=(20, 30, 1), activation='Sigmoid', dtype=torch.float32,
                 variables=None, seed=0, layout=None):
        self.equation = equation
        self.ndims, self.nparams = ndims, nparams
        self.total = ndims + nparams
        self.has_ic = initial_condition is not None
        self.ndims_spatial = ndims - 1 if self.has_ic else ndims          # model_torch.py:25
        self.ic = initial_condition
        self.bc = boundary_condition
        if isinstance(domain[0], (int, float)):
            domain = [tuple(domain)] * ndims                               # model_torch.py:37-39
        self.domain = [tuple(map(float, d)) for d in domain]
        self.dtype = dtype
        self.activation = activation
        self.features = list(features)
        # batchflow layout letters: f dense, a activation, R save tensor, + add it back (model_torch.py:142-156)
        self.layout = (layout or 'fa' * (len(self.features) - 1) + 'f').replace(' ', '')
        gen = torch.Generator().manual_seed(seed)
        self.weights, self.biases = [], []
        n_in = self.total
        for n_out in self.features:                                       # nn.Linear default init
            bound = 1.0 / math.sqrt(n_in)
            w = (torch.rand((n_out, n_in), generator=gen) * 2 - 1) * bound
            b = (torch.rand((n_out,), generator=gen) * 2 - 1) * bound
            self.weights.append(w.to(dtype).requires_grad_())
            self.biases.append(b.to(dtype).requires_grad_())
            n_in = n_out
        self.log_scale = torch.zeros((), dtype=dtype, requires_grad=True)  # model_torch.py:50
        self.var_names = list(variables or {})
        self.vars = {k: torch.tensor([float(v)], dtype=dtype, requires_grad=True)
                     for k, v in (variables or {}).items()}

    # ---- flat parameter vector in the engine's layout ----
    def param_list(self):
        out = []
        for w, b in zip(self.weights, self.biases):
            out += [w, b]
        out.append(self.log_scale)
        out += [self.vars[k] for k in self.var_names]
        return out

    def n_params(self):
        return sum(p.numel() for p in self.param_list())

    def flat_params(self):
        flat = torch.cat([p.detach().reshape(-1) for p in self.param_list()])
        pad = (-flat.numel()) % 4
        return torch.cat([flat, flat.new_zeros(pad)])

    def load_flat(self, flat):
        flat = torch.as_tensor(flat)
        off = 0
        with torch.no_grad():
            for p in self.param_list():
                n = p.numel()
                p.copy_(flat[off:off + n].reshape(p.shape).to(p.dtype))
                off += n

    def flat_grads(self):
        parts = []
        for p in self.param_list():
            g = p.grad if p.grad is not None else torch.zeros_like(p)
            parts.append(g.detach().reshape(-1))
        flat = torch.cat(parts)
        pad = (-flat.numel()) % 4
        return torch.cat([flat, flat.new_zeros(pad)])

    def zero_grad(self):
        for p in self.param_list():
            p.grad = None

    # ---- model ----
    def net(self, xs_concat):
        h = xs_concat
        # `activation`: an nn.* name, a module class, or a sequence of those, one per 'a' (model_torch.py:150-151)
        n_a = self.layout.count('a')
        spec = list(self.activation) if isinstance(self.activation, (list, tuple)) else [self.activation] * n_a
        acts = [getattr(torch.nn, a)() if isinstance(a, str) else a() for a in spec]
        saved, i_f, i_a = [], 0, 0
        for letter in self.layout:
            if letter == 'f':
                h = torch.nn.functional.linear(h, self.weights[i_f], self.biases[i_f])
                i_f += 1
            elif letter == 'a':
                h = acts[i_a](h)
                i_a += 1
            elif letter == 'R':
                saved.append(h)
            elif letter == '+':
                h = h + saved.pop()
        return h

    def ansatz(self, u, xs_concat):
        """ model_torch.py:107-128 """
        nsp = self.ndims_spatial
        x_sp = xs_concat[:, :nsp]
        t = xs_concat[:, self.ndims - 1:self.ndims]
        lo = torch.tensor([d[0] for d in self.domain][:nsp], dtype=self.dtype).reshape(1, -1)
        hi = torch.tensor([d[1] for d in self.domain][:nsp], dtype=self.dtype).reshape(1, -1)
        t0 = self.domain[-1][0]
        if self.bc is not None:
            left = torch.prod((x_sp - lo) / (hi - lo), dim=1, keepdim=True)
            right = torch.prod((hi - x_sp) / (hi - lo), dim=1, keepdim=True)
            u = u * (left * right) + self.bc
        if self.has_ic:
            cols = [x_sp[:, i] for i in range(x_sp.shape[1])]
            if callable(self.ic):
                ic_val = self.ic(*cols)
            else:
                ic_val = torch.tensor(self.ic, dtype=torch.float32).to(self.dtype)
            gate = torch.sigmoid((t - t0) / torch.exp(self.log_scale)) - .5
            u = gate * u + ic_val.view(-1, 1)
        return u

    def V(self, name, init=None):
        return self.vars[name]

    # ---- one evaluation of residual / loss ----
    def residual(self, points):
        """ points: [B, total] tensor -> (residual [B,1], columns) with the autograd graph alive. """
        pts = torch.as_tensor(points, dtype=self.dtype)
        xs = [pts[:, i:i + 1].clone().requires_grad_() for i in range(self.total)]   # separate leaves, :435-436
        xs_concat = torch.cat(xs, dim=1)
        u = self.ansatz(self.net(xs_concat), xs_concat)
        r = self.equation(u, *xs, D=D, V=self.V)
        return r, xs

    def loss_and_grads(self, points, criterion=None):
        """ -> (loss float, residual [B] ndarray, flat grads tensor); `criterion(residual, zeros)` as in the
        reference's fit (:448), default MSELoss """
        self.zero_grad()
        r, xs = self.residual(points)
        loss = (criterion or torch.nn.functional.mse_loss)(r, torch.zeros_like(xs[0]))
        loss.backward()
        return float(loss.detach()), r.detach().reshape(-1).numpy().copy(), self.flat_grads()

    def predict(self, points):
        pts = torch.as_tensor(points, dtype=self.dtype)
        with torch.no_grad():
            return self.ansatz(self.net(pts), pts).reshape(-1).numpy().copy()


def default_points(batch_size, total, dtype=torch.float32):
    """ The reference's default sampler: one torch.rand((B,1)) per column, in order (:431). """
    return torch.cat([torch.rand((batch_size, 1)) for _ in range(total)], dim=1).to(dtype)


def fit(problem, niters, batch_size, lr=0.005, optimizer='Adam', point_stream=None, criterion=None, **opt_kwargs):
    """ The reference's training loop (:419-464) on the port.  `point_stream(i)` supplies the batch of
    iteration i ([B,total]); default: torch.rand per column like the reference.  Returns losses. """
    trainable = [p for p in problem.param_list() if p.requires_grad]
    opt = getattr(torch.optim, optimizer)(trainable, lr=lr, **opt_kwargs)
    losses = []
    for i in range(niters):
        opt.zero_grad()
        pts = point_stream(i) if point_stream is not None else default_points(batch_size, problem.total,
                                                                              problem.dtype)
        r, xs = problem.residual(pts)
        loss = (criterion or torch.nn.functional.mse_loss)(r, torch.zeros_like(xs[0]))
        loss.backward()
        opt.step()
        losses.append(float(loss.detach()))
    return np.asarray(losses)
