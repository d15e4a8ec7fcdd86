""" Model side of the pydens API: network container, ansatz, D / V tokens.

Mirrors the public surface of pydens/model_torch.py:15-188 (`TorchModel`, `ConvBlockModel`, `D`, `V`,
`current_model`) so user code keeps working, but the modules here are parameter CONTAINERS: during
`Solver.fit` the fused CUDA step reads the parameters from one flat buffer the module's tensors are
views of.  `forward` / `anzatc` remain as a device-aware autograd implementation — it serves
constraints, custom equations the tracer cannot lower, and as the in-repo torch reference.
"""
from abc import ABC, abstractmethod
from contextvars import ContextVar

import torch
from torch import nn
from torch.autograd import grad

from . import tracer

current_model = ContextVar('current_model')
_tracing = ContextVar('pydens_b200_tracing', default=False)


class TorchModel(ABC, nn.Module):
    """ Base model: problem dimensions, boundary / initial condition binding (reference
    pydens/model_torch.py:17-128). """

    def __init__(self, ndims, initial_condition=None, boundary_condition=None, domain=(0, 1), nparams=0, **kwargs):
        _ = kwargs
        super().__init__()
        self.ndims = ndims
        self.nparams = nparams
        self.total = ndims + nparams
        self.ndims_spatial = ndims - 1 if initial_condition is not None else ndims
        self.variables = {}

        self.raw_initial_condition = initial_condition
        if initial_condition is None or callable(initial_condition):
            self.initial_condition = initial_condition
        else:
            value = float(initial_condition)
            self.initial_condition = lambda *args: torch.tensor(value, dtype=torch.float32)
        self.boundary_condition = boundary_condition

        if not isinstance(domain, (tuple, list)) or len(domain) == 0:
            raise ValueError('Should be either 1d or 2d-sequence of float/ints.')
        if isinstance(domain[0], (float, int)):
            domain = [tuple(domain)] * ndims
        elif not isinstance(domain[0], (tuple, list)):
            raise ValueError('Should be either 1d or 2d-sequence of float/ints.')
        self.domain = [tuple(lims) for lims in domain]

        # trainable time-scale of the initial-condition gate
        self.log_scale = nn.Parameter(torch.tensor(0.0))

    @abstractmethod
    def forward(self, xs):
        """ Network + ansatz on an (N x total) tensor. """

    def _set_trainable(self, layers, variables, flag):
        for name in (layers or []):
            for param in getattr(self, name).parameters():
                param.requires_grad = flag
        for name in (variables or []):
            getattr(self, name).requires_grad = flag

    def freeze_trainable(self, layers=None, variables=None):
        """ Exclude named sub-modules / variables from training (reference :56-82). """
        self._set_trainable(layers, variables, False)

    def unfreeze_trainable(self, layers=None, variables=None):
        """ Undo `freeze_trainable` (reference :84-105). """
        self._set_trainable(layers, variables, True)

    # spelling used by the reference README (README.md:126)
    freeze_layers = freeze_trainable
    unfreeze_layers = unfreeze_trainable

    def anzatc(self, u, xs):
        """ Bind boundary and initial conditions (reference :107-128):
            u <- u * prod_i (x_i-lo_i)(hi_i-x_i)/(hi_i-lo_i)^2 + bc
            u <- (sigmoid((t-t0)/exp(log_scale)) - 1/2) * u + ic(x_spatial)
        """
        nsp = self.ndims_spatial
        spatial = xs[:, :nsp]
        if self.boundary_condition is not None:
            lo = xs.new_tensor([lims[0] for lims in self.domain[:nsp]]).reshape(1, -1)
            hi = xs.new_tensor([lims[1] for lims in self.domain[:nsp]]).reshape(1, -1)
            width = hi - lo
            factor = (torch.prod((spatial - lo) / width, dim=1, keepdim=True)
                      * torch.prod((hi - spatial) / width, dim=1, keepdim=True))
            u = u * factor + self.boundary_condition
        if self.initial_condition is not None:
            t = xs[:, self.ndims - 1:self.ndims]
            t0 = self.domain[-1][0]
            gate = torch.sigmoid((t - t0) / torch.exp(self.log_scale)) - .5
            ic = self.initial_condition(*[spatial[:, i] for i in range(nsp)])
            if not isinstance(ic, torch.Tensor):
                ic = torch.as_tensor(ic, dtype=xs.dtype)
            u = gate * u + ic.to(xs.device).view(-1, 1)
        return u


# activations the fused kernel covers: type name (lower case) -> (kernel name, closed form).  A module is accepted
# only if it also BEHAVES like the closed form on a probe (nn.Softplus(beta=2), nn.GELU('tanh'), a user class
# that happens to be called Sin … go to the autograd path instead).
_ACT_NAMES = {
    'tanh': ('tanh', torch.tanh),
    'sigmoid': ('sigmoid', torch.sigmoid),
    'sin': ('sin', torch.sin),
    'softplus': ('softplus', nn.functional.softplus),
    'silu': ('silu', nn.functional.silu),
    'swish': ('silu', nn.functional.silu),
    'gelu': ('gelu', nn.functional.gelu),
}


def fused_activation_name(module):
    """ Kernel activation name of an activation module, or None if the fused kernel does not cover it. """
    entry = _ACT_NAMES.get(type(module).__name__.lower())
    if entry is None:
        return None
    probe = torch.linspace(-4.0, 4.0, 17, dtype=torch.float64).view(-1, 1)
    try:
        with torch.no_grad():
            got = module(probe)
    except Exception:                                    # pragma: no cover
        return None
    if got.shape != probe.shape or not torch.allclose(got, entry[1](probe), rtol=1e-9, atol=1e-12):
        return None
    return entry[0]


class Sin(nn.Module):
    def forward(self, x):
        return torch.sin(x)


def _make_activation(act):
    if isinstance(act, str):
        if act.lower() == 'sin':
            return Sin()
        return getattr(nn, act)()
    if isinstance(act, nn.Module):
        return act
    if isinstance(act, type) and issubclass(act, nn.Module):
        return act()
    if act is torch.sin:
        return Sin()
    raise ValueError('unknown activation %r' % (act,))


class DenseBlock(nn.Module):
    """ In-repo equivalent of the dense subset of batchflow's `Block` (the reference builds its network
    with `Block(inputs=..., layout=..., features=..., activation=...)`, model_torch.py:164-168):
    'f' dense layer, 'a' activation, 'R' … '+' residual sum.  Spaces in the layout are ignored.
    """

    def __init__(self, in_features, layout='fafaf', features=(20, 30, 1), activation='Sigmoid'):
        super().__init__()
        layout = layout.replace(' ', '')
        n_dense, n_act = layout.count('f'), layout.count('a')
        features = list(features)
        if len(features) != n_dense:
            raise ValueError('layout %r has %d dense layers but %d sizes were given' % (layout, n_dense, len(features)))
        acts = list(activation) if isinstance(activation, (list, tuple)) else [activation] * n_act
        if len(acts) < n_act:
            raise ValueError('not enough activations for layout %r' % layout)
        self.layout = layout
        self.ops = nn.ModuleList()
        self.kinds = []
        i_f = i_a = 0
        width = in_features
        for letter in layout:
            if letter == 'f':
                self.ops.append(nn.Linear(width, features[i_f]))
                width = features[i_f]
                i_f += 1
            elif letter == 'a':
                self.ops.append(_make_activation(acts[i_a]))
                i_a += 1
            elif letter in 'R+':
                self.ops.append(nn.Identity())
            else:
                raise ValueError('layout letter %r is not supported (dense layouts only)' % letter)
            self.kinds.append(letter)
        self.out_features = width

    @property
    def linears(self):
        return [op for op, k in zip(self.ops, self.kinds) if k == 'f']

    def forward(self, x):
        skips = []
        for op, kind in zip(self.ops, self.kinds):
            if kind == 'R':
                skips.append(x)
            elif kind == '+':
                x = x + skips.pop()
            else:
                x = op(x)
        return x

    def dense_chain(self):
        """ [(linear, activation-name or 'none', skip_src or None), …] if the block is a dense chain the fused
        kernel covers, else None.  Residual connections are covered in their usual form — both 'R' and '+'
        directly after an activation ('faR fa fa+ f'): `skip_src` is the index of the dense layer whose
        activated output is added to this layer's activated output. """
        chain, stack = [], []
        kinds, ops = self.kinds, list(self.ops)
        i = 0
        while i < len(kinds):
            kind = kinds[i]
            if kind == 'f':
                lin, act = ops[i], 'none'
                i += 1
                if i < len(kinds) and kinds[i] == 'a':
                    act = fused_activation_name(ops[i])
                    if act is None:
                        return None
                    i += 1
                chain.append([lin, act, None])
            elif kind in 'R+':
                if not chain or kinds[i - 1] not in ('a', '+') or (kind == '+' and kinds[i - 1] != 'a'):
                    return None                       # skip from the raw input / around a bare dense layer
                if kind == 'R':
                    stack.append(len(chain) - 1)      # saves the layer's final output (activation + its own skip)
                else:
                    if not stack or chain[-1][2] is not None:
                        return None
                    src = stack.pop()
                    if src == len(chain) - 1 or chain[src][0].out_features != chain[-1][0].out_features:
                        return None
                    if any(c[2] == src for c in chain):
                        return None                   # one consumer per saved tensor
                    chain[-1][2] = src
                i += 1
            else:
                return None
        if stack:
            return None
        return [tuple(c) for c in chain]


class ConvBlockModel(TorchModel):
    """ Fully-connected network configured by `layout` / `features` (`units` accepted as in the
    reference README.md:42) / `activation` (reference :130-172). """

    def __init__(self, ndims, initial_condition=None, boundary_condition=None, domain=(0, 1), nparams=0,
                 layout='fafaf', features=(20, 30, 1), activation='Sigmoid', units=None, **kwargs):
        super().__init__(ndims=ndims, initial_condition=initial_condition, boundary_condition=boundary_condition,
                         domain=domain, nparams=nparams, **kwargs)
        n_dense = layout.replace(' ', '').count('f')
        sizes = list(features)
        if units is not None and (len(list(units)) == n_dense or len(sizes) != n_dense):
            sizes = list(units)
        self.conv_block = DenseBlock(self.total, layout=layout, features=sizes, activation=activation)

    def forward(self, xs):
        return self.anzatc(self.conv_block(xs), xs)

    def __getattr__(self, name):
        # 'fc1', 'fc2', … address the dense layers (the names the reference README freezes, README.md:126)
        if name.startswith('fc') and name[2:].isdigit():
            linears = nn.Module.__getattr__(self, 'conv_block').linears
            idx = int(name[2:]) - 1
            if 0 <= idx < len(linears):
                return linears[idx]
        return nn.Module.__getattr__(self, name)


def D(y, x):
    """ Differentiation token.  On tensors: `autograd.grad` of the per-point sum (reference :174-178);
    on traced symbols: symbolic differentiation feeding the fused kernel's jet set. """
    if isinstance(y, tracer.Sym) or isinstance(x, tracer.Sym):
        return tracer.sym_D(y, x)
    return grad(y.sum(), x, retain_graph=True, create_graph=True)[0]


def V(name, *args, **kwargs):
    """ Token for a trainable variable, created on first use on the current model (reference :180-188). """
    model = current_model.get()
    if _tracing.get():
        if not hasattr(model, name):
            raise tracer.NotLowerable('variable %r is created outside the equation' % name)
        if getattr(model, name).numel() != 1:
            raise tracer.NotLowerable('variable %r is not a scalar' % name)
        return tracer.Sym(tracer.var(name))
    if not hasattr(model, name):
        param = nn.Parameter(*args, **kwargs)
        ref = next(model.parameters())
        param.data = param.data.to(ref.device)
        setattr(model, name, param)
    return getattr(model, name)
