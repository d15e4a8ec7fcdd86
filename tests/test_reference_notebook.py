""" Drop-in check of the public API: the reference's own test (`pydens/tests/pydens_test.py:15-39`) converts its
tutorial notebook to a script and `exec`s it.  The same here, against THIS package through the `pydens` alias
(`from pydens import Solver, D, V, ConvBlockModel, NumpySampler`), with `niters` cut down and plotting mocked.
The notebook itself stays in the reference checkout: the test is skipped where /root/reference is absent
(the GPU box); nothing is copied into the repository. """
import glob
import json
import os
import re
import sys
import types
from unittest import mock

import pytest

NOTEBOOKS = sorted(glob.glob('/root/reference/tutorials/*.ipynb'))


@pytest.mark.skipif(not NOTEBOOKS, reason='reference checkout not present')
@pytest.mark.parametrize('path', NOTEBOOKS or ['<absent>'])
def test_tutorial_notebook_runs_unchanged_on_this_package(path, monkeypatch):
    monkeypatch.setenv('PYDENS_B200_PROGRESS', '0')
    cells = [''.join(c['source']) for c in json.load(open(path))['cells'] if c['cell_type'] == 'code']
    plt = mock.MagicMock()
    plt.subplots.return_value = (mock.MagicMock(), mock.MagicMock())
    bft = types.ModuleType('batchflow.models.torch')
    bft.Block = bft.MultiLayer = object                      # imported by the notebook, never used
    fake = {'matplotlib': mock.MagicMock(pyplot=plt), 'matplotlib.pyplot': plt, 'batchflow': types.ModuleType('batchflow'),
            'batchflow.models': types.ModuleType('batchflow.models'), 'batchflow.models.torch': bft}
    saved = {k: sys.modules.get(k) for k in fake}           # (mock.patch.dict would also unload what the cells import)
    sys.modules.update(fake)
    try:
        import pydens
        assert os.path.dirname(os.path.abspath(pydens.__file__)).startswith(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        created = []
        real_init = pydens.Solver.__init__

        def recording_init(self, *args, **kwargs):
            real_init(self, *args, **kwargs)
            created.append(self)
        monkeypatch.setattr(pydens.Solver, '__init__', recording_init)
        env, n_fits = {}, 0
        for i, cell in enumerate(cells):
            code = '\n'.join(l for l in cell.split('\n') if not l.lstrip().startswith(('%', '!')))
            n_fits += len(re.findall(r'\.fit\(', code))
            code = re.sub(r'niters=\d+', 'niters=3', code)
            exec(compile(code, 'cell %d' % i, 'exec'), env)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    assert n_fits >= 6 and len(created) >= 5
    # every problem of the tutorial lowers to the fused kernel (what runs on a B200), constraints included
    for sv in created:
        assert sv._traced is not None, sv._lower_error
        for num in range(len(sv.constraints)):
            assert sv._lower_constraint(num) is not None
    solver = env['solver']
    assert len(solver.losses) > 0 and all(l == l for l in solver.losses)      # finite losses from the last problem


@pytest.mark.skipif(not os.path.exists('/root/reference/README.md'), reason='reference checkout not present')
def test_readme_snippets_run_unchanged_on_this_package(monkeypatch):
    """ The reference README's python blocks, in order (README.md:25-128; `D` / `V` are used there without being
    imported, so they are supplied), `niters` cut down: `units=` alias, `NumpySampler` algebra, `V` in the initial
    condition, a constraint, `freeze_layers` with the README's layer names. """
    monkeypatch.setenv('PYDENS_B200_PROGRESS', '0')
    text = open('/root/reference/README.md').read()
    blocks = re.findall(r'```python\n(.*?)```', text, flags=re.S)
    assert len(blocks) >= 5
    import pydens
    env = {'D': pydens.D, 'V': pydens.V}
    solvers = []
    for i, block in enumerate(blocks):
        exec(compile(re.sub(r'niters=\d+', 'niters=3', block), 'README block %d' % i, 'exec'), env)
        if 'solver' in env and (not solvers or solvers[-1] is not env['solver']):
            solvers.append(env['solver'])
    assert len(solvers) == 3
    for sv in solvers:
        assert sv._traced is not None, sv._lower_error
    last = solvers[-1]
    assert len(last.losses) == 6 and last._lower_constraint(0) is not None
    frozen = [n for n, p in last.model.named_parameters() if not p.requires_grad]
    assert frozen and all(('conv_block' in n) or n == 'log_scale' for n in frozen)
    assert last.model.init.requires_grad


@pytest.mark.skipif(not os.path.exists('/root/reference/pydens/model_torch.py'), reason='reference checkout not present')
def test_reshape_and_concat_equals_the_reference_on_random_inputs():
    """ `Solver.reshape_and_concat` (reference model_torch.py:328-362) decides how `predict` and constraints read
    numbers / arrays / lists / tensors: same output as the reference's own classmethod on random argument mixes. """
    import numpy as np
    import torch
    here = os.path.dirname(os.path.abspath(__file__))
    standin = os.path.join(os.path.dirname(here), 'oracle', 'batchflow_standin')
    saved_path, saved_mods = list(sys.path), {k: sys.modules.get(k) for k in ('pydens', 'pydens.model_torch', 'batchflow')}
    for k in list(sys.modules):
        if k == 'pydens' or k.startswith('pydens.') or k == 'batchflow' or k.startswith('batchflow.'):
            sys.modules.pop(k)
    sys.path[:0] = [standin, '/root/reference']
    try:
        import pydens as ref
        assert ref.__file__.startswith('/root/reference')
        ref_fn = ref.Solver.reshape_and_concat
    finally:
        sys.path[:] = saved_path
        for k in list(sys.modules):
            if k == 'pydens' or k.startswith('pydens.') or k == 'batchflow' or k.startswith('batchflow.'):
                sys.modules.pop(k)
        for k, v in saved_mods.items():
            if v is not None:
                sys.modules[k] = v
    from pydens_b200 import Solver
    rng = np.random.RandomState(0)
    for case in range(200):
        n = int(rng.choice([1, 2, 5, 16]))
        args = []
        for _ in range(int(rng.randint(1, 5))):
            kind = int(rng.randint(7))
            if kind == 0:
                args.append(float(np.round(rng.uniform(-3, 3), 3)))
            elif kind == 1:
                args.append(int(rng.randint(-3, 4)))
            elif kind == 2:
                args.append(rng.uniform(-1, 1, size=n).astype(np.float32))
            elif kind == 3:
                args.append(rng.uniform(-1, 1, size=(n, 1)))
            elif kind == 4:
                args.append(list(np.round(rng.uniform(-1, 1, size=n), 3)))
            elif kind == 5:
                args.append(torch.tensor(rng.uniform(-1, 1, size=n), dtype=torch.float32))
            else:
                args.append(torch.tensor(rng.uniform(-1, 1, size=(n, 1)), dtype=torch.float32))
        try:
            want = ref_fn(list(args))
        except Exception as exc:                            # the reference rejects the mix: so must we
            with pytest.raises(Exception):
                Solver.reshape_and_concat(list(args))
            continue
        got = Solver.reshape_and_concat(list(args))
        assert tuple(got.shape) == tuple(want.shape), (case, [type(a).__name__ for a in args])
        assert torch.allclose(got.to(torch.float64), want.to(torch.float64), rtol=1e-6, atol=1e-7), case
