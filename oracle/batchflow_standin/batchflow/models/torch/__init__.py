""" Stand-in for `batchflow.models.torch.Block`, restricted to dense layouts ('f', 'a', 'R', '+').

Call site being served: pydens/model_torch.py:164-168
    Block(inputs=fake_inputs, layout=..., features=[...], activation=..., **user_kwargs)
"""
from torch import nn


class Block(nn.Module):
    def __init__(self, inputs=None, layout='', features=None, units=None, activation='Sigmoid', **kwargs):
        super().__init__()
        _ = kwargs
        layout = layout.replace(' ', '')
        n_dense = layout.count('f')
        sizes = None
        for cand in (units, features):       # README.md:42 spells it `units`, the code `features`
            if cand is not None and len(list(cand)) == n_dense:
                sizes = list(cand)
                break
        if sizes is None:
            raise ValueError('number of dense layers in layout does not match features/units')
        acts = activation if isinstance(activation, (list, tuple)) else [activation] * layout.count('a')
        in_features = inputs.shape[1]
        layers, i_f, i_a = [], 0, 0
        for letter in layout:
            if letter == 'f':
                layers.append(nn.Linear(in_features, sizes[i_f]))
                in_features = sizes[i_f]
                i_f += 1
            elif letter == 'a':
                act = acts[i_a]
                i_a += 1
                layers.append(getattr(nn, act)() if isinstance(act, str) else act())
            elif letter in 'R+':                      # residual: R saves the tensor, + adds it back
                layers.append(nn.Identity())
            else:
                raise ValueError('stand-in Block supports only dense layouts, got %r' % letter)
        self.layers = nn.ModuleList(layers)
        self.letters = layout

    def forward(self, x):
        saved = []
        for letter, layer in zip(self.letters, self.layers):
            if letter == 'R':
                saved.append(x)
            elif letter == '+':
                x = x + saved.pop()
            else:
                x = layer(x)
        return x
