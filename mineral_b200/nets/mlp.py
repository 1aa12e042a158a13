"""MLP tower container.  Mirrors mineral/nets/mlp.py:25-77 closely enough that `state_dict()` keys
(`mlp.<i>.weight` ...) and layer order (Linear -> [Dropout] -> [Norm] -> [Act]) are identical, so
reference checkpoints load.  It is a PARAMETER CONTAINER (initialisation, checkpoint keys): the agents run
these layers on the hand-written kernels (nets/fused_towers.py); `forward` is the plain torch composition,
kept for inspection and for the tests' autograd comparisons."""

import torch.nn as nn


def _norm(norm_type, size, kwargs):
    return getattr(nn, norm_type)(size, **kwargs)


def _act(act_type, kwargs):
    return getattr(nn.modules.activation, act_type)(**kwargs)


class MLP(nn.Module):
    def __init__(self, in_dim, out_dim=None, units=(512, 256, 128), dropout=None, dropout_kwargs=None, where_dropout='every',
                 norm_type=None, norm_kwargs=None, act_type='ReLU', act_kwargs=None, bias=True, plain_last=False):
        super().__init__()
        units = list(units) + ([out_dim] if out_dim is not None else [])
        dropout_kwargs = dict(inplace=False) if dropout_kwargs is None else dropout_kwargs
        norm_kwargs = {} if norm_kwargs is None else norm_kwargs
        act_kwargs = dict(inplace=True) if act_kwargs is None else act_kwargs
        self.in_dim, self.out_dim, self.units = in_dim, units[-1], units
        self.norm_type, self.act_type = norm_type, act_type

        last = len(units) - 1
        layers, width = [], in_dim
        for i, u in enumerate(units):
            layers.append(nn.Linear(width, u, bias=bias))
            width = u
            if plain_last and i == last:
                break
            if dropout is not None:
                first, final = i == 0, i == last
                if (first and where_dropout in ('every', 'first')) or (final and where_dropout in ('every', 'last')) or (
                        not first and not final and where_dropout == 'every'):
                    layers.append(nn.Dropout(dropout, **dropout_kwargs))
            if norm_type is not None:
                layers.append(_norm(norm_type, u, norm_kwargs))
            if act_type is not None:
                layers.append(_act(act_type, act_kwargs))
        self.mlp = nn.Sequential(*layers)

    def forward(self, x):
        return self.mlp(x)
