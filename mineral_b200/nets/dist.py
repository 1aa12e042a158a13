"""Action-distribution head.  Mirrors mineral/nets/dist.py:9-58: 'normal' / 'dreamerv3_normal' (PPO) and the squashed
variants of the SAC / SAPO actors ('squashed_normal', 'dreamerv3_squashed_normal', mineral/nets/distributions.py:10-58).
This module is the torch-side description of the head (rollout-time `ActorCritic.act`, inspection); during the updates
the same maths runs inside the kernels (csrc/ppo_loss.cu for PPO, mnr_squashed_sample / mnr_squashed_bwd in
csrc/offpolicy.cu for SAC)."""

import math

import numpy as np
import torch
import torch.distributions as D
import torch.nn as nn
import torch.nn.functional as F

_KINDS = ('normal', 'dreamerv3_normal', 'squashed_normal', 'dreamerv3_squashed_normal')


class _Tanh(D.Transform):
    """tanh bijector with the numerically stable log-det of the reference (distributions.py:10-37)."""
    domain = D.constraints.real
    codomain = D.constraints.interval(-1.0, 1.0)
    bijective = True
    sign = +1

    def __init__(self, cache_size=1):
        super().__init__(cache_size=cache_size)

    def __eq__(self, other):
        return isinstance(other, _Tanh)

    def _call(self, x):
        return x.tanh()

    def _inverse(self, y):
        return 0.5 * (y.log1p() - (-y).log1p())

    def log_abs_det_jacobian(self, x, y):
        return 2.0 * (math.log(2.0) - x - F.softplus(-2.0 * x))


class SquashedNormal(D.TransformedDistribution):
    def __init__(self, loc, scale, validate_args=None):
        self.loc, self.scale = loc, scale
        super().__init__(D.Normal(loc, scale), [_Tanh()], validate_args=validate_args)

    @property
    def mean(self):
        return self.loc.tanh()


class Dist(nn.Module):
    def __init__(self, dist_type='normal', minstd=1.0, maxstd=1.0, minlogstd=None, maxlogstd=None, validate_args=None):
        super().__init__()
        if dist_type not in _KINDS:
            raise NotImplementedError(dist_type)
        self.dist_type = dist_type
        self.minstd = float(np.exp(minlogstd)) if minlogstd is not None else minstd
        self.maxstd = float(np.exp(maxlogstd)) if maxlogstd is not None else maxstd
        self.minlogstd, self.maxlogstd = minlogstd, maxlogstd
        self.validate_args = validate_args

    def forward(self, mu, logstd):
        if self.dist_type == 'normal':
            sigma = torch.exp(logstd)
            return mu, sigma, D.Normal(mu, sigma, validate_args=self.validate_args)
        if self.dist_type == 'squashed_normal':
            if self.minlogstd is not None or self.maxlogstd is not None:
                logstd = torch.clamp(logstd, self.minlogstd, self.maxlogstd)
            sigma = logstd.exp()
            return mu, sigma, SquashedNormal(mu, sigma, validate_args=self.validate_args)
        sigma = (self.maxstd - self.minstd) * torch.sigmoid(logstd + 2.0) + self.minstd
        if self.dist_type == 'dreamerv3_normal':
            mu = torch.tanh(mu)
            return mu, sigma, D.Normal(mu, sigma, validate_args=self.validate_args)
        return mu, sigma, SquashedNormal(mu, sigma, validate_args=self.validate_args)

    def __repr__(self):
        return f'Dist(dist_type={self.dist_type}, minstd={self.minstd}, maxstd={self.maxstd})'
