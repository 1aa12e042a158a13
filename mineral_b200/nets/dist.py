"""Action-distribution head.  Mirrors mineral/nets/dist.py:9-58 for the two PPO distributions on
the accelerated path ('normal', 'dreamerv3_normal'); the squashed variants belong to SAC/SAPO actors
(next tier) and raise.  Used at rollout time (`ActorCritic.act`); during the
update the same maths runs inside the fused loss kernel (csrc/ppo_loss.cu)."""

import numpy as np
import torch
import torch.distributions as D
import torch.nn as nn


class Dist(nn.Module):
    def __init__(self, dist_type='normal', minstd=1.0, maxstd=1.0, minlogstd=None, maxlogstd=None, validate_args=None):
        super().__init__()
        if dist_type not in ('normal', 'dreamerv3_normal'):
            raise NotImplementedError(f'dist_type {dist_type!r} is not on the B200 PPO path')
        self.dist_type = dist_type
        self.minstd = float(np.exp(minlogstd)) if minlogstd is not None else minstd
        self.maxstd = float(np.exp(maxlogstd)) if maxlogstd is not None else maxstd
        self.minlogstd, self.maxlogstd = minlogstd, maxlogstd
        self.validate_args = validate_args

    def forward(self, mu, logstd):
        if self.dist_type == 'normal':
            sigma = torch.exp(logstd)
        else:  # dreamerv3_normal
            sigma = (self.maxstd - self.minstd) * torch.sigmoid(logstd + 2.0) + self.minstd
            mu = torch.tanh(mu)
        return mu, sigma, D.Normal(mu, sigma, validate_args=self.validate_args)

    def __repr__(self):
        return f'Dist(dist_type={self.dist_type}, minstd={self.minstd}, maxstd={self.maxstd})'
