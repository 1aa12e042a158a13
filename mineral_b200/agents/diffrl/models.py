"""State-value critics of SHAC / SAPO with the reference's module structure and state_dict keys
(mineral/agents/diffrl/models.py:149-256): `Critic.critic_mlp.mlp.*`, `EnsembleCritic.critics.<k>.mlp.*`.
`forward` is the plain torch path used while rolling out (simulator-bound, out of scope); the regression in
`SHAC.update_critic` runs on `mineral_b200.nets.fused_towers.FusedCritics`."""

import math

import torch
import torch.nn as nn

from ...nets import MLP


def _orthogonal_(module, gain=math.sqrt(2.0)):
    """weight_init_orthogonal_ (models.py:10-14): orthogonal weights with gain sqrt(2), zero biases."""
    for m in module.modules():
        if isinstance(m, nn.Linear):
            nn.init.orthogonal_(m.weight.data, gain=gain)
            if m.bias is not None:
                nn.init.constant_(m.bias.data, 0.0)


class Critic(nn.Module):
    def __init__(self, state_dim, action_dim=None, mlp_kwargs=None, weight_init='orthogonal'):
        super().__init__()
        mlp_kwargs = dict(act_type='ELU', norm_type='LayerNorm') if mlp_kwargs is None else dict(mlp_kwargs)
        self.critic_mlp = MLP(state_dim, out_dim=1, plain_last=True, **mlp_kwargs)
        if weight_init == 'orthogonal':
            _orthogonal_(self)

    def mlps(self):
        return [self.critic_mlp]

    def forward(self, x, return_type=None):
        if isinstance(x, dict):
            x = x['z']
        x = self.critic_mlp(x)
        if return_type == 'all':
            return [x]
        if return_type == 'min_and_avg':
            return x, x.clone()
        return x


class EnsembleCritic(nn.Module):
    def __init__(self, state_dim, action_dim=None, n_critics=1, n_sample=None, with_vmap=False, mlp_kwargs=None,
                 weight_init='orthogonal'):
        super().__init__()
        if with_vmap:
            raise NotImplementedError('with_vmap stacks the ensemble for torch.vmap; FusedCritics batches the ensemble itself')
        mlp_kwargs = dict(act_type='ELU', norm_type='LayerNorm') if mlp_kwargs is None else dict(mlp_kwargs)
        self.n_critics, self.n_sample = n_critics, n_sample
        self.critics = nn.ModuleList([MLP(state_dim, out_dim=1, plain_last=True, **mlp_kwargs) for _ in range(n_critics)])
        if weight_init == 'orthogonal':
            _orthogonal_(self)

    def mlps(self):
        return list(self.critics)

    def forward(self, x, return_type='min'):
        if isinstance(x, dict):
            x = x['z']
        vs = torch.stack([c(x) for c in self.critics])
        if return_type == 'all':
            return vs
        if self.n_sample is not None:
            vs = vs[torch.randperm(self.n_critics)[:self.n_sample].tolist()]
        if return_type == 'min':
            return torch.min(vs, dim=0).values
        if return_type == 'min_and_avg':
            return torch.min(vs, dim=0).values, torch.mean(vs, dim=0)
        return vs
