"""Actor and state-value critics of SHAC / SAPO with the reference's module structure and state_dict keys
(mineral/agents/diffrl/models.py:87-256): `Actor.actor_mlp.mlp.*` / `Actor.mu` / `Actor.sigma`,
`Critic.critic_mlp.mlp.*`, `EnsembleCritic.critics.<k>.mlp.*`.  They are PARAMETER CONTAINERS (checkpoint keys,
initialisation): `SHAC.update_critic` and the bootstrap values of `SHAC.collect_rollout` run on
`mineral_b200.nets.fused_towers.FusedCritics`; `forward` is the plain torch composition of the same modules, kept for
inspection and API compatibility (the actor's differentiable rollout through a simulator is out of scope)."""

import math

import torch
import torch.nn as nn

from ...nets import MLP, Dist


def _orthogonal_(module, gain=math.sqrt(2.0)):
    """weight_init_orthogonal_ (models.py:10-14): orthogonal weights with gain sqrt(2), zero biases."""
    for m in module.modules():
        if isinstance(m, nn.Linear):
            nn.init.orthogonal_(m.weight.data, gain=gain)
            if m.bias is not None:
                nn.init.constant_(m.bias.data, 0.0)


class Actor(nn.Module):
    """models.py:87-146 (orthogonal init; `sigma` a parameter initialised to `init_sigma`, or a Linear head)."""

    def __init__(self, state_dim, action_dim, fixed_sigma=True, init_sigma=-1.0, mlp_kwargs=None, dist_kwargs=None,
                 weight_init='orthogonal', weight_init_last_layers=False):
        super().__init__()
        mlp_kwargs = dict(norm_type='LayerNorm', act_type='ELU') if mlp_kwargs is None else dict(mlp_kwargs)
        dist_kwargs = dict(dist_type='normal') if dist_kwargs is None else dict(dist_kwargs)
        self.fixed_sigma, self.init_sigma = fixed_sigma, init_sigma
        self.actor_mlp = MLP(state_dim, **mlp_kwargs)
        self.mu = nn.Linear(self.actor_mlp.out_dim, action_dim)
        if fixed_sigma:
            self.sigma = nn.Parameter(torch.ones(action_dim, dtype=torch.float32), requires_grad=True)
        else:
            self.sigma = nn.Linear(self.actor_mlp.out_dim, action_dim)
        self.dist = Dist(**dist_kwargs)
        if weight_init != 'orthogonal':
            raise NotImplementedError(f'weight_init {weight_init}')
        _orthogonal_(self)
        if fixed_sigma:
            nn.init.constant_(self.sigma, init_sigma)
        elif weight_init_last_layers:
            nn.init.orthogonal_(self.mu.weight, gain=0.01)
            nn.init.orthogonal_(self.sigma.weight, gain=0.01)
            nn.init.zeros_(self.sigma.bias)

    def forward(self, x):
        if isinstance(x, dict):
            x = x['z']
        x = self.actor_mlp(x)
        mu = self.mu(x)
        sigma = self.sigma if self.fixed_sigma else self.sigma(x)
        return self.dist(mu, sigma)


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
