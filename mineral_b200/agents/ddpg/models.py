"""Actor / twin-Q modules of SAC and DDPG with the reference's structure and state_dict keys
(mineral/agents/ddpg/models.py:8-111): `Actor.actor_mlp.mlp.*`, `Actor.mu`, `Actor.sigma`, `EnsembleQ.critics.<k>.mlp.*`.
They are PARAMETER CONTAINERS: the agents run them through `mineral_b200.nets.fused_offpolicy` (tcgen05 / SIMT linear
kernels, LayerNorm+activation passes, the squashed-normal kernels of csrc/offpolicy.cu).  `forward` is the plain torch
composition of the same modules, kept for inspection and API compatibility (`actor(z) -> (mu, sigma, distr)`); neither
agent calls it."""

import math

import torch
import torch.nn as nn

from ...nets import MLP, Dist


def weight_init_orthogonal_(m):
    """mineral/agents/ddpg/utils.py:45-54 (Linear part)."""
    if isinstance(m, nn.Linear):
        nn.init.orthogonal_(m.weight.data)
        if hasattr(m.bias, 'data'):
            m.bias.data.fill_(0.0)


def weight_init_uniform_(m):
    """mineral/agents/ddpg/utils.py:57-79: fan-in variance scaling, uniform."""
    if isinstance(m, nn.Linear):
        fan_in = m.weight.shape[1]
        limit = math.sqrt(3.0 * 1.0 / max(1.0, fan_in))
        nn.init.uniform_(m.weight.data, -limit, limit)
        if hasattr(m.bias, 'data'):
            m.bias.data.fill_(0.0)


def _init(module, weight_init, last_linear=None):
    if weight_init is None:
        return
    if weight_init == 'orthogonal':
        module.apply(weight_init_orthogonal_)
    elif weight_init == 'uniform':
        module.apply(weight_init_uniform_)
        if last_linear is not None:
            nn.init.uniform_(last_linear.weight, -0.003, 0.003)
    else:
        raise NotImplementedError(weight_init)


class Actor(nn.Module):
    def __init__(self, state_dim, action_dim, tanh_policy=True, fixed_sigma=None, mlp_kwargs=None, dist_kwargs=None, weight_init=None):
        super().__init__()
        mlp_kwargs = dict(units=[512, 256, 128], act_type='ELU') if mlp_kwargs is None else dict(mlp_kwargs)
        self.tanh_policy, self.fixed_sigma = tanh_policy, fixed_sigma
        self.actor_mlp = MLP(state_dim, **mlp_kwargs)
        self.mu = nn.Linear(self.actor_mlp.out_dim, action_dim)
        if not tanh_policy:
            if fixed_sigma is None:
                pass
            elif fixed_sigma:
                self.sigma = nn.Parameter(torch.zeros(action_dim, dtype=torch.float32), requires_grad=True)
            else:
                self.sigma = nn.Linear(self.actor_mlp.out_dim, action_dim)
            self.dist = Dist(**(dist_kwargs or {}))
        self.weight_init = weight_init
        _init(self, weight_init, self.mu)

    def forward(self, x, std=None):
        if isinstance(x, dict):
            x = x['z']
        x = self.actor_mlp(x)
        mu = self.mu(x)
        if self.tanh_policy:  # DDPG
            return mu.tanh(), None, None
        if self.fixed_sigma is None:
            assert std is not None
            sigma = std
        elif self.fixed_sigma:
            sigma = self.sigma
        else:
            sigma = self.sigma(x)
        return self.dist(mu, sigma)


class EnsembleQ(nn.Module):
    def __init__(self, state_dim, action_dim, n_critics=2, mlp_kwargs=None, weight_init=None):
        super().__init__()
        mlp_kwargs = dict(units=[512, 256, 128], act_type='ELU') if mlp_kwargs is None else dict(mlp_kwargs)
        self.n_critics = n_critics
        self.critics = nn.ModuleList([MLP(state_dim + action_dim, out_dim=1, plain_last=True, **mlp_kwargs) for _ in range(n_critics)])
        self.weight_init = weight_init
        for critic in self.critics:
            _init(critic, weight_init, critic.mlp[-1])

    def forward(self, state, action):
        if isinstance(state, dict):
            state = state['z']
        x = torch.cat((state, action), dim=1)
        return [critic(x) for critic in self.critics]

    def get_q_min(self, state, action):
        return torch.min(torch.stack(self.forward(state, action)), dim=0).values

    def get_q_values(self, state, action):
        return self.forward(state, action)


class DistributionalEnsembleQ(nn.Module):
    """Categorical (C51) critics (mineral/agents/ddpg/models.py:117-170): every critic ends in `num_atoms` logits over the fixed
    support linspace(v_min, v_max, num_atoms).  Parameter container like the classes above; DDPG runs it through
    `FusedQ` (n_out = num_atoms) and the row-wise kernels of csrc/distl.cu."""

    def __init__(self, state_dim, action_dim, v_min=-10, v_max=10, num_atoms=51, n_critics=2, mlp_kwargs=None, weight_init=None):
        super().__init__()
        mlp_kwargs = dict(units=[512, 256, 128], act_type='ELU') if mlp_kwargs is None else dict(mlp_kwargs)
        self.v_min, self.v_max, self.num_atoms = v_min, v_max, num_atoms
        self.z_atoms = torch.linspace(v_min, v_max, num_atoms)
        self.n_critics = n_critics
        self.critics = nn.ModuleList([MLP(state_dim + action_dim, out_dim=num_atoms, plain_last=True, **mlp_kwargs) for _ in range(n_critics)])
        self.weight_init = weight_init
        if weight_init is not None:
            raise NotImplementedError(weight_init)   # models.py:152-154

    @property
    def distl(self):
        return True

    def forward(self, state, action):
        if isinstance(state, dict):
            state = state['z']
        x = torch.cat((state, action), dim=1)
        return [critic(x) for critic in self.critics]

    def get_q_values(self, state, action):
        return [torch.softmax(q, dim=1) for q in self.forward(state, action)]

    def get_q_min(self, state, action):
        qs = [torch.sum(q * self.z_atoms.to(q.device), dim=1) for q in self.get_q_values(state, action)]
        return torch.min(torch.stack(qs), dim=0).values
