"""ActorCritic parameter container + rollout-time forward.  Mirrors
mineral/agents/ppo/models.py:29-151 (same ctor kwargs, same sub-module names -> same state_dict
keys, same orthogonal initialisation gains).  Only flat 'obs' observation spaces are on this path
(dict/point-cloud encoders are out of scope)."""

import torch
import torch.nn as nn

from ...nets import MLP, Dist

_SQRT2 = 1.41421356237


def weight_init_orthogonal_(m):
    if isinstance(m, nn.Linear):
        nn.init.orthogonal_(m.weight, gain=_SQRT2)
        if getattr(m, 'bias', None) is not None:
            nn.init.zeros_(m.bias)


class ActorCritic(nn.Module):
    def __init__(self, obs_space, action_dim, mlp_kwargs=None, critic_mlp_kwargs=None, separate_value_mlp=True, fixed_sigma=True,
                 actor_dist_kwargs=None, weight_init='orthogonal', encoder=None, encoder_kwargs=None):
        super().__init__()
        if 'obs' not in obs_space:
            raise NotImplementedError("only a flat 'obs' observation is on the B200 path (MultiEncoder is out of scope)")
        mlp_kwargs = dict(units=[512, 256, 128], act_type='ELU') if mlp_kwargs is None else dict(mlp_kwargs)
        actor_dist_kwargs = dict(dist_type='normal') if actor_dist_kwargs is None else dict(actor_dist_kwargs)
        self.obs_space = obs_space
        self.separate_value_mlp = separate_value_mlp
        self.fixed_sigma = fixed_sigma
        self.encoder = nn.Identity()
        in_dim = obs_space['obs'][0]

        self.actor_mlp = MLP(in_dim, **mlp_kwargs)
        if separate_value_mlp:
            v_kwargs = {**mlp_kwargs, **critic_mlp_kwargs} if critic_mlp_kwargs is not None else mlp_kwargs
            self.value_mlp = MLP(in_dim, **v_kwargs)
        self.value = nn.Linear(self.value_mlp.out_dim if separate_value_mlp else self.actor_mlp.out_dim, 1)
        self.mu = nn.Linear(self.actor_mlp.out_dim, action_dim)
        if fixed_sigma:
            self.sigma = nn.Parameter(torch.zeros(action_dim, dtype=torch.float32), requires_grad=True)
        else:
            self.sigma = nn.Linear(self.actor_mlp.out_dim, action_dim)
        self.dist = Dist(**actor_dist_kwargs)
        self.weight_init = weight_init
        self.reset_parameters()

    def reset_parameters(self):
        """models.py:75-91: gain sqrt(2) towers, value head gain 1, policy heads gain 0.01."""
        assert self.weight_init == 'orthogonal'
        self.actor_mlp.apply(weight_init_orthogonal_)
        if self.separate_value_mlp:
            self.value_mlp.apply(weight_init_orthogonal_)
        # same call order as the reference so an equal torch seed yields equal initial weights
        self.value.apply(weight_init_orthogonal_)
        self.mu.apply(weight_init_orthogonal_)
        nn.init.orthogonal_(self.value.weight, gain=1.0)
        nn.init.orthogonal_(self.mu.weight, gain=0.01)
        if self.fixed_sigma:
            nn.init.constant_(self.sigma, 0)
        else:
            nn.init.orthogonal_(self.sigma.weight, gain=0.01)
            nn.init.zeros_(self.sigma.bias)

    def _actor_critic(self, obs_dict):
        """Pre-distribution outputs (mu_raw, logstd, value) (models.py:136-151); logstd is broadcast to
        (rows, A) exactly as the reference does (`mu * 0 + sigma`)."""
        z = obs_dict['obs']
        x = self.actor_mlp(z)
        mu = self.mu(x)
        sigma = self.sigma if self.fixed_sigma else self.sigma(x)
        if self.separate_value_mlp:
            x = self.value_mlp(z)
        return mu, mu * 0 + sigma, self.value(x)

    def forward(self, input_dict, **kwargs):
        """Reference-compatible eager forward (models.py:93-108): torch ops, for API parity and tests;
        PPO.train_epoch does not call it (it feeds `_actor_critic` to the fused loss kernel)."""
        prev_actions = input_dict.get('prev_actions', None)
        mu, logstd, value = self._actor_critic(input_dict)
        mu, sigma, distr = self.dist(mu, logstd)
        result = {'values': value, 'entropy': distr.entropy().sum(dim=-1), 'mu': mu, 'sigma': sigma, 'distr': distr}
        if prev_actions is not None:
            result['prev_neglogp'] = torch.squeeze(-distr.log_prob(prev_actions).sum(dim=-1))
        return result

    @torch.no_grad()
    def act(self, obs_dict, sample=True, **kwargs):
        """models.py:110-126."""
        mu, logstd, value = self._actor_critic(obs_dict)
        if not sample:
            return mu
        mu, sigma, distr = self.dist(mu, logstd)
        action = distr.sample()
        return {'neglogp': -distr.log_prob(action).sum(1), 'values': value, 'actions': action, 'mu': mu, 'sigma': sigma,
                'distr': distr}
