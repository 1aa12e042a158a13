"""SHAC / SAPO: the critic-target and critic-regression half of mineral/agents/diffrl/shac.py.

On the B200 path: the (T,B) rollout buffers (`create_buffers`,
shac.py:209-224), `compute_target_values` / `compute_target_values_with_entropy` (shac.py:877-916) as
one scan kernel, `CriticDataset` and the critic minibatch regression loop (`update_critic`,
shac.py:824-875).  The actor update back-propagates through a differentiable simulator
(shac.py:632-822), which is third-party and not in the reference repo: `update_actor` raises.
"""

import collections

import torch
import torch.nn as nn
import torch.nn.functional as F

from ... import ops
from ...nets import MLP
from .utils import CriticDataset, grad_norm


class Critic(nn.Module):
    """State-value MLP (mineral/agents/diffrl/models.py:149-178): `forward(z, return_type)`."""

    def __init__(self, in_dim, units=(64, 64), act_type='ELU', norm_type=None, num_critics=1):
        super().__init__()
        self.critics = nn.ModuleList([MLP(in_dim, out_dim=1, units=list(units), act_type=act_type, norm_type=norm_type,
                                          plain_last=True) for _ in range(num_critics)])

    def forward(self, z, return_type='min'):
        vs = [c(z) for c in self.critics]
        if return_type == 'all':
            return vs
        return torch.min(torch.stack(vs), dim=0).values


class SHAC:
    def __init__(self, num_envs, num_obs, num_actions, horizon_len, device, gamma=0.99, lam=0.95, critic_method='td-lambda',
                 critic_iterations=16, num_critic_batches=4, critic_units=(64, 64), num_critics=1, critic_lr=5e-4,
                 max_grad_norm=1.0, truncate_grads=True, with_logprobs=False, use_distr_ent=True, entropy_coef=None,
                 offset_by_target_entropy=False, scale_by_target_entropy=False, target_entropy_scalar=1.0):
        self.device = device
        self.num_envs, self.num_obs, self.num_actions = num_envs, num_obs, num_actions
        self.action_dim = num_actions
        self.obs_space = {'obs': (num_obs,)}
        self.horizon_len = horizon_len
        self.gamma, self.lam, self.critic_method = gamma, lam, critic_method
        self.critic_iterations, self.num_critic_batches = critic_iterations, num_critic_batches
        self.critic_batch_size = num_envs * horizon_len // num_critic_batches  # shac.py:73
        self.max_grad_norm, self.truncate_grads = max_grad_norm, truncate_grads
        self.with_logprobs = with_logprobs
        self.use_distr_ent = use_distr_ent
        self.entropy_coef = entropy_coef
        self.with_autoent = False
        self.offset_by_target_entropy = offset_by_target_entropy
        self.scale_by_target_entropy = scale_by_target_entropy
        self.target_entropy = -self.action_dim * target_entropy_scalar  # shac.py:184-185
        self.critic = Critic(num_obs, critic_units, num_critics=num_critics).to(device)
        self.encoder = lambda obs: obs['obs']
        self.critic_optim = torch.optim.AdamW(self.critic.parameters(), lr=critic_lr, betas=(0.7, 0.95))
        self.create_buffers(horizon_len, num_envs)

    def create_buffers(self, T, B):
        """shac.py:209-224 (done_mask is FLOAT here)."""
        f = dict(dtype=torch.float32, device=self.device)
        self.obs_buf = {k: torch.zeros((T, B) + v, **f) for k, v in self.obs_space.items()}
        self.action_buf = torch.zeros((T, B, self.action_dim), **f)
        self.rew_buf = torch.zeros((T, B), **f)
        self.done_mask = torch.zeros((T, B), **f)
        self.next_values = torch.zeros((T, B), **f)
        self.avg_next_values = torch.zeros((T, B), **f)
        self.target_values = torch.zeros((T, B), **f)
        self.ret = torch.zeros((B), **f)
        self.mus = torch.zeros((T, B, self.num_actions), **f)
        self.sigmas = torch.zeros((T, B, self.num_actions), **f)
        if self.with_logprobs:
            self.logprobs = torch.zeros((T, B), **f)
            self.distr_ent = torch.zeros((T, B), **f)

    @torch.no_grad()
    def compute_target_values(self):
        """shac.py:877-891: one-step or TD-lambda targets, one launch, bit-exact."""
        ops.critic_targets(self.rew_buf, self.done_mask, self.next_values, self.gamma, getattr(self, 'lam', 0.95),
                           self.critic_method, out=self.target_values)

    def get_alpha(self, scalar=True):
        return self.entropy_coef

    @torch.no_grad()
    def compute_target_values_with_entropy(self):
        """shac.py:893-916 (entropy-augmented reward), same kernel with the entropy term enabled."""
        ent = self.distr_ent if self.use_distr_ent else self.logprobs
        alpha = self.get_alpha(scalar=True) if self.with_autoent else self.entropy_coef
        ops.critic_targets(self.rew_buf, self.done_mask, self.next_values, self.gamma, getattr(self, 'lam', 0.95),
                           self.critic_method, alpha=alpha, entropy=ent, negate_entropy=not self.use_distr_ent,
                           offset_by_target_entropy=self.offset_by_target_entropy,
                           scale_by_target_entropy=self.scale_by_target_entropy, target_entropy=self.target_entropy,
                           out=self.target_values)

    def compute_critic_loss(self, obs, target_v):
        """shac.py:869-875: mean over critics of the MSE."""
        z = self.encoder(obs)
        pred_vs = self.critic(z, return_type='all')
        return torch.stack([F.mse_loss(p.squeeze(-1), target_v) for p in pred_vs]).mean()

    def update_critic(self, dataset=None):
        """shac.py:824-867: critic_iterations x num_critic_batches regression steps."""
        if dataset is None:
            dataset = CriticDataset(self.critic_batch_size, self.obs_buf, self.target_values)
        results = collections.defaultdict(list)
        for _ in range(self.critic_iterations):
            total, before, after = 0.0, [], []
            for i in range(len(dataset)):
                b_obs, b_t = dataset[i]
                self.critic_optim.zero_grad()
                loss = self.compute_critic_loss(b_obs, b_t)
                loss.backward()
                for p in self.critic.parameters():
                    p.grad.nan_to_num_(0.0, 0.0, 0.0)
                if self.truncate_grads:
                    before.append(grad_norm(self.critic.parameters()))
                    if self.max_grad_norm is not None:
                        nn.utils.clip_grad_norm_(self.critic.parameters(), self.max_grad_norm)
                    after.append(grad_norm(self.critic.parameters()))
                self.critic_optim.step()
                total = total + loss.detach()
            results['value_loss'].append(total / len(dataset))
            if before:
                results['grad_norm_before_clip/critic'].append(torch.mean(torch.stack(before)))
                results['grad_norm_after_clip/critic'].append(torch.mean(torch.stack(after)))
        return results

    def update_actor(self):
        raise NotImplementedError('SHAC actor BPTT needs a differentiable simulator (dflex/warp), which is not part of the '
                                  'reference repository (SURVEY.md section 2 row 6)')
