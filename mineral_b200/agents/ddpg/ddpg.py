"""DDPG / TD3 on the B200 kernels.  Mirrors mineral/agents/ddpg/ddpg.py:20-345: same constructor, config keys
(`full_cfg.agent.{network,ddpg}` as in mineral/cfgs/agent/AllegroHandDDPG.yaml) and methods (`get_actions /
get_tgt_policy_actions / explore_env / train / update_net / update_critic / update_actor / set_train / set_eval`).

Reference line -> here (no autograd, no cuBLAS, no torch elementwise op on the update path):
  ddpg.py:98-117 get_actions (+ exploration noise)      -> actor tower GEMMs (tcgen05), mnr_tanh_action
  ddpg.py:132-147 target policy smoothing (TD3)         -> mnr_tanh_action with noise / output bounds
  ddpg.py:244-249 sample_batch + normalize x2           -> one gather into the padded critic-input buffers
  ddpg.py:291-297 target + critic regression            -> target towers, mnr_td_target_stats, critic towers fwd/bwd,
                                                           mnr_critic_mse_sum_fwd_bwd, mnr_fused_clip_adam
  ddpg.py:299-310 actor step                            -> critic fwd + input gradient only, mnr_actor_seed, mnr_tanh_action_bwd,
                                                           actor bwd, mnr_fused_clip_adam
  ddpg.py:260-263 soft_update                           -> mnr_soft_update over flat buffers
  ddpg.py:267-290 categorical critic (DistributionalEnsembleQ) -> n_atoms-wide critic heads on the same towers, mnr_distl_target
                                                           (softmax + C51 projection + ensemble min), mnr_distl_bce_fwd_bwd,
                                                           mnr_distl_actor_seed (csrc/distl.cu)"""

import os
from copy import deepcopy

import torch

from ... import ops
from ...nets.fused_offpolicy import FusedActor, FusedQ, flatten_like
from ..agent import cfg_get
from ..sac.sac import OffPolicyAgent, make_optim
from . import models
from .schedule_util import ExponentialSchedule, LinearSchedule


class DDPG(OffPolicyAgent):
    r"""Deep Deterministic Policy Gradients."""

    cfg_key = 'ddpg'

    def __init__(self, full_cfg, **kwargs):
        self._init_common(full_cfg, kwargs)
        self.ddpg_config = cfg = self.algo_config

        ActorCls = getattr(models, self.network_config.actor)
        CriticCls = getattr(models, self.network_config.critic)
        self.actor = ActorCls(self.obs_dim, self.action_dim, **cfg_get(self.network_config, 'actor_kwargs', {})).to(self.device)
        self.critic = CriticCls(self.obs_dim, self.action_dim, **cfg_get(self.network_config, 'critic_kwargs', {})).to(self.device)
        if not self.actor.tanh_policy:
            raise NotImplementedError('DDPG.get_actions raises for a stochastic actor (ddpg.py:104-105)')

        self.critic_target = deepcopy(self.critic)
        self.actor_target = deepcopy(self.actor) if not cfg.no_tgt_actor else self.actor
        self.actor_optim = make_optim(cfg.optim_type, self.actor.parameters(), cfg_get(cfg, 'actor_optim_kwargs', {}))
        self.critic_optim = make_optim(cfg.optim_type, self.critic.parameters(), cfg_get(cfg, 'critic_optim_kwargs', {}))
        self.critic_target_flat = flatten_like(self.critic_target, self.critic_optim)
        self.actor_target_flat = flatten_like(self.actor_target, self.actor_optim) if not cfg.no_tgt_actor else None

        self._init_buffers()

        # --- exploration-noise schedule (ddpg.py:82-96) ---
        noise = cfg.noise
        if noise.decay == 'linear':
            self.noise_scheduler = LinearSchedule(start_val=noise.std_max, end_val=noise.std_min, total_iters=noise.lin_decay_iters)
        elif noise.decay == 'exp':
            self.noise_scheduler = ExponentialSchedule(start_val=noise.std_max, gamma=cfg.exp_decay_rate, end_val=noise.std_min)
        else:
            self.noise_scheduler = None

        B, dev = self.batch_size, self.device
        self._fa = FusedActor(self.actor, max(B, self.num_actors), dev)
        self._fa_t = FusedActor(self.actor_target, B, dev) if not cfg.no_tgt_actor else self._fa
        self._fq = FusedQ(self.critic, B, dev)
        self._fq_t = FusedQ(self.critic_target, B, dev)
        self._attach_optimizer_splits()
        self.distl = bool(getattr(self.critic, 'distl', False))   # ddpg.py:267
        if self.distl:
            n = self.critic.num_atoms
            self._z_atoms = self.critic.z_atoms.to(device=dev, dtype=torch.float32).contiguous()
            self._target_dist = torch.empty(B, n, dtype=torch.float32, device=dev)
            self._distl_ws = ops.distl_workspace(dev)
        elif self._fq.n_out != 1:
            raise NotImplementedError('a scalar-Q critic must end in a 1-wide layer')

    # --- noise draws (hooks; torch.normal(0, std) == std * standard normal) ---
    def _exploration_noise(self, rows):
        """ddpg.py:106-116, noise.py:22-55: Normal(0, std) per element; 'mixed' = per-env std from linspace(std_min, std_max)."""
        noise = self.ddpg_config.noise
        eps = self._randn((rows, self.action_dim))
        if noise.type == 'fixed':
            return eps * self.get_noise_std()
        if noise.type == 'mixed':
            std_seq = torch.linspace(noise.std_min, noise.std_max, rows).to(self.device).unsqueeze(-1)
            return eps * std_seq
        raise NotImplementedError(noise.type)

    def _target_noise(self, rows):
        pre = self._pre
        if pre is not None:
            return pre['eps'][pre['i'], 0]
        return self._randn((rows, self.action_dim)) * self.ddpg_config.noise.tgt_pol_std

    def get_noise_std(self):
        return self.ddpg_config.noise.std_max if self.noise_scheduler is None else self.noise_scheduler.val()

    def update_noise(self):
        if self.noise_scheduler is not None:
            self.noise_scheduler.step()

    # ------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def get_actions(self, obs, sample=True):
        """ddpg.py:98-117: tanh(mu(norm(obs))) [+ exploration noise, clamped to [-1, 1]]."""
        x = obs['obs']
        rows = x.shape[0]
        self._normalize_into(x, self.xn, rows)
        fa = self._fa
        fa.forward(self.xn, self.x_ld, rows, split=self._actor_dirty)
        self._actor_dirty = False
        actions = torch.empty(rows, self.action_dim, dtype=torch.float32, device=self.device)
        ops.tanh_action(fa.mu_raw, self._exploration_noise(rows) if sample else None, actions, None, (-1.0, 1.0), rows)
        return actions

    @torch.no_grad()
    def get_tgt_policy_actions(self, obs=None, sample=True):
        """ddpg.py:131-147 on the normalised next observations in xn[:, :O]: actions land in xn[:, O:O+A] (TD3 smoothing:
        noise clamped to +-tgt_pol_noise_bound, result to [-1, 1])."""
        B, O, A = self.batch_size, self.obs_dim, self.action_dim
        if obs is not None:
            ops.copy_cols(obs.contiguous(), self.xn[:B, :O])
        fa = self._fa_t
        own = fa is self._fa
        fa.forward(self.xn, self.x_ld, B, split=self._actor_dirty if own else True)
        if own:
            self._actor_dirty = False
        nb = self.ddpg_config.noise.tgt_pol_noise_bound
        ops.tanh_action(fa.mu_raw, self._target_noise(B) if sample else None, self.xn[:B, O:O + A], (-nb, nb), (-1.0, 1.0), B)
        return self.xn[:B, O:O + A]

    # ------------------------------------------------------------------------------------------------
    def train(self):
        """ddpg.py:201-237."""
        self._train_prologue()
        cfg = self.ddpg_config
        while self.agent_steps < self.max_agent_steps:
            self.epoch += 1
            self.set_eval()
            trajectory, steps = self.explore_env(self.env, cfg.horizon_len, sample=True)
            self.agent_steps += steps
            self.memory.add_to_buffer(trajectory)
            self.set_train()
            results = self.update_net(self.memory)
            metrics = {k: torch.mean(torch.stack(v)).item() for k, v in results.items()}
            metrics.update({'epoch': self.epoch, 'mini_epoch': self.mini_epoch})
            metrics = {f'train_stats/{k}': v for k, v in metrics.items()}
            metrics.update(self._episode_metrics())
            self.writer.add(self.agent_steps, metrics)
            self.writer.write()
        self.save(os.path.join(self.ckpt_dir, 'final.pth'))

    noise_draws = 1   # per mini-epoch: the target-policy smoothing noise

    def _draw_noise(self, buf):
        buf.normal_(0.0, float(self.ddpg_config.noise.tgt_pol_std))

    def _mini_epoch(self, results, do_actor, do_targets):
        """One iteration of the loop in DDPG.update_net (ddpg.py:241-263) on the batch already gathered into xc / xn."""
        critic_loss, critic_grad_norm = self.update_critic()
        results['loss/critic'].append(critic_loss)
        results['grad_norm/critic'].append(critic_grad_norm)
        if do_actor:
            actor_loss, actor_grad_norm = self.update_actor()
            results['loss/actor'].append(actor_loss)
            results['grad_norm/actor'].append(actor_grad_norm)
        if do_targets:
            self._soft_updates()

    @torch.no_grad()
    def update_critic(self):
        """ddpg.py:266-297 (MSE critics) on the batch in xc / xn / reward_b / done_b."""
        cfg, B = self.ddpg_config, self.batch_size
        self.get_tgt_policy_actions()
        q_t = self._fq_t.forward(self.xn, B)
        out = torch.empty(2, dtype=torch.float32, device=self.device)
        if self.distl:   # ddpg.py:275-290,295-296: projected target distribution, BCE against every critic's softmax
            ct, n = self.critic_target, self.critic.num_atoms
            ops.distl_target(q_t, self.reward_b.view(-1), self.done_b.view(-1), self._z_atoms, cfg.gamma ** cfg.nstep, float(ct.v_min),
                             float(ct.v_max), n, self._target_dist)
            fq = self._fq
            fq.forward(self.xc, B, split=self._split_needed())
            ops.distl_bce_fwd_bwd(fq.pred, self._target_dist, fq.g_pred, out[0:1], self._distl_ws, B, n)
            fq.backward_from_seeds(self.xc, B)
            self.optimizer_update(self.critic_optim, out[1:2])
            return out[0], out[1]
        ops.td_target_stats([q.view(-1)[:B] for q in q_t], None, self.reward_b.view(-1), self.done_b.view(-1), None, 0.0,
                            cfg.gamma ** cfg.nstep, self.target_q.view(-1), None)
        self._fq.loss_fwd_bwd(self.xc, self.target_q, out[0:1], split=self._split_needed())
        self.optimizer_update(self.critic_optim, out[1:2])
        return out[0], out[1]

    @torch.no_grad()
    def update_actor(self):
        """ddpg.py:299-310: actor_loss = -mean(min_k Q_k(s, tanh(mu(s)))) through the frozen critics."""
        B, O, A = self.batch_size, self.obs_dim, self.action_dim
        fa, fq = self._fa, self._fq
        fa.forward(self.xc, self.x_ld, B, split=self._actor_dirty)
        ops.tanh_action(fa.mu_raw, None, self.xc[:B, O:O + A], None, None, B)
        fq.forward(self.xc, B, split=self._split_needed())
        out = torch.empty(5, dtype=torch.float32, device=self.device)
        if self.distl:   # Q = min_k E_k[z] (models.py:160-163); seeds through the minimal critic's softmax
            ops.distl_actor_seed(fq.pred, self._z_atoms, fq.g_pred, out[0:1], self._distl_ws, B, self.critic.num_atoms)
        else:
            ops.actor_seed([q.view(-1)[:B] for q in fq.pred], [g.view(-1)[:B] for g in fq.g_pred], None, None, 0.0, 0.0, out[0:4], None)
        dx = fq.input_grad(self.xc, B)
        ops.tanh_action_bwd(fa.mu_raw, dx[0][:B, O:O + A], fa.g_mu, B, g_actions2=dx[1][:B, O:O + A] if len(dx) > 1 else None)
        fa.backward(self.xc, B)
        self.optimizer_update(self.actor_optim, out[4:5])
        self._actor_dirty = self._split_needed()
        return out[0], out[4]

    def eval(self):
        raise NotImplementedError  # ddpg.py:312-313

    def set_train(self):
        for m in (self.actor, self.critic, self.actor_target, self.critic_target):
            m.train()

    def set_eval(self):
        for m in (self.actor, self.critic, self.actor_target, self.critic_target):
            m.eval()

    def save(self, f):
        pass  # ddpg.py:341-342

    def load(self, f, ckpt_keys=''):
        pass  # ddpg.py:344-345
