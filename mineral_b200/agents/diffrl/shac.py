"""SHAC / SAPO on the B200 kernels: everything of mineral/agents/diffrl/shac.py that sits on the rollout->update
path, behind the reference's agent interface.

  * `SHAC(Agent)` with the reference's constructor (`SHAC(full_cfg, logdir=, accelerator=, datasets=, env=)`,
    shac.py:30-207), config keys (`full_cfg.agent.{network,shac}`, mineral/cfgs/agent/DFlexAnt{SHAC,SHAC2,SAPO}.yaml),
    normaliser configurations (shac.py:77-90), target networks, `save` / `load` with the reference's checkpoint
    keys (shac.py:969-1005);
  * the observation-normaliser protocol of the rollout (shac.py:639-654,694-700): every env step updates the LIVE
    `obs_rms` while the step's observation is normalised with a `deepcopy` SNAPSHOT taken at rollout start
    (`rollout_begin` / `rollout_observe`), and the critic-side bookkeeping of the rollout (`collect_rollout`:
    obs_buf / rew_buf / done_mask / next_values incl. early-termination and time-out bootstrap, shac.py:662-798);
  * `compute_target_values` / `compute_target_values_with_entropy` (shac.py:877-916): one scan kernel;
  * `CriticDataset` + `update_critic` (shac.py:824-875): critic MLP forward / backward (tcgen05 or SIMT linear layers +
    LayerNorm/activation passes), MSE + gradient, grad.nan_to_num_ + clip_grad_norm_ + Adam in one launch, the whole
    critic_iterations x num_critic_batches loop replayed as ONE CUDA graph;
  * the Polyak target update (shac.py:429-434): one launch over the flat parameter buffers.

The actor update back-propagates through a differentiable simulator (shac.py:504-630, 800-822), which is third-party and
not part of the reference repository: `update_actor` / `train` raise, and `collect_rollout` takes the policy as a callable.
A critic architecture `FusedCritics` cannot run is an error at construction (no torch-module / autograd path).
"""

import collections
import re
from copy import deepcopy

import torch
import torch.nn as nn

from ... import ops
from ...common import normalizers
from ...common.reward_shaper import RewardShaper
from ...nets.fused_offpolicy import flatten_like
from ...nets.fused_towers import FusedCritics
from ..agent import Agent, Tracker, cfg_get
from ..ppo.ppo import FusedAdam
from . import models
from .utils import CriticDataset


class _ShapeEnv:
    """Stand-in for the simulator when only shapes are known (`SHAC.from_shapes`): the attributes shac.py:38-43 reads."""

    def __init__(self, num_envs, num_obs, num_actions, max_episode_length=1000):
        import types
        self.num_envs, self.num_obs, self.num_actions, self.max_episode_length = num_envs, num_obs, num_actions, max_episode_length
        self.observation_space = types.SimpleNamespace(shape=(num_obs,))
        self.action_space = types.SimpleNamespace(shape=(num_actions,))


class SHAC(Agent):
    r"""Short-Horizon Actor-Critic (critic half + rollout bookkeeping)."""

    def __init__(self, full_cfg, **kwargs):
        self.network_config = full_cfg.agent.network
        self.shac_config = cfg = full_cfg.agent.shac
        self.num_actors = cfg.num_actors
        self.max_agent_steps = int(cfg.max_agent_steps)
        super().__init__(full_cfg, **kwargs)
        net = self.network_config

        self.num_envs, self.num_obs, self.num_actions = self.env.num_envs, self.env.num_obs, self.env.num_actions
        self.max_episode_length = getattr(self.env, 'max_episode_length', 1000)

        # --- SAPO parameters (shac.py:45-55) ---
        self.with_logprobs = cfg_get(cfg, 'with_logprobs', False)
        self.with_autoent = cfg_get(cfg, 'with_autoent', False)
        self.entropy_coef = cfg_get(cfg, 'entropy_coef', None)
        self.offset_by_target_entropy = cfg_get(cfg, 'offset_by_target_entropy', False)
        self.scale_by_target_entropy = cfg_get(cfg, 'scale_by_target_entropy', False)
        self.unscale_entropy_alpha = cfg_get(cfg, 'unscale_entropy_alpha', False)
        self.use_distr_ent = cfg_get(cfg, 'use_distr_ent', False)
        self.entropy_in_return = cfg_get(cfg, 'entropy_in_return', False)
        self.entropy_in_targets = cfg_get(cfg, 'entropy_in_targets', False)

        # --- SHAC parameters (shac.py:57-74) ---
        self.tanh_clamp = cfg_get(net, 'tanh_clamp', False)
        self.gamma = cfg_get(cfg, 'gamma', 0.99)
        self.critic_method = cfg_get(cfg, 'critic_method', 'one-step')
        if self.critic_method == 'td-lambda':
            self.lam = cfg_get(cfg, 'lambda', 0.95)
        self.critic_iterations = cfg_get(cfg, 'critic_iterations', 16)
        self.target_critic_alpha = cfg_get(cfg, 'target_critic_alpha', 0.4)
        self.no_target_critic = bool(cfg_get(cfg, 'no_target_critic', False))
        self.horizon_len = cfg.horizon_len
        self.max_epochs = cfg_get(cfg, 'max_epochs', 0)
        self.num_critic_batches = cfg_get(cfg, 'num_critic_batches', 4)
        self.critic_batch_size = self.num_envs * self.horizon_len // self.num_critic_batches
        self.max_grad_norm = cfg_get(cfg, 'max_grad_norm', None)
        self.truncate_grads = bool(cfg_get(cfg, 'truncate_grads', False))

        # --- Normalizers (shac.py:76-90): legacy DiffRL statistics under tanh_clamp, PPO-style f64 otherwise ---
        if self.tanh_clamp:
            rms_config = dict(eps=1e-5, correction=0, initial_count=1e-4, dtype=torch.float32)
        else:
            rms_config = dict(eps=1e-5, initial_count=1, dtype=torch.float64)
        if self.normalize_input:
            rms = {}
            for k, v in self.obs_space.items():
                rms[k] = normalizers.RunningMeanStd(v, **rms_config) if re.match(self.obs_rms_keys, k) else normalizers.Identity()
            self.obs_rms = nn.ModuleDict(rms).to(self.device)
        else:
            self.obs_rms = None
        self._obs_rms_snapshot = None

        # --- Encoder (shac.py:92-110): flat observations only ---
        if cfg_get(net, 'encoder', None) is not None:
            raise NotImplementedError('encoders are outside the rollout->update hot path (SURVEY.md section 2 row 18)')
        self.encoder = _ObsEncoder()
        self.share_encoder = cfg_get(cfg, 'share_encoder', True)
        self.actor_encoder = self.encoder if self.share_encoder else deepcopy(self.encoder)
        obs_dim = self.obs_space['obs'][0]
        assert obs_dim == self.num_obs and self.action_dim == self.num_actions

        # --- Model (shac.py:112-127) ---
        ActorCls = getattr(models, cfg_get(net, 'actor', 'Actor'))
        critic_name = cfg_get(net, 'critic', 'Critic')
        if critic_name not in ('Critic', 'EnsembleCritic'):
            raise NotImplementedError(f'critic {critic_name}')
        CriticCls = getattr(models, critic_name)
        self.actor = ActorCls(obs_dim, self.action_dim, **(cfg_get(net, 'actor_kwargs', {}) or {})).to(self.device)
        self.critic = CriticCls(obs_dim, self.action_dim, **(cfg_get(net, 'critic_kwargs', {}) or {})).to(self.device)
        if FusedCritics.parse(self.critic.mlps()) is None:
            raise NotImplementedError('critic architecture outside the fused path: 1-2 state-value MLPs of equal shape, '
                                      'Linear [-> LayerNorm] -> ELU | ReLU | SiLU blocks and a plain 1-wide last layer')

        # --- Optim (shac.py:129-150): flat-buffer Adam / AdamW with device-resident scalars ---
        if cfg.optim_type not in ('Adam', 'AdamW'):
            raise NotImplementedError(f'optim_type {cfg.optim_type}: only Adam / AdamW have a fused kernel')
        decoupled = cfg.optim_type == 'AdamW'
        akw, ckw = dict(cfg_get(cfg, 'actor_optim_kwargs', {}) or {}), dict(cfg_get(cfg, 'critic_optim_kwargs', {}) or {})
        for kw in (akw, ckw):
            if 'betas' in kw:
                kw['betas'] = tuple(kw['betas'])
        self.actor_optim = FusedAdam(self.actor.parameters(), decoupled=decoupled, **akw)
        self.critic_optim = FusedAdam(self.critic.parameters(), decoupled=decoupled, **ckw)
        self.actor_lr, self.critic_lr = self.actor_optim.param_groups[0]['lr'], self.critic_optim.param_groups[0]['lr']
        self.min_lr, self.max_lr = cfg_get(cfg, 'min_lr', 1e-5), cfg_get(cfg, 'max_lr', self.actor_lr)
        self.last_lr = self.actor_lr

        # --- Target networks (shac.py:163-165); the target critic lives in ONE flat buffer laid out like the optimizer's,
        #     so the Polyak update is a single launch ---
        self.encoder_target = self.encoder
        if self.no_target_critic:
            self.critic_target, self._critic_target_flat = self.critic, None
        else:
            self.critic_target = deepcopy(self.critic)
            self._critic_target_flat = flatten_like(self.critic_target, self.critic_optim)

        # --- Buffers (shac.py:167-170) ---
        assert self.num_actors == self.num_envs
        self.create_buffers(self.horizon_len, self.num_envs)
        self.reward_shaper = RewardShaper(**cfg_get(cfg, 'reward_shaper', dict(fn='scale', scale=1.0)))

        # --- Entropy (shac.py:172-187) ---
        if self.with_autoent and cfg_get(cfg, 'alpha', None) is None:
            import math
            self.log_alpha = nn.Parameter(torch.tensor(math.log(cfg.init_alpha), device=self.device, dtype=torch.float32))
        self.target_entropy = -self.action_dim * cfg_get(cfg, 'target_entropy_scalar', 1.0)

        # --- Episode metrics (shac.py:189-204) ---
        f = dict(dtype=torch.float32, device=self.device)
        self.episode_rewards = torch.zeros(self.num_envs, **f)
        self.episode_lengths = torch.zeros(self.num_envs, dtype=torch.int64, device=self.device)
        self.episode_discounted_rewards = torch.zeros(self.num_envs, **f)
        self.episode_gamma = torch.ones(self.num_envs, **f)
        self.episode_rewards_tracker, self.episode_lengths_tracker = Tracker(100), Tracker(100)
        self.episode_discounted_rewards_tracker = Tracker(100)
        self.num_episodes = 0

        self.use_cuda_graph = bool(cfg_get(cfg, 'cuda_graph', True))
        self._critic_engine = None
        self._target_fc = None

    @classmethod
    def from_shapes(cls, num_envs, num_obs, num_actions, horizon_len, device, gamma=0.99, lam=0.95, critic_method='td-lambda',
                    critic_iterations=16, num_critic_batches=4, critic_units=(64, 64), num_critics=1, critic_lr=2e-3,
                    critic_betas=(0.7, 0.95), critic_mlp_kwargs=None, optim_type='Adam', max_grad_norm=1.0, truncate_grads=True,
                    cuda_graph=True, normalize_input=False, tanh_clamp=True, logdir=None, **shac_overrides):
        """The same agent built from explicit shapes (benchmarks, kernel tests): a config with the defaults of
        mineral/cfgs/agent/DFlexAntSHAC.yaml:18-50 (critic [64,64] LayerNorm ELU, Adam lr 2e-3 betas (0.7, 0.95), 16 iterations
        x 4 batches, max_grad_norm 1.0) and a shape-only stand-in for the simulator.  num_critics > 1: SAPO's EnsembleCritic."""
        from ...common.attrdict import to_attr
        mlp_kwargs = dict(units=list(critic_units), norm_type='LayerNorm', act_type='ELU') if critic_mlp_kwargs is None else dict(critic_mlp_kwargs)
        if num_critics == 1:
            critic, critic_kwargs = 'Critic', dict(mlp_kwargs=mlp_kwargs)
        else:
            critic, critic_kwargs = 'EnsembleCritic', dict(n_critics=num_critics, mlp_kwargs=mlp_kwargs)
        shac = dict(num_actors=num_envs, max_agent_steps=1e12, horizon_len=horizon_len, gamma=gamma, critic_method=critic_method,
                    critic_iterations=critic_iterations, num_critic_batches=num_critic_batches, optim_type=optim_type,
                    actor_optim_kwargs=dict(lr=critic_lr, betas=tuple(critic_betas)),
                    critic_optim_kwargs=dict(lr=critic_lr, betas=tuple(critic_betas)), max_grad_norm=max_grad_norm,
                    truncate_grads=truncate_grads, cuda_graph=cuda_graph, reward_shaper=dict(fn='scale', scale=1.0))
        shac['lambda'] = lam
        shac.update(shac_overrides)
        cfg = to_attr(dict(rl_device=device, multi_gpu=False, task=dict(env_autoresets=True),
                           agent=dict(network=dict(normalize_input=normalize_input, tanh_clamp=tanh_clamp, actor='Actor',
                                                   actor_kwargs=dict(mlp_kwargs=dict(units=[32], norm_type='LayerNorm', act_type='ELU')),
                                                   critic=critic, critic_kwargs=critic_kwargs),
                                      shac=shac)))
        return cls(cfg, logdir=logdir, env=_ShapeEnv(num_envs, num_obs, num_actions))

    # ---- a13 -----------------------------------------------------------------------------------
    def create_buffers(self, T, B):
        """shac.py:209-224 (done_mask is FLOAT here)."""
        f = dict(dtype=torch.float32, device=self.device)
        self.obs_buf = {k: torch.zeros((T, B) + tuple(v), **f) for k, v in self.obs_space.items()}
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

    # ---- a17: observation normaliser protocol of the rollout -----------------------------------
    @torch.no_grad()
    def rollout_begin(self):
        """shac.py:638-640: freeze a copy of the statistics; the whole rollout normalises with THIS snapshot."""
        self._obs_rms_snapshot = deepcopy(self.obs_rms) if self.obs_rms is not None else None
        return self._obs_rms_snapshot

    @torch.no_grad()
    def rollout_observe(self, obs):
        """shac.py:648-654,694-700: update the LIVE statistics with the raw observation, return it normalised with the
        snapshot of `rollout_begin` (so the statistics of step t never see the data of step t)."""
        if self.obs_rms is None:
            return obs
        for k, v in obs.items():
            self.obs_rms[k].update(v)
        snap = self._obs_rms_snapshot
        return {k: snap[k].normalize(v) for k, v in obs.items()}

    @torch.no_grad()
    def _bootstrap_values(self, obs_norm):
        """critic_target(z, return_type='min_and_avg') (shac.py:707-709) on the critic kernels."""
        if self._target_fc is None:
            self._target_fc = FusedCritics(self.critic_target.mlps(), self.num_envs, self.device)
            self._x_ld = (self.num_obs + 3) // 4 * 4
            self._x_pad = torch.zeros(self.num_envs, self._x_ld, dtype=torch.float32, device=self.device)
        fc, x = self._target_fc, obs_norm['obs']
        rows = x.shape[0]
        self._x_pad[:rows, :self.num_obs].copy_(x)
        preds = fc.forward(self._x_pad[:rows], self._x_ld, rows)
        vs = torch.stack([p.view(-1)[:rows] for p in preds])
        return vs.min(0).values, vs.mean(0)   # (G, N) -> (N,): plumbing-sized (G <= 2)

    @torch.no_grad()
    def collect_rollout(self, act_fn):
        """The critic-side data path of `compute_actor_loss` (shac.py:632-798) for a policy given as a callable:
        `act_fn(step, normalised_obs) -> actions` or `-> dict(actions=, mu=, sigma=, logprob=, distr_ent=)`.  Fills
        obs_buf / action_buf / rew_buf / done_mask / next_values / avg_next_values (/ logprobs / distr_ent) exactly as the
        reference does - live obs_rms update with snapshot normalisation, critic-target bootstrap, zero value on early
        termination or non-finite terminal observation, time-out bootstrap from `info['obs_before_reset']`, last step of
        the horizon marked done - and the episode statistics.  The differentiable return (the actor loss) is not formed."""
        T, N = self.horizon_len, self.num_envs
        self.rollout_begin()
        obs = self._convert_obs(self.env.initialize_trajectory())
        obs = self.rollout_observe(obs)
        for i in range(T):
            for k, v in obs.items():
                self.obs_buf[k][i] = v
            pre_step_obs = obs   # normalised observation the actor saw (the reference's `z`, shac.py:668)
            out = act_fn(i, obs)
            if not isinstance(out, dict):
                out = dict(actions=out)
            actions = out['actions']
            self.action_buf[i] = actions
            if 'mu' in out:
                self.mus[i], self.sigmas[i] = out['mu'], out['sigma']
            obs, rew, done, info = self.env.step(actions)
            obs = self._convert_obs(obs)
            raw_rew = rew.clone()
            rew = self.reward_shaper(rew)
            self.episode_rewards += raw_rew
            self.episode_lengths += 1
            self.episode_discounted_rewards += self.episode_gamma * raw_rew
            self.episode_gamma *= self.gamma
            obs = self.rollout_observe(obs)
            # shac.py:702-709: with the (default) shared encoder the reference bootstraps from `z`, the encoding it computed
            # BEFORE env.step - i.e. from the pre-step observation - and only a separate target encoder sees the new one.
            # Parity copies that quirk (SURVEY.md Appendix B style; verified against the recorded reference rollout).
            nv_min, nv_avg = self._bootstrap_values(pre_step_obs if self.share_encoder else obs)
            done_ids = done.nonzero(as_tuple=False).squeeze(-1)
            if len(done_ids) > 0:
                terminal = self._convert_obs(info['obs_before_reset'])
                bad = torch.zeros(N, dtype=torch.bool, device=self.device)
                for v in terminal.values():
                    flat = v.reshape(N, -1)
                    bad |= (~torch.isfinite(flat)).any(1) | (flat.abs() > 1e6).any(1)
                early = self.episode_lengths < self.max_episode_length
                zero = done & (bad | early)
                boot = done & ~zero
                nv_min, nv_avg = nv_min.clone(), nv_avg.clone()
                nv_min[zero], nv_avg[zero] = 0.0, 0.0
                if boot.any():   # time-out: value of the TRUE last observation, normalised with the snapshot
                    ids = boot.nonzero(as_tuple=False).squeeze(-1)
                    real = {k: v[ids].contiguous() for k, v in terminal.items()}
                    if self.obs_rms is not None:
                        real = {k: self._obs_rms_snapshot[k].normalize(v) for k, v in real.items()}
                    r_min, r_avg = self._bootstrap_values(real)
                    nv_min[ids], nv_avg[ids] = r_min, r_avg
            if (nv_min.abs() > 1e6).any() or (nv_avg.abs() > 1e6).any():
                raise ValueError('next value error')
            self.rew_buf[i] = rew
            if i < T - 1:
                self.done_mask[i] = done.to(torch.float32)
            else:
                self.done_mask[i, :] = 1.0   # shac.py:789-792
            self.next_values[i], self.avg_next_values[i] = nv_min, nv_avg
            if self.with_logprobs and 'logprob' in out:
                self.logprobs[i], self.distr_ent[i] = out['logprob'], out['distr_ent']
            if len(done_ids) > 0:
                ids = done_ids.cpu()
                self.episode_rewards_tracker.update(self.episode_rewards[ids].tolist())
                self.episode_lengths_tracker.update(self.episode_lengths[ids].tolist())
                self.episode_discounted_rewards_tracker.update(self.episode_discounted_rewards[ids].tolist())
                self.num_episodes += len(ids)
                self.episode_rewards[done_ids] = 0.0
                self.episode_lengths[done_ids] = 0
                self.episode_discounted_rewards[done_ids] = 0.0
                self.episode_gamma[done_ids] = 1.0
        self.agent_steps += T * N

    # ---- a14 -----------------------------------------------------------------------------------
    @torch.no_grad()
    def compute_target_values(self):
        """shac.py:877-891: one-step or TD-lambda targets, one launch, bit-exact."""
        ops.critic_targets(self.rew_buf, self.done_mask, self.next_values, self.gamma, getattr(self, 'lam', 0.95),
                           self.critic_method, out=self.target_values)

    def get_alpha(self, detach=True, scalar=False):
        """shac.py:449-463."""
        if not self.with_autoent:
            return self.entropy_coef
        alpha = self.log_alpha.detach().exp()
        return alpha.item() if scalar else alpha

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

    # ---- a15 / a16 -----------------------------------------------------------------------------
    def update_critic(self, dataset=None):
        """shac.py:824-867: critic_iterations x len(dataset) regression steps, all on our kernels (one CUDA graph once
        warm).  Returns the reference's results dict (per-iteration value_loss and mean grad norms)."""
        if dataset is None:
            dataset = CriticDataset(self.critic_batch_size, self.obs_buf, self.target_values)
        if list(dataset.obs) != ['obs']:
            raise NotImplementedError('dict observations are outside the hot path (SURVEY.md section 2 row 18)')
        if self._critic_engine is None:
            self._critic_engine = _CriticEngine(self)
        return self._critic_engine.run(dataset)

    @torch.no_grad()
    def update_target(self):
        """shac.py:429-434: soft_update(critic, critic_target, alpha) = target * alpha + (1 - alpha) * critic, one launch."""
        if self.no_target_critic:
            return
        from ..._native import check, lib, ptr, stream
        alpha = float(self.target_critic_alpha)
        tgt, cur = self._critic_target_flat, self.critic_optim.flat
        check(lib.mnr_soft_update(ptr(tgt, torch.float32), ptr(cur, torch.float32), tgt.numel(), alpha, 1.0 - alpha, stream()),
              'mnr_soft_update')

    def update_actor(self):
        raise NotImplementedError('SHAC actor BPTT needs a differentiable simulator (dflex/warp), which is not part of the '
                                  'reference repository (SURVEY.md section 2 row 6)')

    def train(self):
        raise NotImplementedError('SHAC.train drives update_actor through a differentiable simulator (out of scope, SURVEY.md '
                                  'section 2 row 6); drive collect_rollout / compute_target_values / update_critic / update_target')

    def eval(self):
        raise NotImplementedError

    def set_train(self):
        self.actor.train(), self.critic.train(), self.critic_target.train()

    def set_eval(self):
        self.actor.eval(), self.critic.eval(), self.critic_target.eval()

    # ---- f4: checkpoints with the reference's keys (shac.py:969-1005) -----------------------------
    def save(self, f):
        ckpt = {
            'epoch': self.epoch,
            'mini_epoch': self.mini_epoch,
            'agent_steps': self.agent_steps,
            'obs_rms': self.obs_rms.state_dict() if self.normalize_input else None,
            'actor_encoder': self.actor_encoder.state_dict() if not self.share_encoder else None,
            'encoder': self.encoder.state_dict(),
            'actor': self.actor.state_dict(),
            'critic': self.critic.state_dict(),
            'encoder_target': self.encoder_target.state_dict() if not self.no_target_critic else None,
            'critic_target': self.critic_target.state_dict() if not self.no_target_critic else None,
        }
        torch.save(ckpt, f)

    def load(self, f, ckpt_keys=''):
        all_keys = ('epoch', 'mini_epoch', 'agent_steps', 'obs_rms', 'actor_encoder', 'encoder', 'actor', 'critic',
                    'encoder_target', 'critic_target')
        ckpt = torch.load(f, map_location=self.device, weights_only=False)
        for k in all_keys:
            if not re.match(ckpt_keys, k):
                print(f'Warning: ckpt skipped loading `{k}`')
                continue
            if (k == 'obs_rms' and not self.normalize_input) or (k == 'actor_encoder' and self.share_encoder):
                continue
            if k in ('encoder_target', 'critic_target') and self.no_target_critic:
                continue
            if hasattr(getattr(self, k), 'load_state_dict'):
                getattr(self, k).load_state_dict(ckpt[k])   # parameters are views into the flat buffers: copied in place
            else:
                setattr(self, k, ckpt[k])


class _ObsEncoder(nn.Module):
    """nets.Lambda(lambda x: x['obs']) of the reference (shac.py:97-99): parameter-free, empty state_dict."""

    def forward(self, x):
        return x['obs']


class _CriticEngine:
    """Pre-allocated buffers + the launch sequence of one `update_critic` call (optionally one CUDA graph)."""

    def __init__(self, agent):
        a = self.a = agent
        dev = a.device
        self.fc = None
        self.n_rows = a.num_envs * a.horizon_len
        self.in_dim = a.num_obs
        self.x_ld = (self.in_dim + 3) // 4 * 4
        f32 = dict(dtype=torch.float32, device=dev)
        # first layer input through a 16 B-multiple row stride (TMA); one strided copy per update
        self.x = torch.zeros(self.n_rows, self.x_ld, **f32)
        self.graph, self._warm, self._key, self._ptrs = None, 0, None, None

    def _plan(self, dataset):
        """Batches exactly as the reference iterates them: `len(dataset)` of them (shac.py:832-835), so a dataset built
        with drop_last=True leaves its ragged tail out and the loss is averaged over the same count."""
        a = self.a
        bs, n_batches = dataset.batch_size, len(dataset)
        key = (bs, n_batches)
        if key != self._key:
            f32 = dict(dtype=torch.float32, device=a.device)
            self.bs, self.n_batches = bs, n_batches
            self.fc = FusedCritics(a.critic.mlps(), bs, a.device)
            self.losses = torch.zeros(a.critic_iterations, n_batches, **f32)
            self.norm_before = torch.zeros(a.critic_iterations, n_batches, **f32)
            self.norm_after = torch.zeros(a.critic_iterations, n_batches, **f32)
            self.graph, self._warm, self._key = None, 0, key

    def _enqueue(self, obs, targets):
        a, fc = self.a, self.fc
        max_norm = float(a.max_grad_norm) if (a.truncate_grads and a.max_grad_norm is not None) else 0.0
        if self.x_ld != self.in_dim:
            self.x[:, :self.in_dim].copy_(obs)
            x_all, ld = self.x, self.x_ld
        else:
            x_all, ld = obs, None
        for j in range(a.critic_iterations):
            for i in range(self.n_batches):
                s, e = i * self.bs, min((i + 1) * self.bs, self.n_rows)
                fc.loss_fwd_bwd(x_all[s:e], targets[s:e], self.losses[j, i:i + 1], ld)
                a.critic_optim.step(max_norm, grad_norm_out=self.norm_before[j, i:i + 1], nan_to_num=True,
                                    clipped_norm_out=self.norm_after[j, i:i + 1])

    def run(self, dataset):
        a = self.a
        obs, targets = dataset.obs['obs'], dataset.target_values
        assert obs.shape[0] == self.n_rows and obs.is_contiguous() and targets.is_contiguous()
        self._plan(dataset)
        ptrs = (obs.data_ptr(), targets.data_ptr())
        with torch.no_grad():
            if a.use_cuda_graph:
                if self.graph is not None and ptrs != self._ptrs:   # the dataset moved (e.g. shuffled copy): re-capture
                    self.graph = None
                if self.graph is None and self._warm >= 1:
                    torch.cuda.synchronize()
                    self.graph = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(self.graph):
                        self._enqueue(obs, targets)
                    self._ptrs = ptrs
                if self.graph is not None:
                    self.graph.replay()
                else:
                    self._enqueue(obs, targets)
                    self._warm += 1
            else:
                self._enqueue(obs, targets)
        results = collections.defaultdict(list)
        for j in range(a.critic_iterations):
            results['value_loss'].append(self.losses[j].mean())
            if a.truncate_grads:
                results['grad_norm_before_clip/critic'].append(self.norm_before[j].mean())
                results['grad_norm_after_clip/critic'].append(self.norm_after[j].mean())
        return results
