"""Soft Actor-Critic on the B200 kernels.  Mirrors mineral/agents/sac/sac.py:22-453: same constructor
(`SAC(full_cfg, logdir=, accelerator=, datasets=, env=)`), config keys (`full_cfg.agent.{network,sac}` as in
mineral/cfgs/agent/*SAC*.yaml), methods (`get_actions / explore_env / train / update_net / update_critic / update_actor /
get_alpha / eval / set_train / set_eval / save / load`) and checkpoint keys, so `getattr(mineral_b200.agents, 'SAC')` drops
in for `mineral.agents.SAC` (mineral/scripts/run.py:140-144).

What runs where (reference line -> here); nothing below touches autograd, cuBLAS or a torch elementwise op:
  sac.py:155-161 obs_rms.update + get_actions      -> RunningMeanStd kernels, mnr_rms_apply_ld, actor tower GEMMs (tcgen05),
                                                      mnr_squashed_sample
  sac.py:171-177,181-183 trajectory stores + shaper  -> mnr_traj_store_step (one launch per env step)
  sac.py:183 / 197,205 n-step + ring insert          -> NStepReplay / ReplayBuffer kernels (csrc/rows.cu)
  sac.py:255 sample_batch + :289-301 normalize x3    -> ONE gather that normalises obs / next_obs straight into the padded
                                                      (batch, obs+act) critic-input buffers (the torch.cat of EnsembleQ.forward)
  sac.py:287-297 target                              -> actor + target-critic towers, mnr_squashed_sample, mnr_td_target_stats
  sac.py:299-304 critic regression                   -> critic towers fwd/bwd, mnr_critic_mse_sum_fwd_bwd, mnr_fused_clip_adam
  sac.py:313-326 actor step                          -> actor fwd, critic fwd + INPUT gradient only (frozen critic), mnr_actor_seed,
                                                      mnr_squashed_bwd, actor bwd, mnr_fused_clip_adam
  sac.py:328-331 temperature step                    -> gradient from mnr_actor_seed, mnr_fused_clip_adam on log_alpha
  sac.py:270-273 soft_update                         -> mnr_soft_update over the flat parameter buffers (one launch per network)
The temperature lives on the device (log_alpha); no value is read on the host inside `update_net`.
"""

import collections
import json
import math
import os
import re
from copy import deepcopy

import torch
import torch.nn as nn

from ... import ops
from ...buffers import NStepReplay, ReplayBuffer
from ...common import normalizers
from ...common.reward_shaper import RewardShaper
from ...common.timer import Timer
from ...nets.fused_offpolicy import FusedActor, FusedQ, flatten_like, pad4
from ..agent import Agent, DeviceMetrics, cfg_get
from ..ddpg import models
from ..ddpg.utils import soft_update_flat
from ..ppo.ppo import FusedAdam

_TIMEOUT_KEYS = ('time_outs', 'TimeLimit.truncated', 'truncated')


def make_optim(optim_type, params, kwargs):
    if optim_type not in ('Adam', 'AdamW'):
        raise NotImplementedError(f'optim_type {optim_type}: only Adam / AdamW have a fused kernel')
    return FusedAdam(params, decoupled=(optim_type == 'AdamW'), **dict(kwargs or {}))


class OffPolicyAgent(Agent):
    """What SAC and DDPG share: normalisers, replay plumbing, the trajectory loop of `explore_env`, the padded critic-input
    buffers and the random streams (overridable hooks, so a test can replay the reference's draws)."""

    cfg_key = None

    def _init_common(self, full_cfg, kwargs):
        self.network_config = full_cfg.agent.network
        self.algo_config = cfg = full_cfg.agent[self.cfg_key]
        self.num_actors = cfg.num_actors
        self.max_agent_steps = int(cfg.max_agent_steps)
        Agent.__init__(self, full_cfg, **kwargs)
        if list(self.obs_space) != ['obs'] or len(self.obs_space['obs']) != 1:
            raise NotImplementedError('the fused SAC / DDPG path takes one flat observation key "obs"')
        if cfg_get(self.network_config, 'encoder', None) is not None:
            raise NotImplementedError('encoders (point clouds, images) are outside the rollout->update hot path (SURVEY.md section 8)')
        self.obs_dim = int(self.obs_space['obs'][0])

        # --- Normalizers (sac.py:33-43, ddpg.py:30-40) ---
        rms_config = dict(eps=1e-4, with_clamp=True, initial_count='eps')
        if self.normalize_input:
            rms = {k: normalizers.RunningMeanStd(v, **rms_config) if re.match(self.obs_rms_keys, k) else normalizers.Identity()
                   for k, v in self.obs_space.items()}
            self.obs_rms = nn.ModuleDict(rms).to(self.device)
        else:
            self.obs_rms = None
        self.encoder = lambda obs: obs['obs']   # nets.Lambda(lambda x: x['obs']) (sac.py:51-52)
        self.encoder_target = self.encoder

    def _init_buffers(self):
        cfg = self.algo_config
        self.memory = ReplayBuffer(self.obs_space, self.action_dim, capacity=int(cfg.memory_size), device=self.device)
        self.n_step_buffer = NStepReplay(self.obs_space, self.action_dim, self.num_actors, cfg.nstep, device=self.device)
        self.reward_shaper = RewardShaper(**cfg.reward_shaper)
        self.batch_size = int(cfg.batch_size)
        f32 = dict(dtype=torch.float32, device=self.device)
        O, A = self.obs_dim, self.action_dim
        self.x_ld = pad4(O + A)
        rows = max(self.batch_size, self.num_actors)
        # critic inputs cat(z, a) of the current / next transition; pad columns stay zero
        self.xc = torch.zeros(rows, self.x_ld, **f32)
        self.xn = torch.zeros(rows, self.x_ld, **f32)
        self.reward_b = torch.empty(self.batch_size, 1, **f32)
        self.done_b = torch.empty(self.batch_size, 1, **f32)
        self.target_q = torch.empty(self.batch_size, 1, **f32)
        self._actor_dirty = True   # actor weights changed since the tensor-core (hi, lo) copies were last refreshed
        self.use_cuda_graph = bool(cfg_get(cfg, 'cuda_graph', True))
        self._graphs, self._graph_warm, self._pre, self._pre_buf = {}, {}, None, None

    def _create_metrics(self, tracker_len):
        scores = {'rewards': torch.zeros(self.num_actors, dtype=torch.float32, device=self.device),
                  'lengths': torch.zeros(self.num_actors, dtype=torch.int64, device=self.device)}
        return DeviceMetrics(scores, tracker_len)

    # --- random streams (torch's generator on the device, like the reference) ---
    def _randn(self, shape):
        return torch.randn(shape, dtype=torch.float32, device=self.device)

    def _rand(self, shape):
        return torch.rand(shape, dtype=torch.float32, device=self.device)

    def _sample_indices(self, memory, batch_size):
        return torch.randint(memory.cur_capacity, size=(batch_size,), device=self.device)

    # --- inputs ---
    def _normalize_into(self, obs, out, rows):
        """obs (rows, O) -> out[:rows, :O] (row stride x_ld), normalised + clamped when normalize_input (sac.py:122-123)."""
        rms = self.obs_rms['obs'] if self.normalize_input else None
        if rms is not None and hasattr(rms, 'running_mean'):
            m, d = rms._prepare()
            clamp = tuple(rms.clamp_range) if rms.with_clamp else None
            ops.rms_apply_ld(obs.contiguous(), m, d, clamp, out, self.x_ld)
        else:
            ops.copy_cols(obs.contiguous(), out[:rows, :self.obs_dim])

    def _sample_into_inputs(self, memory, batch_size):
        """memory.sample_batch (sac.py:255) fused with the three `obs_rms.normalize` calls downstream and the cat with the
        action: one gather into xc = [norm(obs) | action] and xn[:, :O] = norm(next_obs)."""
        B, O, A = batch_size, self.obs_dim, self.action_dim
        idx = self._next_indices(memory, B)
        out = dict(obs={'obs': self.xc[:B, :O]}, next_obs={'obs': self.xn[:B, :O]}, action=self.xc[:B, O:O + A],
                   reward=self.reward_b[:B], done=self.done_b[:B])
        rms = self.obs_rms if self.normalize_input else None
        memory.sample_batch(B, device=self.device, indices=idx, obs_rms=rms, out=out)

    @staticmethod
    def _timeouts(infos):
        for k in _TIMEOUT_KEYS:
            if k in infos:
                return infos[k]
        return None

    @torch.no_grad()
    def explore_env(self, env, timesteps: int, random: bool = False, sample: bool = False):
        """sac.py:138-188 / ddpg.py:149-199."""
        N, T, A = self.num_actors, timesteps, self.action_dim
        f32 = dict(dtype=torch.float32, device=self.device)
        traj_obs = {k: torch.empty((N, T) + v, **f32) for k, v in self.obs_space.items()}
        traj_next_obs = {k: torch.empty((N, T) + v, **f32) for k, v in self.obs_space.items()}
        traj_actions = torch.empty((N, T, A), **f32)
        traj_rewards = torch.empty((N, T), **f32)
        traj_dones = torch.empty((N, T), **f32)
        handle_timeout = bool(self.algo_config.handle_timeout)
        for i in range(T):
            if not self.env_autoresets:
                raise NotImplementedError
            if self.normalize_input:
                for k, v in self.obs.items():
                    self.obs_rms[k].update(v)
            if random:
                actions = self._rand((N, A)) * 2.0 - 1.0
            else:
                actions = self.get_actions(obs=self.obs, sample=sample)
            next_obs, rewards, dones, infos = env.step(actions)
            next_obs = self._convert_obs(next_obs)
            rewards, dones = torch.as_tensor(rewards, device=self.device), torch.as_tensor(dones, device=self.device)
            self.metrics.update(self.epoch, self.env, self.obs, rewards, dones, infos)
            tout = self._timeouts(infos) if handle_timeout else None
            ops.traj_store_step(self.obs['obs'].contiguous(), next_obs['obs'].contiguous(), actions.contiguous(), rewards.contiguous().float(),
                                dones.contiguous(), None if tout is None else torch.as_tensor(tout, device=self.device).contiguous(), i,
                                self.reward_shaper.scale, traj_obs['obs'], traj_next_obs['obs'], traj_actions, traj_rewards, traj_dones)
            self.obs = next_obs
        self.metrics.flush_video(self.epoch)
        data = self.n_step_buffer.add_to_buffer(traj_obs, traj_actions, traj_rewards.reshape(N, T, 1), traj_next_obs,
                                                traj_dones.reshape(N, T, 1))
        return data, T * N

    def _train_prologue(self):
        self.obs = self._convert_obs(self.env.reset())
        self.dones = torch.ones((self.num_actors,), dtype=torch.bool, device=self.device)
        self.set_eval()
        trajectory, steps = self.explore_env(self.env, self.algo_config.warm_up, random=True)
        self.memory.add_to_buffer(trajectory)
        self.agent_steps += steps

    def _episode_metrics(self):
        return {'train_scores/episode_rewards': self.metrics.episode_trackers['rewards'].mean(),
                'train_scores/episode_lengths': self.metrics.episode_trackers['lengths'].mean(),
                'train_scores/num_episodes': self.metrics.num_episodes, **self.metrics.result(prefix='train')}

    # --- update_net: eager launches, or ONE CUDA graph per (actor / target update pattern) once warm -----------------------
    def _draw_indices(self, memory, buf):
        """Replay indices of all mini-epochs of one update_net call, drawn up front: buf (mini_epochs, B) int64."""
        buf.random_(0, memory.cur_capacity)

    def _draw_noise(self, buf):
        """Standard-normal noise of all mini-epochs, drawn up front: buf (mini_epochs, draws_per_mini_epoch, B, A)."""
        buf.normal_()

    def _next_indices(self, memory, batch_size):
        pre = self._pre
        return pre['idx'][pre['i']] if pre is not None else self._sample_indices(memory, batch_size)

    def _next_eps(self, rows, slot):
        pre = self._pre
        if pre is not None and rows == self.batch_size:
            return pre['eps'][pre['i'], slot]
        return self._randn((rows, self.action_dim))

    @torch.no_grad()
    def update_net(self, memory):
        """sac.py:251-274 / ddpg.py:239-264.  Returns the reference's results dict: lists of 0-dim device tensors, nothing
        synced.  With `cuda_graph` (default) the replay indices and the noise of all mini-epochs are drawn first (two torch RNG
        launches) and everything else - sampling gather, every network pass, losses, optimizer steps, Polyak updates - replays as
        one CUDA graph of our own kernels.  The result tensors of a replayed call live in the graph's memory pool: read them
        (the agents' `train()` takes their means right away) before the next `update_net` overwrites them."""
        cfg = self.algo_config
        flags = []
        for _ in range(cfg.mini_epochs):
            self.mini_epoch += 1
            flags.append((self.mini_epoch % cfg.update_actor_interval == 0, self.mini_epoch % cfg.update_targets_interval == 0))
        flags = tuple(flags)
        if not self.use_cuda_graph:
            self._pre = None
            return self._enqueue_update(memory, flags)
        n, B, A = cfg.mini_epochs, self.batch_size, self.action_dim
        if self._pre_buf is None:
            self._pre_buf = dict(idx=torch.zeros(n, B, dtype=torch.int64, device=self.device),
                                 eps=torch.zeros(n, self.noise_draws, B, A, dtype=torch.float32, device=self.device), i=0)
        self._draw_indices(memory, self._pre_buf['idx'])
        self._draw_noise(self._pre_buf['eps'])
        self._pre = self._pre_buf
        self._actor_dirty = True   # the captured sequence must not depend on what ran before it
        try:
            entry = self._graphs.get(flags)
            if entry is None and self._graph_warm.get(flags, 0) >= 1:
                torch.cuda.synchronize()
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    results = self._enqueue_update(memory, flags)
                entry = self._graphs[flags] = (graph, results)
            if entry is not None:
                entry[0].replay()
                results = collections.defaultdict(list, {k: list(v) for k, v in entry[1].items()})
            else:
                results = self._enqueue_update(memory, flags)
                self._graph_warm[flags] = self._graph_warm.get(flags, 0) + 1
        finally:
            self._pre = None
            self._actor_dirty = True
        return results

    def _attach_optimizer_splits(self):
        """The optimizer kernels keep the tf32 (hi, lo) copies of the actor's / critics' tensor-core weights current
        (FusedAdam.attach_splits): inside update_net no forward pass of a trained network needs its own split launch."""
        self._splits_attached = False
        fa, fq = self._fa.stack, self._fq.stack
        if fa.use_tc and fq.use_tc and len(fa.split_jobs()) <= 8 and len(fq.split_jobs()) <= 8:
            self.actor_optim.attach_splits(fa.split_jobs())
            self.critic_optim.attach_splits(fq.split_jobs())
            self._splits_attached = True
        self._in_update = False

    def _split_needed(self):
        """Does a forward of the critic need a fresh (hi, lo) split?  Not inside update_net once the optimizers write them."""
        return not (self._splits_attached and self._in_update)

    def _enqueue_update(self, memory, flags):
        results = collections.defaultdict(list)
        if self._splits_attached:   # whatever changed the weights since the last call (load_state_dict, ...): split once
            self._fa.stack._split_weights()
            self._fq.stack._split_weights()
            self._actor_dirty = False
        self._in_update = True
        try:
            for i, (do_actor, do_targets) in enumerate(flags):
                if self._pre is not None:
                    self._pre['i'] = i
                self._sample_into_inputs(memory, self.batch_size)
                self._mini_epoch(results, do_actor, do_targets)
        finally:
            self._in_update = False
        return results

    def _soft_updates(self):
        cfg = self.algo_config
        soft_update_flat(self.critic_target_flat, self.critic_optim.flat, cfg.tau)
        if not cfg.no_tgt_actor:
            soft_update_flat(self.actor_target_flat, self.actor_optim.flat, cfg.tau)

    def optimizer_update(self, optimizer, grad_norm_out):
        """sac.py:334-345: clip_grad_norm_ + step (+ zero_grad), one launch; the total norm lands in grad_norm_out."""
        mgn = self.algo_config.max_grad_norm
        optimizer.step(float(mgn) if mgn is not None else 0.0, grad_norm_out=grad_norm_out)


class SAC(OffPolicyAgent):
    r"""Soft Actor-Critic."""

    cfg_key = 'sac'

    def __init__(self, full_cfg, **kwargs):
        self._init_common(full_cfg, kwargs)
        self.sac_config = cfg = self.algo_config

        # --- Model (sac.py:57-71) ---
        ActorCls = getattr(models, self.network_config.actor)
        CriticCls = getattr(models, self.network_config.critic)
        self.actor = ActorCls(self.obs_dim, self.action_dim, **cfg_get(self.network_config, 'actor_kwargs', {})).to(self.device)
        self.critic = CriticCls(self.obs_dim, self.action_dim, **cfg_get(self.network_config, 'critic_kwargs', {})).to(self.device)
        if self.actor.tanh_policy:
            raise NotImplementedError('SAC needs a stochastic actor (tanh_policy: False)')

        # --- Target networks (sac.py:87-90) before the optimizers re-point the parameters into flat buffers ---
        self.critic_target = deepcopy(self.critic)
        self.actor_target = deepcopy(self.actor) if not cfg.no_tgt_actor else self.actor

        # --- Optim (sac.py:73-85): flat-buffer Adam(W) with device-resident step / lr, clip fused in ---
        self.actor_optim = make_optim(cfg.optim_type, self.actor.parameters(), cfg_get(cfg, 'actor_optim_kwargs', {}))
        self.critic_optim = make_optim(cfg.optim_type, self.critic.parameters(), cfg_get(cfg, 'critic_optim_kwargs', {}))
        self.critic_target_flat = flatten_like(self.critic_target, self.critic_optim)
        self.actor_target_flat = flatten_like(self.actor_target, self.actor_optim) if not cfg.no_tgt_actor else None

        self._init_buffers()

        # --- SAC entropy (sac.py:101-109) ---
        target_entropy_scalar = cfg_get(cfg, 'target_entropy_scalar', 1.0)
        self.target_entropy = -self.action_dim * target_entropy_scalar
        if cfg.alpha is None:
            self.log_alpha = nn.Parameter(torch.tensor(math.log(cfg.init_alpha), device=self.device, dtype=torch.float32))
            self.alpha_optim = make_optim(cfg.optim_type, [self.log_alpha], cfg_get(cfg, 'alpha_optim_kwargs', {}))

        # --- fused engines ---
        B, dev = self.batch_size, self.device
        self._fa = FusedActor(self.actor, max(B, self.num_actors), dev)
        self._fq = FusedQ(self.critic, B, dev)
        self._fq_t = FusedQ(self.critic_target, B, dev)
        dk = cfg_get(cfg_get(self.network_config, 'actor_kwargs', {}), 'dist_kwargs', {}) or {}
        self._squash = ops.make_squash_cfg(dk.get('dist_type', 'normal'), dk.get('minlogstd'), dk.get('maxlogstd'),
                                           self.actor.dist.minstd, self.actor.dist.maxstd)
        f32 = dict(dtype=torch.float32, device=dev)
        # --- Timing (sac.py:111-117): same scopes, device time from CUDA events (common/timer.py) ---
        self.timer = Timer()
        self.timer.wrap('agent', self, ['explore_env', 'update_net'])
        self.timer.wrap('memory', self.memory, ['add_to_buffer'])
        if hasattr(self.env, 'step'):
            self.timer.wrap('env', self.env, ['step'])
        self.timer_total_names = ('agent.explore_env', 'memory.add_to_buffer', 'agent.update_net')
        self._attach_optimizer_splits()
        self.log_prob = torch.empty(max(B, self.num_actors), **f32)
        self._eps_actor = None
        self._scalars = torch.zeros(16, **f32)   # scratch for per-step scalar outputs that nobody keeps

    # ------------------------------------------------------------------------------------------------
    def _alpha_args(self):
        """(device log_alpha | None, host alpha) for the kernels (get_alpha, sac.py:275-285)."""
        if self.sac_config.alpha is None:
            return self.log_alpha.data.view(1), 0.0
        return None, float(self.sac_config.alpha)

    def get_alpha(self, detach=True, scalar=False):
        if self.sac_config.alpha is None:
            alpha = self.log_alpha.exp()
            if detach:
                alpha = alpha.detach()
            if scalar:
                alpha = alpha.item()
            return alpha
        return self.sac_config.alpha

    def _policy(self, x, rows, sample, log_prob, actions_out, slot=0):
        """actor(z) -> Dist -> rsample / mean [-> log_prob] on rows of the padded input buffer x (z in columns [0, O))."""
        fa = self._fa
        fa.forward(x, self.x_ld, rows, split=self._actor_dirty)
        self._actor_dirty = False
        eps = self._next_eps(rows, slot) if sample else None
        ops.squashed_sample(fa.mu_raw, fa.logstd, eps, self._squash, actions_out, self.log_prob if log_prob else None, rows)
        return eps

    @torch.no_grad()
    def get_actions(self, obs=None, z=None, sample=True, logprob=False):
        """sac.py:119-136 for rollouts / evaluation: obs dict -> actions (N, A) [, None, log_prob (N, 1)]."""
        if z is not None:
            raise NotImplementedError('pre-encoded inputs: the update path feeds the fused engines directly')
        x = obs['obs']
        rows = x.shape[0]
        assert rows <= self.xn.shape[0]
        self._normalize_into(x, self.xn, rows)
        actions = torch.empty(rows, self.action_dim, dtype=torch.float32, device=self.device)
        self._policy(self.xn, rows, sample, logprob, actions)
        if logprob:
            return actions, None, self.log_prob[:rows].clone().view(rows, 1)
        return actions

    # ------------------------------------------------------------------------------------------------
    def train(self):
        """sac.py:190-249."""
        self._train_prologue()
        cfg = self.sac_config
        while self.agent_steps < self.max_agent_steps:
            self.epoch += 1
            self.set_eval()
            trajectory, steps = self.explore_env(self.env, cfg.horizon_len, sample=True)
            self.agent_steps += steps
            self.memory.add_to_buffer(trajectory)
            self.set_train()
            results = self.update_net(self.memory)
            metrics = {k: torch.mean(torch.stack(v)).item() for k, v in results.items()}
            metrics.update({'epoch': self.epoch, 'mini_epoch': self.mini_epoch, 'alpha': self.get_alpha(scalar=True)})
            metrics = {f'train_stats/{k}': v for k, v in metrics.items()}
            timings = self.timer.stats(step=self.agent_steps, total_names=self.timer_total_names, reset=False)
            metrics.update({f'train_timings/{k}': v for k, v in timings.items()})
            metrics.update(self._episode_metrics())
            self.writer.add(self.agent_steps, metrics)
            self.writer.write()
            self._checkpoint_save(metrics['train_scores/episode_rewards'])
            if self.print_every > 0 and (self.epoch + 1) % self.print_every == 0:
                print(f'Epochs: {self.epoch + 1} | Agent Steps: {int(self.agent_steps):,} | '
                      f'ep_rewards {metrics["train_scores/episode_rewards"]:.2f}, ep_lengths {metrics["train_scores/episode_lengths"]:.2f}')
        self.save(os.path.join(self.ckpt_dir, 'final.pth'))

    noise_draws = 2   # per mini-epoch: a' ~ pi(.|s') for the target, a ~ pi(.|s) for the actor step

    def _mini_epoch(self, results, do_actor, do_targets):
        """One iteration of the loop in SAC.update_net (sac.py:253-273) on the batch already gathered into xc / xn."""
        critic_loss, critic_grad_norm, target_values = self.update_critic()
        results['loss/critic'].append(critic_loss)
        results['grad_norm/critic'].append(critic_grad_norm)
        for k, v in target_values.items():
            results[k].append(v)
        if do_actor:
            actor_loss, alpha_loss, entropy, actor_grad_norm = self.update_actor()
            results['loss/actor'].append(actor_loss)
            results['loss/alpha'].append(alpha_loss)
            results['entropy'].append(entropy)
            results['grad_norm/actor'].append(actor_grad_norm)
        if do_targets:
            self._soft_updates()

    @torch.no_grad()
    def update_critic(self, obs=None, action=None, reward=None, next_obs=None, done=None):
        """sac.py:287-311 on the batch `_sample_into_inputs` left in xc / xn / reward_b / done_b (explicit tensors are copied in)."""
        cfg, B, O, A = self.sac_config, self.batch_size, self.obs_dim, self.action_dim
        if obs is not None:
            self._load_batch(obs, action, reward, next_obs, done)
        # target: a' ~ pi(.|s'), min_k Q_target_k(s', a') - alpha log pi(a'|s')
        self._policy(self.xn, B, True, True, self.xn[:B, O:O + A])
        q_t = self._fq_t.forward(self.xn, B)
        log_alpha, alpha_host = self._alpha_args()
        out = torch.empty(6, dtype=torch.float32, device=self.device)   # target stats [0:4], critic loss [4], grad norm [5]
        ops.td_target_stats([q.view(-1)[:B] for q in q_t], self.log_prob[:B] if cfg.backup_entropy else None, self.reward_b.view(-1),
                            self.done_b.view(-1), log_alpha, alpha_host, cfg.gamma ** cfg.nstep, self.target_q.view(-1), out[0:4])
        # regression of both critics on the shared target + clip + optimizer step
        self._fq.loss_fwd_bwd(self.xc, self.target_q, out[4:5], split=self._split_needed())
        self.optimizer_update(self.critic_optim, out[5:6])
        target_values = {'target_values/mean': out[0], 'target_values/std': out[1], 'target_values/max': out[2], 'target_values/min': out[3]}
        return out[4], out[5], target_values

    @torch.no_grad()
    def update_actor(self, obs=None):
        """sac.py:313-332 on the observations in xc[:, :O] (the sampled actions in xc are overwritten by fresh policy samples)."""
        cfg, B, O, A = self.sac_config, self.batch_size, self.obs_dim, self.action_dim
        fa, fq = self._fa, self._fq
        if obs is not None:
            self._normalize_into(obs['obs'] if isinstance(obs, dict) else obs, self.xc, B)
        eps = self._policy(self.xc, B, True, True, self.xc[:B, O:O + A], slot=1)
        fq.forward(self.xc, B, split=self._split_needed())   # the critic step just changed the weights
        log_alpha, alpha_host = self._alpha_args()
        out = torch.empty(5, dtype=torch.float32, device=self.device)   # actor loss, entropy, alpha loss, alpha, grad norm
        learn_alpha = cfg.alpha is None
        ops.actor_seed([q.view(-1)[:B] for q in fq.pred], [g.view(-1)[:B] for g in fq.g_pred], self.log_prob[:B], log_alpha, alpha_host,
                       self.target_entropy, out[0:4], self.log_alpha.grad.view(1) if learn_alpha else None)
        dx = fq.input_grad(self.xc, B)               # d(-mean Q_min)/d(z, a) per tower; only the action block is used (z is data)
        ops.squashed_bwd(fa.mu_raw, fa.logstd, eps, dx[0][:B, O:O + A], log_alpha, alpha_host, self._squash, fa.g_mu, fa.g_ls, B,
                         g_actions2=dx[1][:B, O:O + A] if len(dx) > 1 else None)
        fa.backward(self.xc, B)
        self.optimizer_update(self.actor_optim, out[4:5])
        self._actor_dirty = self._split_needed()
        if learn_alpha:
            self.optimizer_update(self.alpha_optim, self._scalars[0:1])
        return out[0], out[2], out[1], out[4]

    def _load_batch(self, obs, action, reward, next_obs, done):
        B, O, A = self.batch_size, self.obs_dim, self.action_dim
        self._normalize_into(obs['obs'], self.xc, B)
        self._normalize_into(next_obs['obs'], self.xn, B)
        ops.copy_cols(action.contiguous(), self.xc[:B, O:O + A])
        self.reward_b.copy_(reward.view(B, 1))
        self.done_b.copy_(done.view(B, 1))

    # ------------------------------------------------------------------------------------------------
    def eval(self):
        """sac.py:347-398."""
        self.set_eval()
        obs = self._convert_obs(self.env.reset())
        total_eval_episodes = self.num_actors * 2
        eval_metrics = self._create_metrics(total_eval_episodes)
        old, self.metrics = self.metrics, eval_metrics
        try:
            with torch.no_grad():
                while self.metrics.num_episodes < total_eval_episodes:
                    if not self.env_autoresets:
                        raise NotImplementedError
                    actions = self.get_actions(obs=obs, sample=True)
                    next_obs, rewards, dones, infos = self.env.step(actions)
                    next_obs = self._convert_obs(next_obs)
                    rewards, dones = torch.as_tensor(rewards, device=self.device), torch.as_tensor(dones, device=self.device)
                    self.metrics.update(self.epoch, self.env, obs, rewards, dones, infos)
                    obs = next_obs
            metrics = {'eval_scores/num_episodes': self.metrics.num_episodes,
                       'eval_scores/episode_rewards': self.metrics.episode_trackers['rewards'].mean(),
                       'eval_scores/episode_lengths': self.metrics.episode_trackers['lengths'].mean()}
            self.writer.add(self.agent_steps, metrics)
            self.writer.write()
            scores = dict(metrics, epoch=self.epoch, mini_epoch=self.mini_epoch, agent_steps=self.agent_steps)
            json.dump(scores, open(os.path.join(self.logdir, 'scores.json'), 'w'), indent=4)
        finally:
            self.metrics = old
        return metrics

    def set_train(self):
        for m in (self.actor, self.critic, self.actor_target, self.critic_target):
            m.train()

    def set_eval(self):
        for m in (self.actor, self.critic, self.actor_target, self.critic_target):
            m.eval()

    def save(self, f):
        """sac.py:414-428: same keys (the encoder is the parameter-free `x['obs']` lambda: empty state dicts)."""
        cfg = self.sac_config
        ckpt = {
            'epoch': self.epoch, 'mini_epoch': self.mini_epoch, 'agent_steps': self.agent_steps,
            'obs_rms': self.obs_rms.state_dict() if self.normalize_input else None,
            'encoder': {}, 'actor': self.actor.state_dict(), 'critic': self.critic.state_dict(), 'encoder_target': {},
            'actor_target': self.actor_target.state_dict() if not cfg.no_tgt_actor else None,
            'critic_target': self.critic_target.state_dict(),
            'log_alpha': self.log_alpha.data.clone() if cfg.alpha is None else None,
        }
        torch.save(ckpt, f)

    def load(self, f, ckpt_keys=''):
        """sac.py:430-453.  Parameters are copied INTO the flat optimizer buffers (load_state_dict copies in place)."""
        cfg = self.sac_config
        keys = ('epoch', 'mini_epoch', 'agent_steps', 'obs_rms', 'encoder', 'actor', 'critic', 'encoder_target', 'actor_target',
                'critic_target', 'log_alpha')
        ckpt = torch.load(f, map_location=self.device)
        for k in keys:
            if not re.match(ckpt_keys, k):
                print(f'Warning: ckpt skipped loading `{k}`')
                continue
            if (k == 'obs_rms' and not self.normalize_input) or (k == 'actor_target' and cfg.no_tgt_actor) or \
                    (k == 'log_alpha' and cfg.alpha is not None) or k in ('encoder', 'encoder_target'):
                continue
            if k == 'log_alpha':
                self.log_alpha.data.copy_(ckpt[k])
            elif hasattr(getattr(self, k), 'load_state_dict'):
                getattr(self, k).load_state_dict(ckpt[k])
            else:
                setattr(self, k, ckpt[k])
        self._actor_dirty = True
