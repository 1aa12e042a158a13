"""PPO on the B200 kernels.  Mirrors mineral/agents/ppo/ppo.py:19-498: same constructor
(`PPO(full_cfg, logdir=, accelerator=, datasets=, env=)`), same config keys
(`full_cfg.agent.{network,ppo}` as in mineral/cfgs/agent/*.yaml), same methods
(`train / eval / play_steps / train_epoch / model_act / set_train / set_eval / save / load`) and the
same module-level loss functions, so `getattr(mineral_b200.agents, cfg.agent.algo)` drops in for
`mineral.agents.PPO` (mineral/scripts/run.py:140-144).

What runs where during an update (reference line -> here); every compute launch is one of this repo's kernels:
  ppo.py:172-183 compute_return + prepare_training + value RMS -> ExperienceBuffer.compute_return_and_prepare
                                                                  (mnr_post_rollout_fused + _finalize)
  ppo.py:258-262 obs_rms.update + normalize per minibatch       -> mnr_rms_batch_moments_f32 once per rollout,
                                                                  mnr_rms_schedule (all steps' statistics up
                                                                  front), mnr_rms_apply_ld_f32 per step
  ppo.py:270     model(batch)                                    -> mnr_tc_linear_fwd per tower layer (tcgen05 3xTF32,
                                                                  both towers in one launch; mnr_linear_fwd for shapes
                                                                  the tensor-core kernel cannot take; mnr_ln_act_fwd
                                                                  behind a LayerNorm)
  ppo.py:271-285,302-303,313 heads/dist/entropy/neglogp/losses/KL/write-back + the head gradients and d loss / d Y
                                                                 -> mnr_ppo_heads_loss_fwd_bwd (one kernel)
  ppo.py:291-292 zero_grad / backward                            -> mnr_tc_linear_bwd_weight_partials + mnr_tc_linear_bwd_input
                                                                  per layer, mnr_tc_reduce_partials once (no autograd graph)
  ppo.py:294-300 clip_grad_norm_ + optimizer.step                -> mnr_fused_clip_adam / mnr_peer_reduce_clip_adam (one launch)
  ppo.py:335-348 avg_kl.item() + LR schedule                     -> mnr_kl_lr_update ('kl') / mnr_lr_table_update ('cos', 'linear')
No value is read on the host inside `train_epoch` until the results are assembled at its end.  There is no torch-module /
autograd / NCCL data path: a network `FusedTowers.supported()` rejects is an error at construction.
"""

import collections
import os
import re

import torch
import torch.nn as nn

from ... import ops
from ..._native import PPO_METRICS
from ...common import normalizers
from ...common.reward_shaper import RewardShaper
from ..agent import Agent, cfg_get
from . import models
from .experience import ExperienceBuffer
from .utils import AdaptiveScheduler, LinearScheduler, adjust_learning_rate_cos


def peer_tail_floats():
    from ...peer import TAIL_FLOATS
    return TAIL_FLOATS


class FusedAdam:
    """Adam / AdamW over ONE flat f32 buffer with device-resident lr / step / clip coefficient.
    Exposes `param_groups`, `state_dict()`, `load_state_dict()` in torch.optim.Adam's format so PPO
    checkpoints (ppo.py:416-446) stay interchangeable."""

    def __init__(self, params, decoupled, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=None, comm=None, **other):
        """comm: optional `mineral_b200.peer.PeerComm` whose 'grad' region ((n + 32) floats) holds the flat gradient, so
        the peers can read it over NVLink (multi-GPU); the step is then ONE fused all-reduce + clip + Adam kernel.
        Any torch.optim keyword this kernel does not implement (amsgrad, maximize, ...) is an error unless it is switched
        off: a config must not silently train differently from the reference."""
        unsupported = {k: v for k, v in other.items() if v not in (False, None) and k not in ('foreach', 'fused', 'capturable')}
        if unsupported:
            raise NotImplementedError(f'optim_kwargs not implemented by the fused Adam/AdamW kernel: {sorted(unsupported)}')
        self.params = list(params)
        self.comm = comm
        dev = self.params[0].device
        self.decoupled = decoupled
        self.betas, self.eps = betas, eps
        self.weight_decay = (1e-2 if decoupled else 0.0) if weight_decay is None else weight_decay
        sizes = [p.numel() for p in self.params]
        self.offsets = [0]
        for s in sizes:
            self.offsets.append(self.offsets[-1] + ((s + 3) // 4) * 4)  # keep every tensor 16 B aligned
        n = self.offsets[-1]
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        if comm is None:
            self.flat_grad = torch.zeros(n, dtype=torch.float32, device=dev)
            self.tail = self.reduced = None
        else:
            # PING-PONG gradient regions ('grad', 'grad1'): consecutive optimizer steps write their gradients in turn, so a
            # region is only rewritten two steps after the peers read it and the fused kernel needs ONE cross-GPU barrier per
            # step instead of two (use_grad_buffer / step(end_barrier=...); csrc/peer.cu)
            self._regions = [r for r in ('grad', 'grad1') if r in comm.offsets]
            self._grad_bufs, self._tails = [], []
            for r in self._regions:
                region = comm.tensor(r)
                assert region.numel() >= n + peer_tail_floats(), 'peer gradient region too small'
                self._grad_bufs.append(region[:n])
                self._tails.append(region[n:n + peer_tail_floats()])   # this rank's loss metrics, summed by the fused kernel
            self._cur = 0
            self.flat_grad, self.tail = self._grad_bufs[0], self._tails[0]
            self.reduced = torch.zeros(n, dtype=torch.float32, device=dev)
            self.trace = torch.zeros(512, dtype=torch.int64, device=dev) if os.environ.get('MNR_PEER_TRACE') else None
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=dev)
        for p, o in zip(self.params, self.offsets):
            view = self.flat[o:o + p.numel()].view_as(p)
            view.copy_(p.data)
            p.data = view
            p.grad = self.flat_grad[o:o + p.numel()].view_as(p)
        if comm is not None and len(self._grad_bufs) > 1:   # the `.grad` views of both regions, built once
            self._grad_views = [[buf[o:o + p.numel()].view_as(p) for p, o in zip(self.params, self.offsets)] for buf in self._grad_bufs]
        self.state = ops.optim_state_tensor(lr, dev)
        self.ws = torch.zeros(ops.lib.mnr_grad_norm_workspace_bytes(n), dtype=torch.uint8, device=dev)
        self.fused_ws = torch.zeros(ops.lib.mnr_fused_adam_workspace_bytes(), dtype=torch.uint8, device=dev)
        self.param_groups = [dict(lr=float(lr), betas=betas, eps=eps, weight_decay=self.weight_decay)]

    def use_grad_buffer(self, i):
        """Multi-GPU: point every parameter's `.grad` (and `flat_grad` / `tail`) at ping-pong region i & 1 for the NEXT backward
        pass.  Host-side pointer bookkeeping only (done while the launch sequence is enqueued / captured)."""
        if self.comm is None or len(self._grad_bufs) < 2:
            return
        self._prev, self._cur = self._cur, i & 1
        self.flat_grad, self.tail = self._grad_bufs[self._cur], self._tails[self._cur]
        for p, g in zip(self.params, self._grad_views[self._cur]):
            p.grad = g

    def step(self, max_grad_norm, kl=None, grad_norm_out=None, metrics_out=None, kl_index=-1, fused=True, nan_to_num=False,
             clipped_norm_out=None, end_barrier=True):
        """clip_grad_norm_ + optimizer.step + zero_grad with no host sync: ONE launch (csrc/peer.cu), which with a peer
        communicator also all-reduces the gradient and the metrics tail over NVLink (`metrics_out` receives the sums,
        the metric at `kl_index` feeds the KL accumulator).  fused=False: the two-launch form (csrc/optim.cu)."""
        if self.comm is not None:
            pp = len(self._grad_bufs) > 1
            self.comm.reduce_clip_adam(self.flat.numel(), self.reduced, self.flat, self.exp_avg, self.exp_avg_sq, max_grad_norm,
                                       self.betas[0], self.betas[1], self.eps, self.weight_decay, self.decoupled, True, self.state,
                                       metrics_out, kl_index, grad_norm_out, self.fused_ws, region=self._regions[self._cur],
                                       zero_region=self._regions[1 - self._cur] if pp else None,
                                       end_barrier=end_barrier or not pp, trace=self.trace)
        elif fused:
            ops.fused_clip_adam(self.flat, self.flat_grad, self.exp_avg, self.exp_avg_sq, max_grad_norm, self.betas[0], self.betas[1],
                                self.eps, self.weight_decay, self.decoupled, True, self.state, kl, grad_norm_out, self.fused_ws,
                                nan_to_num=nan_to_num, clipped_norm_out=clipped_norm_out)
        else:
            ops.grad_norm_prepare(self.flat_grad, max_grad_norm, self.betas[0], self.betas[1], self.state, kl, grad_norm_out, self.ws)
            ops.adam_step(self.flat, self.flat_grad, self.exp_avg, self.exp_avg_sq, self.betas[0], self.betas[1], self.eps,
                          self.weight_decay, self.decoupled, True, self.state)

    def attach_splits(self, jobs):
        """jobs: [(weight Parameter living in this optimizer's flat buffer, hi, lo, ld)] for up to 8 tensor-core layers: the
        fused step kernel then also writes the tf32 (hi, lo) copies of the updated weights (the table of
        mnr_fused_adam_split_table_offset in include/mineral_b200.h), so no separate split launch is needed per step."""
        import struct
        assert len(jobs) <= 8
        base, blob = self.flat.data_ptr(), struct.pack('<4i', len(jobs), 0, 0, 0)
        for w, hi, lo, ld in jobs:
            off = (w.data.data_ptr() - base) // 4
            n, k = w.shape
            assert 0 <= off and off + n * k <= self.flat.numel() and off % 4 == 0 and hi.is_contiguous() and lo.is_contiguous()
            blob += struct.pack('<q4i2Q', off, n * k, k, ld, 0, hi.data_ptr(), lo.data_ptr())
        o = ops.lib.mnr_fused_adam_split_table_offset()
        self.fused_ws[o:o + len(blob)].copy_(torch.frombuffer(bytearray(blob), dtype=torch.uint8))
        self._split_refs = jobs   # keep the (hi, lo) buffers alive
        self.splits_attached = len(jobs) > 0

    def zero_grad(self):
        self.flat_grad.zero_()

    def read_state(self):
        return ops.read_optim_state(self.state)

    def set_lr(self, lr):
        st = self.read_state()
        st.lr, st.lr_f32 = float(lr), float(lr)
        self.state.copy_(torch.frombuffer(bytearray(bytes(st)), dtype=torch.uint8))
        self.param_groups[0]['lr'] = float(lr)

    def state_dict(self):
        st = self.read_state()
        state = {}
        for i, (p, o) in enumerate(zip(self.params, self.offsets)):
            n = p.numel()
            state[i] = dict(step=torch.tensor(float(st.step)), exp_avg=self.exp_avg[o:o + n].view_as(p).clone(),
                            exp_avg_sq=self.exp_avg_sq[o:o + n].view_as(p).clone())
        group = dict(self.param_groups[0], lr=st.lr, params=list(range(len(self.params))))
        return dict(state=state, param_groups=[group])

    def load_state_dict(self, sd):
        step = 0
        for i, (p, o) in enumerate(zip(self.params, self.offsets)):
            s = sd['state'].get(i)
            if s is None:
                continue
            n = p.numel()
            self.exp_avg[o:o + n].copy_(s['exp_avg'].reshape(-1))
            self.exp_avg_sq[o:o + n].copy_(s['exp_avg_sq'].reshape(-1))
            step = int(s['step'])
        st = self.read_state()
        st.step = step
        st.lr = st.lr_f32 = float(sd['param_groups'][0]['lr'])
        self.state.copy_(torch.frombuffer(bytearray(bytes(st)), dtype=torch.uint8))
        self.param_groups[0]['lr'] = st.lr


class PPO(Agent):
    r"""Proximal Policy Optimization."""

    @property
    def last_lr(self):
        """ppo.py:338-345 keeps the last learning rate as a Python float; here it is device-resident and fetched lazily."""
        if getattr(self, '_lr_stale', False):
            st = self.optim.read_state()   # device -> host read (synchronises)
            self._last_lr = st.lr
            self.optim.param_groups[0]['lr'] = st.lr
            self._lr_stale = False
        return self._last_lr

    @last_lr.setter
    def last_lr(self, value):
        self._last_lr, self._lr_stale = value, False

    def __init__(self, full_cfg, **kwargs):
        self.network_config = full_cfg.agent.network
        self.ppo_config = full_cfg.agent.ppo
        self.num_actors = self.ppo_config.num_actors
        self.max_agent_steps = int(self.ppo_config.max_agent_steps)
        super().__init__(full_cfg, **kwargs)
        cfg = self.ppo_config
        if cfg_get(full_cfg.agent, 'dapg', None) is not None or cfg_get(cfg, 'dapg', None) is not None:
            raise NotImplementedError('DAPG needs demonstration datasets and is out of scope (SURVEY.md section 2 row 12)')
        self.dapg_config = None

        # --- Normalizers (ppo.py:29-43) ---
        rms_config = dict(eps=1e-5, with_clamp=True, initial_count=1, dtype=torch.float64)
        if self.normalize_input:
            rms = {}
            for k, v in self.obs_space.items():
                rms[k] = normalizers.RunningMeanStd(v, **rms_config) if re.match(self.obs_rms_keys, k) else normalizers.Identity()
            self.obs_rms = nn.ModuleDict(rms).to(self.device)
        else:
            self.obs_rms = None
        self.value_rms = normalizers.RunningMeanStd((1,), **rms_config).to(self.device)

        # --- Model (ppo.py:45-52) ---
        ModelCls = getattr(models, cfg_get(self.network_config, 'actor_critic', 'ActorCritic'))
        model_kwargs = cfg_get(self.network_config, 'actor_critic_kwargs', {})
        self.model = ModelCls(self.obs_space, self.action_dim, encoder=cfg_get(self.network_config, 'encoder', None),
                              encoder_kwargs=cfg_get(self.network_config, 'encoder_kwargs', None), **model_kwargs)
        self.model.to(self.device)

        # --- Optim (ppo.py:54-61): Adam / AdamW re-implemented over a flat buffer ---
        optim_kwargs = dict(cfg_get(cfg, 'optim_kwargs', {}))
        self.init_lr = float(optim_kwargs.get('lr', 3e-4))
        self.last_lr = self.init_lr
        optim_kwargs['lr'] = self.init_lr
        if cfg.optim_type not in ('Adam', 'AdamW'):
            raise NotImplementedError(f'optim_type {cfg.optim_type}: only Adam / AdamW have a fused kernel')

        # --- Networks the kernels cannot run are an error HERE, not a silent torch-module / autograd path ---
        from ...nets.fused_towers import FusedTowers
        if not FusedTowers.supported(self.model) or list(self.obs_space) != ['obs'] or len(self.obs_space['obs']) != 1:
            raise NotImplementedError(
                'mineral_b200 PPO runs separate actor / value MLP towers of equal shape (Linear [-> LayerNorm] -> ELU | ReLU | SiLU, '
                'no dropout) on a flat `obs` vector; shared trunks, encoders and dict observations are outside the hot path '
                '(SURVEY.md section 2 rows 5, 18)')
        if self.action_dim > 32:
            raise NotImplementedError('action_dim > 32: the policy-step / heads kernels hold one action vector per thread')

        # --- Multi-GPU: one process per GPU; the data path runs over NVLink peer memory with our own kernels
        #     (mineral_b200/peer.py, csrc/peer.cu); torch.distributed (NCCL) only carries the set-up handshake ---
        self.world, self.comm = 1, None
        if self.multi_gpu:
            import torch.distributed as dist
            assert dist.is_initialized(), 'multi_gpu=True needs torch.distributed initialised (torchrun)'
            self.world = dist.get_world_size()
            if self.world > 1:
                from ...peer import PeerComm, TAIL_FLOATS
                n_flat = sum((p.numel() + 3) // 4 * 4 for p in self.model.parameters())
                self.comm = PeerComm({'grad': (n_flat + TAIL_FLOATS) * 4, 'grad1': (n_flat + TAIL_FLOATS) * 4, 'stage': 1 << 20}, self.device)
        self.optim = FusedAdam(self.model.parameters(), decoupled=(cfg.optim_type == 'AdamW'), comm=self.comm, **optim_kwargs)

        # --- PPO train params (ppo.py:63-76) ---
        self.e_clip = cfg['e_clip']
        self.use_smooth_clamp = cfg['use_smooth_clamp']
        self.clip_value_loss = cfg['clip_value_loss']
        self.entropy_coef = cfg['entropy_coef']
        self.critic_coef = cfg['critic_coef']
        self.bounds_loss_coef = cfg['bounds_loss_coef']
        self.bounds_type = cfg['bounds_type']
        self.gamma = cfg['gamma']
        self.tau = cfg['tau']
        self.truncate_grads = cfg['truncate_grads']
        self.max_grad_norm = cfg['max_grad_norm']
        self.value_bootstrap = cfg['value_bootstrap']
        self.normalize_advantage = cfg['normalize_advantage']  # read but, as in the reference, never consulted
        self.normalize_value = cfg['normalize_value']

        # --- Collect params (ppo.py:78-83) ---
        self.horizon_len = cfg['horizon_len']
        self.batch_size = self.horizon_len * self.num_actors
        self.minibatch_size = cfg['minibatch_size']
        self.mini_epochs = cfg['mini_epochs']
        assert self.batch_size % self.minibatch_size == 0 or 'train' not in cfg_get(full_cfg, 'run', 'train')

        # --- LR schedule (ppo.py:85-93): 'fixed' | 'linear' | 'kl' | 'cos' (mineral/cfgs/agent/DFlexAntPPO.yaml:63) ---
        self.lr_schedule = cfg['lr_schedule']
        if self.lr_schedule not in ('fixed', 'constant', 'linear', 'kl', 'cos'):
            raise NotImplementedError(f'lr_schedule {self.lr_schedule!r}')
        self.min_lr, self.max_lr = cfg_get(cfg, 'min_lr', 1e-6), cfg_get(cfg, 'max_lr', 1e-2)
        if self.lr_schedule == 'kl':
            self.kl_threshold = cfg['kl_threshold']
            self.scheduler = AdaptiveScheduler(self.kl_threshold, self.min_lr, self.max_lr)
        elif self.lr_schedule == 'linear':
            self.scheduler = LinearScheduler(self.init_lr, self.min_lr, cfg['max_agent_steps'])

        # --- Storage (ppo.py:95-107) ---
        self.obs, self.dones = None, None
        self.storage = ExperienceBuffer(self.num_actors, self.horizon_len, self.batch_size, self.minibatch_size, self.obs_space,
                                        self.action_dim, self.device, self.cpu_obs_keys)
        self.reward_shaper = RewardShaper(**cfg['reward_shaper'])

        self._engine = None
        self._rollout = None
        self.use_cuda_graph = bool(cfg_get(cfg, 'cuda_graph', False))

    # -------------------------------------------------------------------------------------------
    @torch.no_grad()
    def model_act(self, obs_dict, sample=True, **kwargs):
        """ppo.py:125-131 on the rollout kernels: obs normalise -> tower GEMMs -> policy-step kernel (heads, sample,
        neglogp, un-normalised value).  Returns the reference's keys `actions / neglogp / values / mu / sigma`."""
        ro = self._rollout_engine()
        ro.begin_rollout()
        N, A = self.num_actors, self.action_dim
        assert obs_dict['obs'].shape[0] == N, 'model_act expects one observation per actor'
        f32 = dict(dtype=torch.float32, device=self.device)
        out = dict(actions=torch.empty(N, A, **f32), neglogp=torch.empty(N, **f32), values=torch.empty(N, 1, **f32),
                   mu=torch.empty(N, A, **f32), sigma=torch.empty(N, A, **f32))
        ro._forward(obs_dict)
        if not sample:   # models.py:113-114: the raw `mu` head output, nothing else
            return ro.towers.heads_forward()[0].clone()
        ro.eps.normal_()
        ro._policy(ro.eps, out['actions'], out['neglogp'], out['values'].view(-1), out['mu'], out['sigma'], None,
                   unnormalize_value=self.normalize_value)
        return out

    def _rollout_engine(self):
        if self._rollout is None:
            if not _RolloutEngine.supported(self):
                raise NotImplementedError('the rollout kernels need a heads input of <= 512 features and a `scale` reward shaper')
            self._rollout = _RolloutEngine(self)
        return self._rollout

    @torch.no_grad()
    def play_steps(self):
        """ppo.py:133-183.  The agent side of every env step runs on our kernels (`_RolloutEngine`: obs normalise -> tower
        GEMMs -> one policy-step kernel writing into the storage rows -> env.step -> one post-step kernel; SURVEY.md
        section 8f row 1); the tail is the fused post-rollout pass."""
        self._rollout_engine()
        return self._play_steps_fused()

    @torch.no_grad()
    def _play_steps_fused(self):
        ro = self._rollout
        ro.begin_rollout()
        for n in range(self.horizon_len):
            if not self.env_autoresets and any(self.dones):   # simulators without auto-reset need the indices on the host
                done_indices = torch.where(self.dones)[0].tolist()
                obs_reset = self._convert_obs(self.env.reset_idx(done_indices))
                for k, v in obs_reset.items():
                    self.obs[k][done_indices] = v
            self.storage.update_data('obses', n, self.obs)      # a contiguous device copy (plumbing)
            actions = ro.act(n, self.obs)                          # heads / sample / neglogp / values -> storage rows of step n
            obs, r, self.dones, infos = self.env.step(actions)
            self.obs = self._convert_obs(obs)
            r, self.dones = torch.as_tensor(r, device=self.device), torch.as_tensor(self.dones, device=self.device)
            ro.after_env_step(n, r, self.dones, infos)             # shaped reward (+ time-out bootstrap), dones -> storage rows
            ro.metrics.update(self.epoch, self.env, self.obs, r.reshape(-1).float(), self.dones.reshape(-1), infos)  # no host sync
        self.metrics.flush_video(self.epoch)
        self.finish_rollout(ro.values_only(self.obs))

    @torch.no_grad()
    def finish_rollout(self, last_values):
        """ppo.py:172-183 in one fused pass + one normalise pass."""
        self.storage.compute_return_and_prepare(last_values, self.gamma, self.tau, self.value_rms if self.normalize_value else None,
                                                world=self.world, comm=self.comm)

    def train(self):
        """ppo.py:185-246 without the TensorBoard/W&B plumbing."""
        self.obs = self._convert_obs(self.env.reset())
        self.dones = torch.zeros((self.num_actors,), dtype=torch.bool, device=self.device)
        while self.agent_steps < self.max_agent_steps:
            self.epoch += 1
            self.set_eval()
            self.play_steps()
            self.agent_steps += self.batch_size * self.world
            self.set_train()
            results = self.train_epoch()
            self.storage.data_dict = None
            if self.comm is not None and self.comm.error():   # host read; the metrics below synchronise anyway
                raise RuntimeError(f'rank {self.rank}: a cross-GPU wait timed out (peer {self.comm.error() - 1} died or diverged); '
                                   'the optimizer skipped the affected steps and left the parameters untouched')
            if not self.multi_gpu or self.rank == 0:
                metrics = {f'train_stats/{k}': torch.mean(torch.stack(v)).item() for k, v in results.items() if v[0].dim() == 0}
                metrics['train_stats/last_lr'] = self.last_lr
                metrics['train_scores/episode_rewards'] = self.metrics.episode_trackers['rewards'].mean()
                metrics['train_scores/episode_lengths'] = self.metrics.episode_trackers['lengths'].mean()
                self.writer.add(self.agent_steps, metrics)
                self.writer.write()
                self._checkpoint_save(metrics['train_scores/episode_rewards'])
                if self.print_every > 0 and (self.epoch + 1) % self.print_every == 0:
                    print(f'Epochs: {self.epoch + 1} | Agent Steps: {int(self.agent_steps):,} | '
                          f'ep_rewards {metrics["train_scores/episode_rewards"]:.2f}')
        self.save(os.path.join(self.ckpt_dir, 'final.pth'))

    # -------------------------------------------------------------------------------------------
    def train_epoch(self):
        """ppo.py:248-350: mini_epochs x (B / minibatch) optimizer steps on the prepared data_dict."""
        if self._engine is None:
            self._engine = _UpdateEngine(self)
        results = self._engine.run()
        self.mini_epoch += self.mini_epochs
        return results

    @torch.no_grad()
    def eval(self):
        """ppo.py:352-408: roll the (sampling) policy until 2 x num_actors episodes have finished, on the same kernels as
        `play_steps` (obs normalise -> tower GEMMs -> policy-step kernel; episode accounting on the device), write the
        reference's `eval_scores/*` metrics and `scores.json`."""
        import json
        from ..agent import DeviceMetrics
        self.set_eval()
        obs = self._convert_obs(self.env.reset())
        total_eval_episodes = self.num_actors * 2
        ro = self._rollout_engine()
        scores = {'rewards': torch.zeros(self.num_actors, dtype=torch.float32, device=self.device),
                  'lengths': torch.zeros(self.num_actors, dtype=torch.int64, device=self.device)}
        eval_metrics = DeviceMetrics(scores, total_eval_episodes)
        ro.begin_rollout()
        sink = dict(actions=torch.empty(self.num_actors, self.action_dim, device=self.device), neglogp=torch.empty(self.num_actors, device=self.device),
                    values=torch.empty(self.num_actors, device=self.device), mu=torch.empty(self.num_actors, self.action_dim, device=self.device),
                    sigma=torch.empty(self.num_actors, self.action_dim, device=self.device))
        steps, check_every = 0, 16   # the episode counter lives on the device: read it every few steps, not every step
        while True:
            if not self.env_autoresets:
                raise NotImplementedError
            ro._forward(obs)
            ro.eps.normal_()
            ro._policy(ro.eps, sink['actions'], sink['neglogp'], sink['values'], sink['mu'], sink['sigma'], ro.clamped)
            actions = ro.clamped
            obs, r, dones, infos = self.env.step(actions)
            obs = self._convert_obs(obs)
            r, dones = torch.as_tensor(r, device=self.device), torch.as_tensor(dones, device=self.device)
            eval_metrics.update(self.epoch, self.env, obs, r.reshape(-1).float(), dones.reshape(-1), infos)
            steps += 1
            if steps % check_every == 0 and eval_metrics.num_episodes >= total_eval_episodes:
                break
        metrics = {'eval_scores/num_episodes': eval_metrics.num_episodes,
                   'eval_scores/episode_rewards': eval_metrics.episode_trackers['rewards'].mean(),
                   'eval_scores/episode_lengths': eval_metrics.episode_trackers['lengths'].mean()}
        self.writer.add(self.agent_steps, metrics)
        self.writer.write()
        out = dict(metrics, epoch=self.epoch, mini_epoch=self.mini_epoch, agent_steps=self.agent_steps)
        out['eval_scores/episode_rewards'] = eval_metrics.win['rewards'].tolist()
        out['eval_scores/episode_lengths'] = eval_metrics.win['lengths'].tolist()
        json.dump(out, open(os.path.join(self.logdir, 'scores.json'), 'w'), indent=4)
        return metrics

    def set_train(self):
        self.model.train()

    def set_eval(self):
        self.model.eval()

    def save(self, f):
        """ppo.py:416-427, same keys."""
        ckpt = {
            'epoch': self.epoch,
            'mini_epoch': self.mini_epoch,
            'agent_steps': self.agent_steps,
            'model': self.model.state_dict(),
            'optim': self.optim.state_dict(),
            'obs_rms': self.obs_rms.state_dict() if self.normalize_input else None,
            'value_rms': self.value_rms.state_dict() if self.normalize_value else None,
        }
        torch.save(ckpt, f)

    def load(self, f, ckpt_keys=''):
        """ppo.py:429-446 (regex filter on the checkpoint keys)."""
        ckpt = torch.load(f, map_location=self.device, weights_only=False)
        for k in ('epoch', 'mini_epoch', 'agent_steps', 'model', 'optim', 'obs_rms', 'value_rms'):
            if not re.match(ckpt_keys, k):
                print(f'Warning: ckpt skipped loading `{k}`')
                continue
            if (k == 'obs_rms' and not self.normalize_input) or (k == 'value_rms' and not self.normalize_value):
                continue
            if hasattr(getattr(self, k), 'load_state_dict'):
                getattr(self, k).load_state_dict(ckpt[k])
            else:
                setattr(self, k, ckpt[k])
        self.last_lr = self.optim.param_groups[0]['lr']


class _RolloutEngine:
    """Agent side of one env step (ppo.py:144-163 minus env.step and the host-side episode metrics) on our kernels:
    `mnr_rms_apply_ld_f32` (obs normalise) -> tower GEMMs (`TowerStack.forward`, hi / lo weights split once per rollout) ->
    `mnr_policy_step` -> [env.step] -> `mnr_post_step`.  ~7 launches per env step instead of ~50 torch ops."""

    @staticmethod
    def supported(agent):
        from ...nets.fused_towers import FusedTowers
        return agent.model.mu.weight.shape[1] <= 512 and agent.reward_shaper.fn == 'scale'   # the rest: PPO.__init__

    def __init__(self, agent):
        from ...nets.fused_towers import FusedTowers
        a = self.a = agent
        dev, N, A = a.device, a.num_actors, a.action_dim
        self.N, self.A = N, A
        self.towers = FusedTowers(a.model, N, dev)
        f32 = dict(dtype=torch.float32, device=dev)
        self.in_dim = a.obs_space['obs'][0]
        self.x_ld = (self.in_dim + 3) // 4 * 4
        self.x = torch.zeros(N, self.x_ld, **f32)       # normalised obs, TMA-friendly row stride (pad columns stay zero)
        self.eps = torch.empty(N, A, **f32)
        self.clamped = torch.empty(N, A, **f32)
        self.scratch = [torch.empty(N, A, **f32) for _ in range(3)] + [torch.empty(N, **f32)]   # values_only() sinks
        self.last_values = torch.empty(N, 1, **f32)
        dk = a.model.dist
        self.cfg = ops.make_loss_cfg(dist_type=dk.dist_type, minstd=dk.minstd, maxstd=dk.maxstd)
        self.eps_source = None   # tests inject recorded noise: callable(step) -> (N, A) tensor; default torch's generator
        # episode accounting on the device; the agent's `metrics` keeps the reference's attributes (`episode_trackers`, ...)
        from ..agent import DeviceMetrics, Metrics
        if type(a.metrics) is Metrics:
            a.metrics = DeviceMetrics(a.metrics.current, a.tracker_len)
        self.metrics = a.metrics

    def begin_rollout(self):
        a = self.a
        self.obs_tab = a.obs_rms['obs']._prepare() if (a.normalize_input and hasattr(a.obs_rms['obs'], 'running_mean')) else None
        self.val_tab = a.value_rms._prepare() if a.normalize_value else (None, None)
        st = self.towers.stack
        if st.use_tc:
            st._split_weights()    # the weights are frozen for the whole rollout

    def _forward(self, obs):
        a, tw = self.a, self.towers
        o = obs['obs']
        if self.obs_tab is not None:
            rms = a.obs_rms['obs']
            ops.rms_apply_ld(o.contiguous(), self.obs_tab[0], self.obs_tab[1], tuple(rms.clamp_range) if rms.with_clamp else None,
                             self.x, self.x_ld)
        else:
            self.x[:, :self.in_dim].copy_(o)
        tw.stack.forward(self.x, self.x_ld, rows=self.N, split=False)

    def _policy(self, eps, actions, neglogp, values, mu, sigma, clamped, unnormalize_value=True):
        a, m, tw = self.a, self.a.model, self.towers
        fixed = m.fixed_sigma
        val_tab = self.val_tab if unnormalize_value else (None, None)
        clamp = tuple(a.value_rms.clamp_range) if (val_tab[0] is not None and a.value_rms.with_clamp) else None
        ops.policy_step(tw.Y[-1][0], tw.Y[-1][1], m.mu.weight.data, m.mu.bias.data, m.value.weight.data, m.value.bias.data,
                        None if fixed else m.sigma.weight.data, m.sigma.data if fixed else m.sigma.bias.data, eps,
                        val_tab[0], val_tab[1], clamp, actions, neglogp, values, mu, sigma, clamped, self.cfg, rows=self.N)

    def act(self, t, obs):
        sd = self.a.storage.storage_dict
        self._forward(obs)
        if self.eps_source is not None:
            self.eps.copy_(self.eps_source(t))
        else:
            self.eps.normal_()   # the reference's Normal.sample() draws the same way: randn, then * sigma + mu
        self._policy(self.eps, sd['actions'][t], sd['neglogp'][t], sd['values'][t].view(-1), sd['mu'][t], sd['sigma'][t], self.clamped)
        return self.clamped

    def after_env_step(self, t, rewards, dones, infos):
        a = self.a
        sd = a.storage.storage_dict
        tout = None
        if a.value_bootstrap and 'time_outs' in infos:
            tout = torch.as_tensor(infos['time_outs'], device=a.device).reshape(-1).contiguous()
            if tout.dtype not in (torch.bool, torch.uint8):
                tout = tout != 0
        d = dones.reshape(-1).contiguous()
        if d.dtype not in (torch.bool, torch.uint8):
            d = d != 0
        ops.post_step(rewards.reshape(-1).float().contiguous(), d, tout, sd['values'][t].view(-1), float(a.reward_shaper.scale), float(a.gamma),
                      sd['rewards'][t].view(-1), sd['dones'][t])

    def values_only(self, obs):
        """Bootstrap values of the final observation (ppo.py:169-171): same kernels, the other outputs go to scratch."""
        self._forward(obs)
        self.eps.zero_()
        s = self.scratch
        self._policy(self.eps, s[0], s[3], self.last_values.view(-1), s[1], s[2], None)
        return self.last_values


class _UpdateEngine:
    """Pre-allocated buffers + the launch sequence of one PPO update (optionally one CUDA graph)."""

    def __init__(self, agent):
        a = self.a = agent
        dev = a.device
        self.mb, self.B, self.A = a.minibatch_size, a.batch_size, a.action_dim
        self.num_mb = len(a.storage)
        self.steps = a.mini_epochs * self.num_mb
        f32 = dict(dtype=torch.float32, device=dev)
        self.metrics = torch.zeros(self.steps, PPO_METRICS, **f32)
        self.mu_mean = torch.zeros(self.steps, self.A, **f32)
        self.sigma_mean = torch.zeros(self.steps, self.A, **f32)
        self.grad_norms = torch.zeros(self.steps, **f32)
        self.avg_kl = torch.zeros(a.mini_epochs, **f32)
        self.lr_out = torch.zeros(a.mini_epochs, dtype=torch.float64, device=dev)
        self.g_mu = torch.zeros(self.mb, self.A, **f32)
        self.g_ls = torch.zeros(self.mb, self.A, **f32)
        self.g_v = torch.zeros(self.mb, 1, **f32)
        self.loss_ws = ops.ppo_loss_workspace(self.mb, self.A, dev)
        dk = a.model.dist
        self.loss_cfg = ops.make_loss_cfg(
            dist_type=dk.dist_type, minstd=dk.minstd, maxstd=dk.maxstd, e_clip=a.e_clip, use_smooth_clamp=a.use_smooth_clamp,
            clip_value_loss=a.clip_value_loss, critic_coef=a.critic_coef, entropy_coef=a.entropy_coef,
            bounds_loss_coef=a.bounds_loss_coef, bounds_type=a.bounds_type, grad_scale=1.0 / a.world, metric_scale=1.0 / a.world)
        self.norm_keys = []
        if a.normalize_input:
            for k, shape in a.obs_space.items():
                if not hasattr(a.obs_rms[k], 'running_mean'):
                    continue
                cols = a.obs_rms[k]._cols
                self.norm_keys.append(k)
                setattr(self, f'_n_{k}', dict(
                    cols=cols, gm=torch.zeros(self.num_mb, cols, **f32), gv=torch.zeros(self.num_mb, cols, **f32),
                    mean_tab=torch.zeros(self.steps, cols, **f32), denom_tab=torch.zeros(self.steps, cols, **f32),
                    x=torch.zeros(self.mb, (cols + 3) // 4 * 4, **f32), x_ld=(cols + 3) // 4 * 4,
                    x2=torch.zeros(self.mb, (cols + 3) // 4 * 4, **f32),   # second buffer: step s + 1 is normalised under step s
                    ws=ops.rms_moments_workspace(self.B, cols, self.mb, dev)))
        # actor / value towers: explicit forward / backward on the hand-written linear kernels (tcgen05 where the widths
        # allow, SIMT twins otherwise); PPO.__init__ has already rejected networks FusedTowers cannot run
        from ...nets.fused_towers import FusedTowers
        self.towers = FusedTowers(a.model, self.mb, dev)
        if self.towers.fused_heads_ok:
            self.heads_ws = ops.ppo_heads_workspace(self.mb, self.A, self.towers.widths[-1], dev)
            # ELU / ReLU towers: the heads kernel writes dZ3 = dY3 * act'(Y3) (it holds the Y3 tiles), the top layer's
            # backward GEMMs then skip their saved-output loads
            self.loss_cfg.dz_act = self.towers.heads_dz_act()
        jobs = self.towers.stack.split_jobs()
        if 0 < len(jobs) <= 8:   # the optimizer kernel keeps the tf32 (hi, lo) weight copies current
            a.optim.attach_splits(jobs)
        if not self.norm_keys:
            cols = a.obs_space['obs'][0]
            self._raw_x = torch.zeros(self.mb, (cols + 3) // 4 * 4, **f32)
        # 'cos' / 'linear': the learning rate that takes effect after each mini-epoch, uploaded by run() before the launches
        self.lr_table = torch.zeros(a.mini_epochs, dtype=torch.float64, device=dev)
        self.lr_table_host = torch.zeros(a.mini_epochs, dtype=torch.float64).pin_memory()
        self.graph = None
        self._warm = 0
        self.graph_launches = 0  # kernels of ours inside one captured update
        self.replays = 0

    # ---- the launch sequence -------------------------------------------------------------------
    def _enqueue(self):
        a = self.a
        d = a.storage.data_dict
        mb = self.mb
        # obs-RMS: per-minibatch moments once, then every step's statistics up front (ppo.py:258-262)
        for k in self.norm_keys:
            n, rms = getattr(self, f'_n_{k}'), a.obs_rms[k]
            ops.rms_batch_moments(d['obses'][k].view(self.B, n['cols']), mb, rms.correction, n['ws'], n['gm'], n['gv'])
            if a.world > 1:
                self._merge_moments_across_ranks(n)
            ops.rms_schedule(rms.running_mean, rms.running_var, rms.running_count, n['gm'], n['gv'], mb * a.world,
                             self.B * a.world, self.steps, rms.eps, n['mean_tab'], n['denom_tab'])
        max_norm = float(a.max_grad_norm) if a.truncate_grads else 0.0
        step = 0
        peer = a.comm is not None
        # tf32 (hi, lo) weight copies: split ONCE here (whatever changed the weights since the last update: a checkpoint
        # load, a test poking parameters); afterwards the optimizer kernel keeps them current (FusedAdam.attach_splits)
        opt_splits = self.towers.use_tc and getattr(a.optim, 'splits_attached', False)
        if opt_splits:
            self.towers.stack._split_weights()
        n, rms = (getattr(self, f'_n_{self.norm_keys[0]}'), a.obs_rms[self.norm_keys[0]]) if self.norm_keys else (None, None)
        clamp = tuple(rms.clamp_range) if (rms is not None and rms.with_clamp) else None
        if n is None:   # no obs normalisation: the towers read the rows through the same padded stride
            x_buf, x_ld = self._raw_x, self._raw_x.shape[1]
        else:
            x_buf, x_ld = n['x'], n['x_ld']

        pingpong = peer and os.environ.get('MNR_PEER_END_BARRIER', '0') != '1'   # =1: two barriers per step (A/B measurements)
        # MNR_RMS_FORK=1 (A/B measurements): the normalised input of step s + 1 depends only on the statistics tables computed
        # above, so its launch can be forked onto a side stream at the top of step s (two input buffers) and join before the
        # forward of step s + 1.  MEASURED SLOWER (round 2, cfg2: 5.02 ms per update against 4.87 ms in line): the persistent
        # GEMM kernels own every SM, so the forked kernel's CTAs only delay their tiles, and the cross-stream graph edges
        # cost more than the 7 us launch they hide.  Default: in line.
        fork = n is not None and os.environ.get('MNR_RMS_FORK', '0') == '1'
        # Default: the normalisation of step s + 1's rows rides along in step s's optimizer launch (extra CTAs beside the
        # optimizer's own, mnr_optim_queue_normalise): only the first step of an update has its own launch.
        # MNR_RMS_RIDE=0: one mnr_rms_apply_ld_f32 launch per step (A/B measurements).
        ride = n is not None and not fork and os.environ.get('MNR_RMS_RIDE', '1') != '0'
        main = torch.cuda.current_stream()
        if fork:
            if getattr(self, '_side', None) is None:
                self._side = torch.cuda.Stream(device=a.device)
            xs = (n['x'], n['x2'])
        v = d['obses']['obs']
        total = a.mini_epochs * self.num_mb

        def normalise(st_, buf):
            j = st_ % self.num_mb
            ops.rms_apply_ld(v[j * mb:(j + 1) * mb], n['mean_tab'][st_], n['denom_tab'][st_], clamp, buf, x_ld)

        ready = None
        if fork:
            normalise(0, xs[0])
        for ep in range(a.mini_epochs):
            for i in range(self.num_mb):
                s, e = i * mb, (i + 1) * mb
                if peer:   # this step's backward writes ping-pong gradient region step & 1
                    a.optim.use_grad_buffer(step)
                if fork:
                    x_buf = xs[step & 1]
                    if ready is not None:
                        main.wait_event(ready)
                    if step + 1 < total:   # xs[(step + 1) & 1] was last read by step - 1, which is complete in `main` order
                        self._side.wait_stream(main)
                        with torch.cuda.stream(self._side):
                            normalise(step + 1, xs[(step + 1) & 1])
                            ready = torch.cuda.Event()
                            ready.record(self._side)
                elif n is not None:   # normalise + clamp into the padded (TMA-friendly) row stride; pad columns stay zero
                    if not ride or step == 0:
                        ops.rms_apply_ld(v[s:e], n['mean_tab'][step], n['denom_tab'][step], clamp, x_buf, x_ld)
                else:
                    x_buf[:, :v.shape[1]].copy_(v[s:e])
                m = self.metrics[step]
                if self.towers.fused_heads_ok:
                    # towers -> fused heads + loss (+ head grads, d/dY3) -> towers backward
                    self.towers.forward_towers(x_buf, x_ld, split=not opt_splits)
                    # the cross-CTA sums of the heads kernel and of LayerNorm's backward run inside the split-reduce launch
                    folds = self.towers.stack.fold_begin()
                    try:
                        self.towers.heads_loss_fwd_bwd(d, s, e, self.g_mu, self.g_ls, self.g_v.view(-1), a.optim.tail if peer else m,
                                                       self.loss_cfg, self.heads_ws, self.mu_mean[step], self.sigma_mean[step])
                        self.towers.backward_towers(x_buf, folds=folds)
                    finally:   # never leave a fold list registered with the library (backward clears it on the normal path)
                        if folds is not None:
                            ops.fold_defer(None)
                else:
                    # heads wider than the fused kernel takes: thin-layer head kernels + the row-wise loss kernel
                    self.towers.stack.forward(x_buf, x_ld, split=not opt_splits)
                    mu_raw, logstd, value = self.towers.heads_forward()
                    ops.ppo_loss_fwd_bwd(mu_raw, logstd.contiguous(), value, d['actions'][s:e], d['neglogp'][s:e], d['advantages'][s:e],
                                         d['values'][s:e], d['returns'][s:e], d['mu'][s:e], d['sigma'][s:e], self.g_mu, self.g_ls,
                                         self.g_v, a.optim.tail if peer else m, self.loss_cfg, self.loss_ws, True,
                                         self.mu_mean[step], self.sigma_mean[step])
                    self.towers.backward(x_buf, self.g_mu, self.g_ls, self.g_v)
                if ride and step + 1 < total:   # x_buf was last read by this step's first-layer weight gradient: done by then
                    j = (step + 1) % self.num_mb
                    ops.optim_queue_normalise(v[j * mb:(j + 1) * mb], n['mean_tab'][step + 1], n['denom_tab'][step + 1], clamp, x_buf, x_ld)
                # clip_grad_norm_ + Adam(W) + zero_grad: one launch; with peers also the gradient + metrics all-reduce over NVLink
                if peer:   # the last step of the update keeps the second barrier: its region is the next update's first target
                    a.optim.step(max_norm, grad_norm_out=self.grad_norms[step:step + 1], metrics_out=m, kl_index=7,
                                 end_barrier=(not pingpong) or step == self.steps - 1)
                else:
                    a.optim.step(max_norm, kl=m[7:8], grad_norm_out=self.grad_norms[step:step + 1])
                step += 1
            if a.lr_schedule == 'kl':
                ops.kl_lr_update(a.optim.state, a.kl_threshold, a.min_lr, a.max_lr, self.avg_kl[ep:ep + 1], self.lr_out[ep:ep + 1])
            elif a.lr_schedule in ('cos', 'linear'):
                ops.lr_table_update(a.optim.state, self.lr_table[ep:ep + 1], self.avg_kl[ep:ep + 1], self.lr_out[ep:ep + 1])
            else:
                # 'fixed': a NaN threshold makes both comparisons of the KL rule false, so it only averages the KL and reports lr
                ops.kl_lr_update(a.optim.state, float('nan'), 0.0, 1.0, self.avg_kl[ep:ep + 1], self.lr_out[ep:ep + 1])

    def _merge_moments_across_ranks(self, n):
        """Global minibatch i = union of the ranks' local minibatch i; peer-memory all-gather."""
        from ... import distributed as mdist
        mdist.merge_group_moments(n['gm'], n['gv'], self.mb, self.a.obs_rms[self.norm_keys[0]].correction, comm=self.a.comm)

    def _upload_lr_table(self):
        """ppo.py:339-348.  'cos': lr after mini-epoch k = adjust_learning_rate_cos(init_lr, k, mini_epochs, agent_steps,
        max_agent_steps).  'linear': the reference re-installs `last_lr` (the value its scheduler produced at the END of the
        previous train_epoch) after every mini-epoch and only then computes the next one, so the new value takes effect one
        mini-epoch into the next update - reproduced by installing the same `last_lr` after every mini-epoch."""
        a = self.a
        if a.lr_schedule == 'cos':
            vals = [adjust_learning_rate_cos(a.init_lr, k, a.mini_epochs, a.agent_steps, a.max_agent_steps) for k in range(a.mini_epochs)]
        else:
            vals = [float(a._last_lr)] * a.mini_epochs
        self.lr_table_host.copy_(torch.tensor(vals, dtype=torch.float64))
        self.lr_table.copy_(self.lr_table_host, non_blocking=True)   # stream-ordered before the launches / the graph replay
        return vals

    def run(self):
        a = self.a
        if a.lr_schedule in ('cos', 'linear'):
            self._host_lr = self._upload_lr_table()[-1]
        with torch.no_grad():
            if a.use_cuda_graph:  # no NCCL call inside the update: capturable at any world size
                if self.graph is None and self._warm >= 1:
                    torch.cuda.synchronize()
                    from ... import _native as nv
                    before = nv.LAUNCHES
                    self.graph = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(self.graph):
                        self._enqueue()
                    self.graph_launches = nv.LAUNCHES - before
                    nv.LAUNCHES = before  # captured, not launched; replays are counted separately
                    # the capture itself did not execute anything; fall through to replay
                if self.graph is not None:
                    self.graph.replay()
                    self.replays += 1
                else:
                    self._enqueue()
                    self._warm += 1
            else:
                self._enqueue()
        return self._results()

    def _results(self):
        a = self.a
        # The learning rate lives on the device (KL schedule, mnr_kl_lr_update).  `agent.last_lr` reads it back ON DEMAND, so an
        # update ends without a host synchronisation and the host-side assembly below overlaps the GPU's work.
        a._lr_stale = True
        if a.lr_schedule == 'cos':
            a.last_lr = self._host_lr                       # known on the host: no read-back needed
        elif a.lr_schedule == 'linear':
            a.last_lr = a.scheduler.update(a.agent_steps)   # ppo.py:347-348 (takes effect after the next mini-epoch)
        m = self.metrics
        res = collections.defaultdict(list)
        names = ('loss/total', 'loss/actor', 'loss/critic', 'loss/bounds', 'loss/entropy', 'clip_frac', 'explained_var')
        # per-step 0-dim views, one unbind per key (the reference's lists of per-step scalars, without 12 x steps indexing ops)
        for j, k in enumerate(names):
            res[k] = list(m[:, j].unbind(0))
        res['kl'] = list(m[:, 7].unbind(0))
        # column means (1, A): torch.cat(...).mean(0) over them equals the reference's value
        res['mu'] = list(self.mu_mean.unsqueeze(1).unbind(0))
        res['sigma'] = list(self.sigma_mean.unsqueeze(1).unbind(0))
        if a.truncate_grads:
            res['grad_norm/all'] = list(self.grad_norms.unbind(0))
        res['avg_kl'] = list(self.avg_kl.unbind(0))
        return res


# -----------------------------------------------------------------------------------------------
# Module-level loss functions with the reference's signatures (ppo.py:449-498), each a thin
# autograd.Function over the per-row C-ABI kernels.
# -----------------------------------------------------------------------------------------------
def _rows_ws(n, dev):
    from ..._native import lib, workspace
    return workspace('rows_reduce', lib.mnr_rows_reduce_workspace_bytes(n), dev)


class _ActorLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, old_nlp, nlp, adv, e_clip, smooth):
        from ..._native import check, lib, ptr, stream
        n = nlp.numel()
        loss, d = torch.empty_like(nlp), torch.empty_like(nlp)
        frac = torch.empty((), dtype=torch.float32, device=nlp.device)
        ws = _rows_ws(n, nlp.device)
        check(lib.mnr_actor_loss_rows(ptr(old_nlp.contiguous()), ptr(nlp.contiguous()), ptr(adv.contiguous()), n, e_clip,
                                      int(bool(smooth)), ptr(loss), ptr(d), ptr(frac), ptr(ws), ws.numel(), stream()))
        ctx.save_for_backward(d)
        ctx.mark_non_differentiable(frac)
        return loss, frac

    @staticmethod
    def backward(ctx, g_loss, g_frac):
        (d,) = ctx.saved_tensors
        return None, g_loss * d, None, None, None


def actor_loss(old_action_log_probs, action_log_probs, advantage, e_clip, use_smooth_clamp):
    return _ActorLoss.apply(old_action_log_probs, action_log_probs, advantage, e_clip, use_smooth_clamp)


class _CriticLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, value_preds, values, e_clip, returns, clip):
        from ..._native import check, lib, ptr, stream
        n = values.numel()
        loss, d = torch.empty_like(values), torch.empty_like(values)
        ev = torch.empty((), dtype=torch.float32, device=values.device)
        ws = _rows_ws(n, values.device)
        check(lib.mnr_critic_loss_rows(ptr(value_preds.contiguous()), ptr(values.contiguous()), ptr(returns.contiguous()), n,
                                       e_clip, int(bool(clip)), ptr(loss), ptr(d), ptr(ev), ptr(ws), ws.numel(), stream()))
        ctx.save_for_backward(d)
        ctx.mark_non_differentiable(ev)
        return loss, ev

    @staticmethod
    def backward(ctx, g_loss, g_ev):
        (d,) = ctx.saved_tensors
        return None, g_loss * d, None, None, None


def critic_loss(value_preds, values, e_clip, returns, clip_value_loss):
    return _CriticLoss.apply(value_preds, values, e_clip, returns, clip_value_loss)


class _BoundsLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mu, btype, soft_bound):
        from ..._native import check, lib, ptr, stream
        n, A = mu.shape
        loss = torch.empty(n, dtype=torch.float32, device=mu.device)
        d = torch.empty_like(mu)
        check(lib.mnr_bounds_loss_rows(ptr(mu.contiguous()), n, A, btype, soft_bound, ptr(loss), ptr(d), stream()))
        ctx.save_for_backward(d)
        return loss

    @staticmethod
    def backward(ctx, g):
        (d,) = ctx.saved_tensors
        return g.unsqueeze(-1) * d, None, None


def bounds_loss(mu, bounds_type, soft_bound=1.1):
    bt = {'bound': 0, 'reg': 1}.get(bounds_type)
    if bt is None:
        raise NotImplementedError(bounds_type)
    return _BoundsLoss.apply(mu, bt, soft_bound)


@torch.no_grad()
def policy_kl(p0_mu, p0_sigma, p1_mu, p1_sigma):
    from ..._native import check, lib, ptr, stream
    n, A = p0_mu.shape
    out = torch.empty((), dtype=torch.float32, device=p0_mu.device)
    ws = _rows_ws(n, p0_mu.device)
    check(lib.mnr_policy_kl_rows(ptr(p0_mu.contiguous()), ptr(p0_sigma.contiguous()), ptr(p1_mu.contiguous()),
                                 ptr(p1_sigma.contiguous()), n, A, None, ptr(out), ptr(ws), ws.numel(), stream()))
    return out
