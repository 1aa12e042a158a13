"""Agent base: the drop-in boundary of the reference (mineral/agents/agent.py:13-167) reduced to what
the rollout->update path needs.  Constructor signature, public attributes (`epoch`, `mini_epoch`,
`agent_steps`, `ckpt_dir`, `device`, `obs_space`, `action_dim`, ...) and the checkpoint policy are
kept; TensorBoard / W&B writers and video capture (observability)
are out of scope and replaced by an in-memory writer with the same `add` / `write` calls."""

import os
import re
import tempfile
from collections import deque

import numpy as np
import torch


def cfg_get(cfg, key, default=None):
    if cfg is None:
        return default
    if hasattr(cfg, 'get'):
        return cfg.get(key, default)
    return getattr(cfg, key, default)


class Tracker:
    """Windowed statistics of finished episodes (mineral/common/tracker.py:9-22): like the reference, the window starts
    as `max_len` zeros, so the mean ramps up until `max_len` episodes have finished."""

    def __init__(self, max_len):
        self.max_len = max_len
        self.window = deque([0 for _ in range(max_len)], maxlen=max_len)

    def update(self, values):
        self.window.extend(values)

    def mean(self):
        return float(np.mean(self.window))


class DeviceTracker:
    """Tracker whose window lives on the GPU (filled by mnr_episode_scores); reading a statistic is the only sync."""

    def __init__(self, window):
        self.window = window   # (max_len,) f32, zero-initialised; order inside the ring does not matter for mean / std / ...

    def mean(self):
        return float(self.window.mean().item())


class DeviceMetrics:
    """Metrics._update_scores (mineral/common/metrics.py:59-67) without the per-step host round trip: ONE kernel per env
    step keeps the per-env accumulators and the windows of the last `tracker_len` finished episodes on the device."""

    def __init__(self, current_scores, tracker_len):
        from .. import ops
        self._ops = ops
        self.current = current_scores
        dev = current_scores['rewards'].device
        self.win = {k: torch.zeros(tracker_len, dtype=torch.float32, device=dev) for k in ('rewards', 'lengths')}
        self.episode_trackers = {k: DeviceTracker(v) for k, v in self.win.items()}
        self.counters = torch.zeros(1, dtype=torch.int64, device=dev)

    @property
    def num_episodes(self):
        return int(self.counters.item())

    def update(self, epoch, env, obs, rewards, dones, infos):
        """`dones`: the (N,) bool / u8 done flags of this step (NOT a host list of indices)."""
        self._ops.episode_scores(rewards.contiguous(), dones.contiguous(), self.current['rewards'], self.current['lengths'],
                                 self.win['rewards'], self.win['lengths'], self.counters)

    def flush_video(self, epoch):
        return

    def result(self, prefix=''):
        return {}


class Metrics:
    """Per-env episode reward/length accumulators on the device (mineral/common/metrics.py:59-67)."""

    def __init__(self, current_scores, tracker_len):
        self.current = current_scores
        self.episode_trackers = {k: Tracker(tracker_len) for k in current_scores}
        self.num_episodes = 0

    def update(self, epoch, env, obs, rewards, done_indices, infos):
        self.current['rewards'] += rewards
        self.current['lengths'] += 1
        if len(done_indices):
            idx = torch.as_tensor(done_indices, device=rewards.device)
            for k in self.current:
                self.episode_trackers[k].update(self.current[k][idx].tolist())
                self.current[k][idx] = 0
            self.num_episodes += len(done_indices)

    def flush_video(self, epoch):
        return

    def result(self, prefix=''):
        return {}


class MemoryWriter:
    def __init__(self):
        self.rows = []

    def add(self, step, metrics):
        self.rows.append((step, metrics))

    def write(self):
        return


class Agent:
    def __init__(self, full_cfg, logdir=None, accelerator=None, datasets=None, env=None):
        assert getattr(self, 'network_config', False)
        assert getattr(self, 'num_actors', False)
        self.full_cfg = full_cfg
        self.logdir = logdir if logdir is not None else tempfile.mkdtemp(prefix='mineral_b200_')

        # --- Device / ranks (agent.py:22-31).  One process per GPU over torch.distributed (NCCL);
        #     the reference's `accelerator` argument is accepted and ignored. ---
        self.device = full_cfg.rl_device
        self.multi_gpu = bool(cfg_get(full_cfg, 'multi_gpu', False))
        self.rank = int(os.getenv('LOCAL_RANK', '0')) if self.multi_gpu else -1
        self.rank_size = int(os.getenv('WORLD_SIZE', '1')) if self.multi_gpu else 1
        self.accelerator = accelerator
        self.datasets = datasets

        # --- Environment (agent.py:37-51) ---
        self.env = env
        self.action_dim = self.env.action_space.shape[0]
        self.env_autoresets = cfg_get(full_cfg.task, 'env_autoresets', True)
        self.normalize_input = cfg_get(self.network_config, 'normalize_input', False)
        self.obs_rms_keys = re.compile(cfg_get(self.network_config, 'obs_rms_keys', ''))
        self.cpu_obs_keys = re.compile(cfg_get(self.network_config, 'cpu_obs_keys', '$^'))
        self.observation_space = self.env.observation_space
        try:
            obs_space = {k: v.shape for k, v in self.observation_space.spaces.items()}
        except AttributeError:
            obs_space = {'obs': self.observation_space.shape}
        self.obs_space = obs_space

        # --- Metrics / logging ---
        self.tracker_len = cfg_get(full_cfg.agent, 'tracker_len', 100)
        self.metrics = self._create_metrics(self.tracker_len)
        self.ckpt_dir = os.path.join(self.logdir, 'ckpt')
        os.makedirs(self.ckpt_dir, exist_ok=True)
        self.writer = MemoryWriter()
        self.print_every = cfg_get(full_cfg.agent, 'print_every', -1)
        self.ckpt_every = cfg_get(full_cfg.agent, 'ckpt_every', -1)
        self.best_stat = None

        # --- Training counters (agent.py:79-81) ---
        self.epoch = -1
        self.mini_epoch = -1
        self.agent_steps = 0

    def _create_metrics(self, tracker_len):
        scores = {'rewards': torch.zeros(self.num_actors, dtype=torch.float32, device=self.device),
                  'lengths': torch.zeros(self.num_actors, dtype=torch.int64, device=self.device)}
        return Metrics(scores, tracker_len)

    def _checkpoint_save(self, stat, stat_name='rewards', higher_better=True, sep=''):
        """agent.py:124-142: periodic + rolling-best checkpoints."""
        if self.ckpt_every > 0 and (self.epoch + 1) % self.ckpt_every == 0:
            name = f'epochs{sep}{self.epoch + 1}_steps{sep}{int(self.agent_steps / 1000)}k_{stat_name}{sep}{stat:.2f}.pth'
            self.save(os.path.join(self.ckpt_dir, name))
            latest = os.path.join(self.ckpt_dir, 'latest.pth')
            if os.path.lexists(latest):
                os.unlink(latest)
            os.symlink(name, latest)
        better = True if self.best_stat is None else (stat > self.best_stat if higher_better else stat < self.best_stat)
        if better:   # unconditional, as in the reference (agent.py:133-142): the rolling best does not depend on ckpt_every
            if self.best_stat is not None:
                prev = os.path.join(self.ckpt_dir, f'best_{stat_name}{sep}{self.best_stat:.2f}.pth')
                if os.path.exists(prev):
                    os.remove(prev)
            self.best_stat = stat
            self.save(os.path.join(self.ckpt_dir, f'best_{stat_name}{sep}{self.best_stat:.2f}.pth'))

    def _convert_obs(self, obs):
        """agent.py:144-156."""
        if not isinstance(obs, dict):
            obs = {'obs': obs}
        out = {}
        for k, v in obs.items():
            out[k] = torch.tensor(v, device=self.device) if isinstance(v, np.ndarray) else v
        return out
