"""Circular replay buffer on the C-ABI row kernels.  Mirrors mineral/buffers/simple_replay.py:4-85
(same ctor, attributes `next_p / if_full / cur_capacity / capacity / buf_*`, same wrap rule, indices
drawn with `torch.randint` so sampling is bit-identical given the same generator state)."""

import math

import torch

from .. import ops
from .._native import ROW_BOOL_TO_F32, ROW_F32, ROW_F32_NORMALIZE, ROW_F32_TO_BOOL, ROW_U8


def create_buffer(capacity, obs_space, action_dim, device='cuda'):
    """simple_replay.py:4-16: f32 obs/action/reward/next_obs, BOOL done."""
    if isinstance(capacity, int):
        capacity = (capacity,)
    buf_obs = {k: torch.empty((*capacity, *v), dtype=torch.float32, device=device) for k, v in obs_space.items()}
    buf_action = torch.empty((*capacity, int(action_dim)), dtype=torch.float32, device=device)
    buf_reward = torch.empty((*capacity, 1), dtype=torch.float32, device=device)
    buf_next_obs = {k: torch.empty((*capacity, *v), dtype=torch.float32, device=device) for k, v in obs_space.items()}
    buf_done = torch.empty((*capacity, 1), dtype=torch.bool, device=device)
    return buf_obs, buf_action, buf_next_obs, buf_reward, buf_done


def _width(shape):
    return int(math.prod(shape)) if len(shape) else 1


class ReplayBuffer:
    def __init__(self, obs_space, action_dim: int, capacity: int, device='cuda'):
        self.obs_space = obs_space
        self.action_dim = action_dim
        self.device = device
        self.next_p = 0
        self.if_full = False
        self.cur_capacity = 0
        self.capacity = int(capacity)
        ret = create_buffer(self.capacity, obs_space, action_dim, device=device)
        self.buf_obs, self.buf_action, self.buf_next_obs, self.buf_reward, self.buf_done = ret

    @torch.no_grad()
    def add_to_buffer(self, trajectory):
        """Ring insert, one launch for all fields (simple_replay.py:32-71)."""
        obs, actions, rewards, next_obs, dones = trajectory
        n = rewards.numel()
        fields = []
        for k in self.buf_obs:
            fields.append((obs[k].contiguous(), self.buf_obs[k], _width(self.obs_space[k]), ROW_F32))
        fields.append((actions.contiguous(), self.buf_action, self.action_dim, ROW_F32))
        fields.append((rewards.contiguous(), self.buf_reward, 1, ROW_F32))
        for k in self.buf_next_obs:
            fields.append((next_obs[k].contiguous(), self.buf_next_obs[k], _width(self.obs_space[k]), ROW_F32))
        dones = dones.contiguous()
        if dones.dtype == torch.float32:
            fields.append((dones, self.buf_done, 1, ROW_F32_TO_BOOL))  # dones.bool(), simple_replay.py:39
        elif dones.dtype in (torch.bool, torch.uint8):
            fields.append((dones, self.buf_done, 1, ROW_U8))
        else:
            raise TypeError(f'dones dtype {dones.dtype}')
        ops.replay_insert(fields, n, self.next_p, self.capacity)

        p = self.next_p + n
        if p > self.capacity:
            self.if_full = True
            p -= self.capacity
        self.next_p = p
        self.cur_capacity = self.capacity if self.if_full else self.next_p

    @torch.no_grad()
    def sample_batch(self, batch_size, device='cuda', indices=None, obs_rms=None, out=None):
        """Uniform sample with replacement (simple_replay.py:73-85): indices from torch.randint on
        `device`, then ONE gather launch for obs / action / reward / next_obs / done->f32.

        Extensions: `indices` replays a given int64 index tensor; `obs_rms` ({key: RunningMeanStd})
        additionally returns normalised+clamped obs / next_obs fused into the same gather (what
        SAC/DDPG do 2-3x per batch downstream, mineral/agents/sac/sac.py:289-301,317-319)."""
        if indices is None:
            indices = torch.randint(self.cur_capacity, size=(batch_size,), device=device)
        n = indices.numel()
        f32 = dict(dtype=torch.float32, device=self.buf_action.device)
        out = {} if out is None else out
        # `out`: destination views for 'obs' / 'next_obs' ({key: view}), 'action', 'reward', 'done' - e.g. column blocks of the
        # padded (batch, obs+act) critic-input buffers, so the `torch.cat((state, action), 1)` of EnsembleQ.forward
        # (mineral/agents/ddpg/models.py:102-106) costs nothing.  Views need unit column stride; rows may be strided.
        obs = {k: out['obs'][k] if 'obs' in out else torch.empty((n, *self.obs_space[k]), **f32) for k in self.buf_obs}
        next_obs = {k: out['next_obs'][k] if 'next_obs' in out else torch.empty((n, *self.obs_space[k]), **f32) for k in self.buf_next_obs}
        action = out['action'] if 'action' in out else torch.empty((n, self.action_dim), **f32)
        reward = out['reward'] if 'reward' in out else torch.empty((n, 1), **f32)
        done = out['done'] if 'done' in out else torch.empty((n, 1), **f32)

        def ld(t):
            return 0 if t.is_contiguous() else int(t.stride(0))
        kind = ROW_F32
        fields = []
        for k in self.buf_obs:
            w = _width(self.obs_space[k])
            if obs_rms is not None and hasattr(obs_rms[k], 'running_mean'):
                m, d = obs_rms[k]._prepare()
                lo, hi = obs_rms[k].clamp_range if obs_rms[k].with_clamp else (1.0, -1.0)
                fields.append((self.buf_obs[k], obs[k], w, ROW_F32_NORMALIZE, m, d, lo, hi, ld(obs[k])))
                fields.append((self.buf_next_obs[k], next_obs[k], w, ROW_F32_NORMALIZE, m, d, lo, hi, ld(next_obs[k])))
            else:
                fields.append((self.buf_obs[k], obs[k], w, kind, None, None, None, None, ld(obs[k])))
                fields.append((self.buf_next_obs[k], next_obs[k], w, kind, None, None, None, None, ld(next_obs[k])))
        fields.append((self.buf_action, action, self.action_dim, ROW_F32, None, None, None, None, ld(action)))
        fields.append((self.buf_reward, reward, 1, ROW_F32))
        fields.append((self.buf_done, done, 1, ROW_BOOL_TO_F32))
        ops.gather_rows(fields, idx=indices.to(self.buf_action.device))
        return obs, action, reward, next_obs, done
