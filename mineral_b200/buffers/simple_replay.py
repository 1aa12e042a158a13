"""Circular replay buffer on the C-ABI row kernels.  Mirrors mineral/buffers/simple_replay.py:4-85
(same ctor, attributes `next_p / if_full / cur_capacity / capacity / buf_*`, same wrap rule, indices
drawn with `torch.randint` so sampling is bit-identical given the same generator state).

Layout (round 2): ONE ring of 128 B-aligned transition records `(C, R)` f32 -
`[obs (pad4) | next_obs (pad4) | action (pad4) | reward | done | pad]`, Ant: 96 floats = three cache lines - and the
reference's five arrays `buf_obs / buf_action / buf_next_obs / buf_reward / buf_done` are strided VIEWS of it with the
reference's shapes and dtypes (`(C, *shape)` f32, `(C, 1)` bool), so indexing them works as before while a uniform sample
reads whole lines (csrc/replay_records.cu: one TMA bulk copy per sampled row, staged in shared memory)."""

import math

import torch

from .. import ops
from .._native import ROW_BOOL_TO_F32, ROW_F32, ROW_F32_NORMALIZE, ROW_F32_TO_BOOL, ROW_U8


def _width(shape):
    return int(math.prod(shape)) if len(shape) else 1


def _pad4(n):
    return (n + 3) // 4 * 4


def record_layout(obs_space, action_dim):
    """-> (offsets {('obs', k) | ('next_obs', k) | 'action' | 'reward' | 'done': first float}, floats per record)."""
    off, o = {}, 0
    for part in ('obs', 'next_obs'):
        for k, v in obs_space.items():
            off[(part, k)] = o
            o += _pad4(_width(v))
        if part == 'obs':
            pass
    off['action'] = o
    o += _pad4(int(action_dim))
    off['reward'] = o
    o += 4
    off['done'] = o
    o += 4
    return off, (o + 31) // 32 * 32


def create_buffer(capacity, obs_space, action_dim, device='cuda', return_records=False):
    """simple_replay.py:4-16: f32 obs / action / reward / next_obs, BOOL done - as views of one record ring."""
    if not isinstance(capacity, int):
        (capacity,) = capacity
    off, R = record_layout(obs_space, action_dim)
    rec = torch.zeros((capacity, R), dtype=torch.float32, device=device)
    buf_obs = {k: rec[:, off[('obs', k)]:off[('obs', k)] + _width(v)].unflatten(1, tuple(v)) for k, v in obs_space.items()}
    buf_next_obs = {k: rec[:, off[('next_obs', k)]:off[('next_obs', k)] + _width(v)].unflatten(1, tuple(v)) for k, v in obs_space.items()}
    buf_action = rec[:, off['action']:off['action'] + int(action_dim)]
    buf_reward = rec[:, off['reward']:off['reward'] + 1]
    d = off['done'] * 4
    buf_done = rec.view(torch.uint8)[:, d:d + 1].view(torch.bool)   # byte 0 of the done slot
    out = (buf_obs, buf_action, buf_next_obs, buf_reward, buf_done)
    return out + (rec, off) if return_records else out


class ReplayBuffer:
    def __init__(self, obs_space, action_dim: int, capacity: int, device='cuda'):
        self.obs_space = obs_space
        self.action_dim = action_dim
        self.device = device
        self.next_p = 0
        self.if_full = False
        self.cur_capacity = 0
        self.capacity = int(capacity)
        ret = create_buffer(self.capacity, obs_space, action_dim, device=device, return_records=True)
        self.buf_obs, self.buf_action, self.buf_next_obs, self.buf_reward, self.buf_done, self.records, self._off = ret

    @torch.no_grad()
    def add_to_buffer(self, trajectory):
        """Ring insert, one launch for all fields (simple_replay.py:32-71)."""
        obs, actions, rewards, next_obs, dones = trajectory
        n = rewards.numel()
        rec, off = self.records, self._off
        fields, offsets = [], []
        for k in self.buf_obs:
            fields.append((obs[k].contiguous(), rec, _width(self.obs_space[k]), ROW_F32)), offsets.append(off[('obs', k)])
        fields.append((actions.contiguous(), rec, self.action_dim, ROW_F32)), offsets.append(off['action'])
        fields.append((rewards.contiguous(), rec, 1, ROW_F32)), offsets.append(off['reward'])
        for k in self.buf_next_obs:
            fields.append((next_obs[k].contiguous(), rec, _width(self.obs_space[k]), ROW_F32)), offsets.append(off[('next_obs', k)])
        dones = dones.contiguous()
        if dones.dtype == torch.float32:
            fields.append((dones, rec, 1, ROW_F32_TO_BOOL))  # dones.bool(), simple_replay.py:39
        elif dones.dtype in (torch.bool, torch.uint8):
            fields.append((dones, rec, 1, ROW_U8))
        else:
            raise TypeError(f'dones dtype {dones.dtype}')
        offsets.append(off['done'])
        ops.replay_insert_records(fields, offsets, rec, n, self.next_p, self.capacity)

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
        rec, off = self.records, self._off
        fields, offsets = [], []
        for k in self.buf_obs:
            w = _width(self.obs_space[k])
            if obs_rms is not None and hasattr(obs_rms[k], 'running_mean'):
                m, d = obs_rms[k]._prepare()
                lo, hi = obs_rms[k].clamp_range if obs_rms[k].with_clamp else (1.0, -1.0)
                fields.append((rec, obs[k], w, ROW_F32_NORMALIZE, m, d, lo, hi, ld(obs[k])))
                fields.append((rec, next_obs[k], w, ROW_F32_NORMALIZE, m, d, lo, hi, ld(next_obs[k])))
            else:
                fields.append((rec, obs[k], w, ROW_F32, None, None, None, None, ld(obs[k])))
                fields.append((rec, next_obs[k], w, ROW_F32, None, None, None, None, ld(next_obs[k])))
            offsets += [off[('obs', k)], off[('next_obs', k)]]
        fields.append((rec, action, self.action_dim, ROW_F32, None, None, None, None, ld(action))), offsets.append(off['action'])
        fields.append((rec, reward, 1, ROW_F32)), offsets.append(off['reward'])
        fields.append((rec, done, 1, ROW_BOOL_TO_F32)), offsets.append(off['done'])
        ops.replay_sample_records(fields, offsets, rec, indices.to(rec.device))
        return obs, action, reward, next_obs, done
