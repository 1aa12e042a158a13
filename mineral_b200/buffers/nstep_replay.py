"""N-step return window on the C-ABI kernel.  Mirrors mineral/buffers/nstep_replay.py:10-94: same
ctor, `add_to_buffer(obs, actions, rewards, next_obs, dones) -> 5-tuple` with step-major rows.

The reference re-concatenates five (N, nstep, ..) FIFO tensors on every env step and calls a
TorchScript function per step; here a call with S steps is ONE launch that reads each transition's
window straight from the inputs (or the carried-over window) and also writes the window that
survives the call.  `dones` is a float tensor in the reference's own use (sac.py:150; its argmax rejects
bool) and `done` then comes back as float 0/1; bool/u8 dones are accepted too and come back as bool.
While the window is still warming up (fewer than nstep steps seen in total) the
reference crashes in torch.cat of an empty list; this class returns empty tensors instead."""

import math

import torch

from .. import ops
from .simple_replay import create_buffer


class NStepReplay:
    def __init__(self, obs_space, action_dim: int, num_envs: int = 1, nstep: int = 3, gamma: float = 0.99, device: str = 'cuda'):
        self.obs_space = obs_space
        self.action_dim = action_dim
        self.num_envs = num_envs
        self.nstep = nstep
        self.gamma = gamma
        self.device = device
        self.nstep_count = 0
        if len(obs_space) != 1:
            raise NotImplementedError('NStepReplay on the B200 path supports a single flat observation key')
        self._key = next(iter(obs_space))
        self._O = int(math.prod(obs_space[self._key]))
        # gamma ** i as Python doubles, rounded to f32 (nstep_replay.py:27)
        self.gamma_array = torch.tensor([self.gamma**i for i in range(self.nstep)]).to(device).view(-1, 1)
        self._win = [self._alloc_window(), self._alloc_window()]
        self._cur = 0

    def _alloc_window(self):
        N, n, dev = self.num_envs, self.nstep, self.device
        f32 = dict(dtype=torch.float32, device=dev)
        return dict(obs=torch.zeros(N, n, self._O, **f32), act=torch.zeros(N, n, self.action_dim, **f32),
                    rew=torch.zeros(N, n, 1, **f32), nobs=torch.zeros(N, n, self._O, **f32),
                    done=torch.zeros(N, n, 1, dtype=torch.uint8, device=dev))

    @torch.no_grad()
    def add_to_buffer(self, obs, actions, rewards, next_obs, dones):
        if self.nstep <= 1:
            return obs, actions, rewards, next_obs, dones
        k, N, n, O, A = self._key, self.num_envs, self.nstep, self._O, self.action_dim
        S = dones.shape[1]
        if dones.dtype not in (torch.float32, torch.bool, torch.uint8):
            raise TypeError(f'dones dtype {dones.dtype}')
        inp = dict(obs=obs[k].contiguous(), act=actions.contiguous(), rew=rewards.contiguous(), nobs=next_obs[k].contiguous(),
                   done=dones.contiguous())
        skip = min(max(0, n - 1 - self.nstep_count), S)
        rows = (S - skip) * N
        f32 = dict(dtype=torch.float32, device=self.device)
        out = dict(obs=torch.empty(rows, *self.obs_space[k], **f32), act=torch.empty(rows, A, **f32), rew=torch.empty(rows, 1, **f32),
                   nobs=torch.empty(rows, *self.obs_space[k], **f32),
                   done=torch.empty(rows, 1, dtype=torch.float32 if dones.dtype == torch.float32 else torch.bool, device=self.device))
        win_in, win_out = self._win[self._cur], self._win[1 - self._cur]
        ops.nstep_push(inp, win_in, win_out, out if rows > 0 else None, self.gamma_array.view(-1), N, S, O, A, n, self.nstep_count)
        self._cur = 1 - self._cur
        self.nstep_count += S
        return {k: out['obs']}, out['act'], out['rew'], {k: out['nobs']}, out['done']
