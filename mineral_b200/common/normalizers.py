"""RunningMeanStd on the C-ABI kernels.  Mirrors mineral/common/normalizers.py:5-81 (same ctor,
methods, buffer names and dtypes, so reference checkpoints load unchanged).

Differences, both deliberate: the running buffers are updated IN PLACE (the reference re-binds new
tensors every update, which defeats CUDA-graph capture), and the batch
moments come from a Welford/Chan reduction instead of torch's cascade sum, so statistics agree with
the reference to rounding (<= 1e-6 relative), not bit for bit.
"""

import math

import torch
import torch.nn as nn

from .. import ops


class Identity(nn.Module):
    def __init__(self, shape=None):
        super().__init__()

    def update(self, x):
        return

    def normalize(self, x):
        return x

    def unnormalize(self, x):
        return x


class RunningMeanStd(nn.Module):
    def __init__(self, shape, eps=1e-4, correction=1, with_clamp=False, clamp_range=(-5.0, 5.0), initial_count='eps',
                 dtype=torch.float32):
        super().__init__()
        if dtype not in (torch.float32, torch.float64):
            raise NotImplementedError(f'RunningMeanStd dtype {dtype}')
        self.shape = shape
        self.eps = eps
        self.correction = correction
        self.with_clamp = with_clamp
        self.clamp_range = clamp_range
        if initial_count == 'eps':
            initial_count = eps
        self.register_buffer('running_mean', torch.zeros(shape, dtype=dtype))
        self.register_buffer('running_var', torch.ones(shape, dtype=dtype))
        self.register_buffer('running_count', torch.tensor(initial_count, dtype=dtype))
        self._cols = int(math.prod(shape)) if len(shape) else 1
        self._scratch = None

    # scratch lives beside the buffers' device; rebuilt after .to(device) / deepcopy onto another device
    def _s(self):
        dev = self.running_mean.device
        s = self._scratch
        if s is None or s['bm'].device != dev:
            f = dict(dtype=torch.float32, device=dev)
            s = dict(bm=torch.empty(self._cols, **f), bv=torch.empty(self._cols, **f), mean=torch.empty(self._cols, **f),
                     denom=torch.empty(self._cols, **f), ws=None, ws_key=None)
            self._scratch = s
        return s

    def _clamp(self):
        return tuple(self.clamp_range) if self.with_clamp else None

    def _rows(self, x):
        if x.dtype != torch.float32:
            raise TypeError(f'RunningMeanStd expects float32 input, got {x.dtype}')
        if not x.is_contiguous():
            x = x.contiguous()
        return x, x.numel() // self._cols

    @torch.no_grad()
    def update(self, x):
        """normalizers.py:57-63: batch mean/var over dim 0, Chan merge into the running state."""
        x, rows = self._rows(x)
        s = self._s()
        key = (rows, self._cols)
        if s['ws_key'] != key:
            s['ws'] = ops.rms_moments_workspace(rows, self._cols, rows, x.device)
            s['ws_key'] = key
        ops.rms_batch_moments(x.view(rows, self._cols), rows, self.correction, s['ws'], s['bm'].view(1, -1), s['bv'].view(1, -1))
        ops.rms_merge(self.running_mean, self.running_var, self.running_count, s['bm'], s['bv'], rows)

    def _prepare(self):
        s = self._s()
        ops.rms_prepare(self.running_mean, self.running_var, self.eps, s['mean'], s['denom'])
        return s['mean'], s['denom']

    @torch.no_grad()
    def normalize(self, x):
        """normalizers.py:65-68."""
        x, _ = self._rows(x)
        m, d = self._prepare()
        return ops.rms_apply(x, m, d, self._clamp(), inverse=False)

    @torch.no_grad()
    def unnormalize(self, x):
        """normalizers.py:70-73 (clamps its INPUT, then scales back)."""
        x, _ = self._rows(x)
        m, d = self._prepare()
        return ops.rms_apply(x, m, d, self._clamp(), inverse=True)

    def clamp(self, x):
        """normalizers.py:75-78; reuses the apply kernel with mean 0 / denom 1."""
        if not self.with_clamp:
            return x
        x, _ = self._rows(x)
        dev = x.device
        zero = torch.zeros(self._cols, dtype=torch.float32, device=dev)
        one = torch.ones(self._cols, dtype=torch.float32, device=dev)
        return ops.rms_apply(x, zero, one, self._clamp(), inverse=False)

    def __repr__(self):
        return f'RunningMeanStd(shape={self.shape})'
