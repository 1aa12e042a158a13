"""On-policy rollout store on the C-ABI kernels.  Mirrors mineral/agents/ppo/experience.py:17-121:
same constructor, same `storage_dict` tensors (time-major (T,N,..), u8 dones), same public methods
and the same env-major `data_dict` (row = n*T + t, contiguous minibatch slices, no shuffle).

What differs underneath:
  * `compute_return` is one scan kernel instead of a T-iteration Python loop (bit-exact);
  * `prepare_training` is one fused pack kernel + one normalise kernel instead of 10 transposed
    copies and 2 reductions, and it writes into persistent env-major buffers (stable addresses, so
    the update can be captured in a CUDA graph; the reference re-allocates every epoch);
  * `compute_return_and_prepare` (new) fuses both and PPO's value-normalisation block
    (mineral/agents/ppo/ppo.py:172-183) into a single pass over the rollout.
"""

import math
import re

import torch

from ... import ops


def transform_op(arr):
    """(T,N,..) -> (N*T,..) of one arbitrary tensor with torch view ops (experience.py:7-14).  Kept for
    API parity only; the accelerated path packs every field in one kernel (prepare_training)."""
    if arr is None:
        return arr
    s = arr.size()
    return arr.transpose(0, 1).reshape(s[0] * s[1], *s[2:])


class ExperienceBuffer:
    def __init__(self, num_envs, horizon_len, batch_size, minibatch_size, obs_space, act_dim, device, cpu_obs_keys=None):
        self.device = device
        self.num_envs = num_envs
        self.transitions_per_env = horizon_len
        self.cpu_obs_keys = cpu_obs_keys
        self.data_dict = None
        self.obs_space = obs_space
        self.act_dim = act_dim
        T, N, A = horizon_len, num_envs, act_dim
        f32 = dict(dtype=torch.float32, device=device)
        self.storage_dict = {
            'rewards': torch.zeros((T, N, 1), **f32),
            'values': torch.zeros((T, N, 1), **f32),
            'neglogp': torch.zeros((T, N), **f32),
            'dones': torch.zeros((T, N), dtype=torch.uint8, device=device),
            'actions': torch.zeros((T, N, A), **f32),
            'mu': torch.zeros((T, N, A), **f32),
            'sigma': torch.zeros((T, N, A), **f32),
            'returns': torch.zeros((T, N, 1), **f32),
        }
        obs_dict = {}
        for k, v in self.obs_space.items():
            if cpu_obs_keys is not None and re.match(cpu_obs_keys, k):
                raise NotImplementedError(f"obs key {k!r} matches cpu_obs_keys: host-resident observations are not on the B200 path")
            obs_dict[k] = torch.zeros((T, N, *v), **f32)
        self.storage_dict['obses'] = obs_dict

        self._stage = None   # double-buffered host upload (stage_from_host / commit_staged)

        self.batch_size = batch_size
        self.minibatch_size = minibatch_size
        self.length = self.batch_size // self.minibatch_size
        self.last_range = None

        # persistent env-major training buffers + scratch
        B = T * N
        self._em = {
            'rewards': torch.zeros((B, 1), **f32),
            'values': torch.zeros((B, 1), **f32),
            'neglogp': torch.zeros((B,), **f32),
            'dones': torch.zeros((B,), dtype=torch.uint8, device=device),
            'actions': torch.zeros((B, A), **f32),
            'mu': torch.zeros((B, A), **f32),
            'sigma': torch.zeros((B, A), **f32),
            'returns': torch.zeros((B, 1), **f32),
            'advantages': torch.zeros((B,), **f32),
            'obses': {k: torch.zeros((B, *v), **f32) for k, v in self.obs_space.items()},
        }
        self._ws = ops.post_rollout_workspace(T, N, device)
        self._ws_all = None  # combined partials of all ranks (multi-GPU), see mineral_b200/distributed.py
        self.stats = torch.zeros(10, dtype=torch.float64, device=device)  # see mnr_post_rollout_finalize

    def __len__(self):
        return self.length

    def __getitem__(self, idx):
        """Minibatch idx = rows [idx*mb, (idx+1)*mb) as zero-copy views (experience.py:61-81)."""
        start = idx * self.minibatch_size
        end = (idx + 1) * self.minibatch_size
        self.last_range = (start, end)
        d = self.data_dict
        return (
            d['values'][start:end],
            d['neglogp'][start:end],
            d['advantages'][start:end],
            d['mu'][start:end],
            d['sigma'][start:end],
            d['returns'][start:end],
            d['actions'][start:end],
            {k: v[start:end] for k, v in d['obses'].items()},
        )

    def update_mu_sigma(self, mu, sigma):
        """experience.py:83-87, as one two-field row-scatter launch."""
        start, end = self.last_range
        A = self.act_dim
        ops.scatter_rows([(mu.contiguous(), self.data_dict['mu'], A, 0), (sigma.contiguous(), self.data_dict['sigma'], A, 0)],
                         idx=None, row0=start, n=end - start)

    # ---- host-resident rollouts (a simulator on the CPU, a remote actor): double-buffered upload ---------------------------
    def stage_from_host(self, host):
        """Start uploading a rollout from pinned host tensors ({name: (T,N,..)}; observations under 'obs' or 'obses': {key: ..})
        into device STAGING buffers on a dedicated copy stream and return at once: the copy engines run while the compute
        stream is still busy with the previous update.  `commit_staged()` makes the staged rollout the current one."""
        if self._stage is None:
            self._stage = dict(bufs={}, stream=torch.cuda.Stream(device=self.device), ready=torch.cuda.Event(), free=torch.cuda.Event())
            self._stage['free'].record(torch.cuda.current_stream())
        st = self._stage
        st['stream'].wait_event(st['free'])   # the previous commit has finished reading the staging buffers
        with torch.cuda.stream(st['stream']):
            for name, src in self._host_fields(host):
                dst = st['bufs'].get(name)
                if dst is None:
                    dst = st['bufs'][name] = torch.empty_like(self._storage_field(name))
                dst.copy_(src, non_blocking=True)
            st['ready'].record(st['stream'])
        st['names'] = [name for name, _ in self._host_fields(host)]

    def commit_staged(self):
        """storage_dict <- staged rollout (device-to-device, on the current stream, after the upload has landed).  The storage
        tensors keep their addresses, so a captured update graph stays valid."""
        st = self._stage
        cur = torch.cuda.current_stream()
        cur.wait_event(st['ready'])
        for name in st['names']:
            self._storage_field(name).copy_(st['bufs'][name], non_blocking=True)
        st['free'].record(cur)

    @staticmethod
    def _host_fields(host):
        for k, v in host.items():
            if k == 'obses':
                for kk, vv in v.items():
                    yield ('obses', kk), vv
            elif k == 'obs':
                yield ('obses', 'obs'), v
            else:
                yield k, v

    def _storage_field(self, name):
        return self.storage_dict['obses'][name[1]] if isinstance(name, tuple) else self.storage_dict[name]

    def update_data(self, name, index, val):
        """Rollout insert storage[name][t, :] = val (experience.py:89-94).  A contiguous device copy:
        memory plumbing, left to torch's copy engine path."""
        if type(val) is dict:
            for k, v in val.items():
                self.storage_dict[name][k][index, :] = v
        else:
            self.storage_dict[name][index, :] = val

    # ---- update path -----------------------------------------------------------------------
    def compute_return(self, last_values, gamma, tau):
        """GAE into storage_dict['returns'] (experience.py:96-108), one kernel, bit-exact."""
        sd = self.storage_dict
        ops.gae(sd['rewards'], sd['values'], sd['dones'], last_values.contiguous(), gamma, tau, returns=sd['returns'])

    def _wide(self):
        sd, em = self.storage_dict, self._em
        wide = [(sd[k], em[k]) for k in ('actions', 'mu', 'sigma')]
        wide += [(sd['obses'][k], em['obses'][k]) for k in self.obs_space]
        return wide

    def _publish(self):
        em = self._em
        self.data_dict = {k: (dict(v) if isinstance(v, dict) else v) for k, v in em.items()}
        return self.data_dict

    def prepare_training(self):
        """Env-major pack + advantage normalisation from the stored returns (experience.py:110-121)."""
        sd, T, N = self.storage_dict, self.transitions_per_env, self.num_envs
        ops.post_rollout_fused(self._wide(), sd['rewards'], sd['values'], sd['dones'], sd['neglogp'], None, sd['returns'],
                               self._em, T, N, 0.0, 0.0, self._ws, do_gae=False)
        ops.post_rollout_finalize(self._em, T * N, None, self._ws, stats_out=self.stats)
        return self._publish()

    def _launch_fused_only(self, last_values, gamma, tau):
        """Just the fused GAE + pack + partial-moments kernel (bench.py times it in isolation)."""
        sd, T, N = self.storage_dict, self.transitions_per_env, self.num_envs
        ops.post_rollout_fused(self._wide(), sd['rewards'], sd['values'], sd['dones'], sd['neglogp'], last_values.contiguous(),
                               sd['returns'], self._em, T, N, gamma, tau, self._ws, do_gae=True)

    def compute_return_and_prepare(self, last_values, gamma, tau, value_rms=None, world=1, comm=None):
        """Fused compute_return + prepare_training (+ PPO's value-RMS block when `value_rms`, a
        RunningMeanStd over shape (1,), is given; ppo.py:175-183): one pass over the rollout, then one
        small normalise pass.  Equivalent to calling the two reference methods and the value block.
        `world` > 1: the statistics cover every rank's shard (one all-gather of the partial moments)."""
        sd, T, N = self.storage_dict, self.transitions_per_env, self.num_envs
        ops.post_rollout_fused(self._wide(), sd['rewards'], sd['values'], sd['dones'], sd['neglogp'], last_values.contiguous(),
                               sd['returns'], self._em, T, N, gamma, tau, self._ws, do_gae=True)
        ws = self._ws
        if world > 1:  # advantage / value statistics over ALL ranks' samples
            from ... import distributed as mdist
            if self._ws_all is None:
                self._ws_all = mdist.combined_workspace(N, world, self.device)
            ws = mdist.gather_partials(self._ws, self._ws_all, N, comm=comm)
        if value_rms is None:
            ops.post_rollout_finalize(self._em, T * N, None, ws, stats_out=self.stats)
        else:
            clamp = tuple(value_rms.clamp_range) if value_rms.with_clamp else None
            ops.post_rollout_finalize(self._em, T * N, (value_rms.running_mean, value_rms.running_var, value_rms.running_count),
                                      ws, eps=value_rms.eps, clamp=clamp, stats_out=self.stats)
        return self._publish()
