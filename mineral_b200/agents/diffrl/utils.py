"""SHAC/SAPO critic-side helpers.  Mirrors mineral/agents/diffrl/utils.py:5-69 for what sits on the
rollout->update path: `CriticDataset` (time-major flatten, optional numpy permutation, contiguous
batch slices, ragged last batch) and the small host/elementwise utilities the critic loop uses."""

import numpy as np
import torch

from ... import ops


def adaptive_scheduler(current_lr, kl_dist, kl_threshold=0.008, min_lr=1e-6, max_lr=1e-2, decay=1.5):
    lr = current_lr
    if kl_dist > 2.0 * kl_threshold:
        lr = max(current_lr / decay, min_lr)
    if kl_dist < 0.5 * kl_threshold:
        lr = min(current_lr * decay, max_lr)
    return lr


def grad_norm(params):
    """sqrt(sum ||g||^2) (utils.py:31-36) via the grad-norm kernel when the grads are one flat buffer,
    else accumulated per tensor."""
    total = None
    for p in params:
        if p.grad is not None:
            s = torch.sum(p.grad**2)
            total = s if total is None else total + s
    return torch.sqrt(total)


def soft_update(module, module_target, alpha: float):
    for param, param_targ in zip(module.parameters(), module_target.parameters()):
        param_targ.data.mul_(alpha)
        param_targ.data.add_((1.0 - alpha) * param.data)


class CriticDataset:
    def __init__(self, batch_size, obs, target_values, shuffle=False, drop_last=False):
        # (T, B, ...) -> (T * B, ...): plain views, row = t * B + b (utils.py:41-43)
        self.obs = {k: v.view(-1, *v.shape[2:]) for k, v in obs.items()}
        self.target_values = target_values.view(-1)
        self.N = self.target_values.shape[0]
        self.batch_size = batch_size
        if shuffle:
            self.shuffle()
        self.length = self.N // self.batch_size if drop_last else ((self.N - 1) // self.batch_size) + 1

    def shuffle(self):
        """`np.random.permutation` on the host (RNG parity by construction, utils.py:55-58), then ONE
        row-gather launch for all fields."""
        index = torch.from_numpy(np.random.permutation(self.N)).to(self.target_values.device)
        fields, new_obs = [], {}
        for k, v in self.obs.items():
            new_obs[k] = torch.empty_like(v)
            fields.append((v.contiguous(), new_obs[k], v[0].numel(), 0))
        new_t = torch.empty_like(self.target_values)
        fields.append((self.target_values.contiguous(), new_t, 1, 0))
        ops.gather_rows(fields, idx=index)
        self.obs, self.target_values = new_obs, new_t

    def __len__(self):
        return self.length

    def __getitem__(self, index):
        s = index * self.batch_size
        e = min((index + 1) * self.batch_size, self.N)
        return {k: v[s:e, ...] for k, v in self.obs.items()}, self.target_values[s:e]
