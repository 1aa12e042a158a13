"""SAC / DDPG update math on the B200 kernels (csrc/offpolicy.cu), with the reference's function names and
signatures (mineral/agents/ddpg/utils.py:8-31) plus the TD target both agents form inline
(mineral/agents/sac/sac.py:293-297, mineral/agents/ddpg/ddpg.py:292-293).  SURVEY.md section 8(f) row 2: the math that
sits on top of the fused replay sample (`ReplayBuffer.sample_batch`); the agents' actor / critic networks and their
optimizer steps are not mirrored here."""

import ctypes as C

import torch

from ... import _native as nv
from ..._native import check, lib, ptr, stream

F32 = torch.float32


@torch.no_grad()
def soft_update(target_net, current_net, tau: float):
    """Polyak update `tar.mul_(1 - tau).add_(cur * tau)` of every parameter, in place (one launch per parameter tensor;
    use `soft_update_flat` when both networks live in flat buffers)."""
    for tar, cur in zip(target_net.parameters(), current_net.parameters()):
        soft_update_flat(tar.data, cur.data, tau)


@torch.no_grad()
def soft_update_flat(target, current, tau: float):
    assert target.numel() == current.numel()
    check(lib.mnr_soft_update(ptr(target, F32, "target"), ptr(current, F32, "current"), target.numel(), 1.0 - tau, tau, stream()),
          "mnr_soft_update")
    return target


@torch.no_grad()
def td_target(q_values, reward, done, gamma, nstep=1, alpha=None, log_prob=None, out=None):
    """reward + (1 - done) * gamma**nstep * (min_k q_k [- alpha * log_prob]).  q_values: list of (B,1) tensors (the target
    critics' outputs), reward / done (float) / log_prob (B,1)."""
    cs = nv.CriticSet()
    for k, q in enumerate(q_values):
        cs.pred[k] = ptr(q, F32, "q").value
    if out is None:
        out = torch.empty_like(reward)
    a = 0.0 if alpha is None else float(alpha)
    check(lib.mnr_td_target(C.byref(cs), len(q_values), ptr(log_prob, F32, "log_prob", allow_none=True) if alpha is not None else None,
                            ptr(reward, F32, "reward"), ptr(done, F32, "done"), a, gamma ** nstep, ptr(out, F32), reward.numel(), stream()),
          "mnr_td_target")
    return out


@torch.no_grad()
def distl_projection(next_dist, reward, done, gamma, v_min=-10, v_max=10, num_atoms=51, support=None):
    """C51 categorical projection (mineral/agents/ddpg/utils.py:14-31), one thread per batch row."""
    out = torch.empty_like(next_dist)
    delta_z = (v_max - v_min) / (num_atoms - 1)
    check(lib.mnr_distl_projection(ptr(next_dist, F32, "next_dist"), ptr(reward, F32, "reward"), ptr(done, F32, "done"),
                                   ptr(support.to(next_dist.device), F32, "support"), gamma, v_min, v_max, delta_z, num_atoms, ptr(out, F32),
                                   reward.numel(), stream()), "mnr_distl_projection")
    return out
