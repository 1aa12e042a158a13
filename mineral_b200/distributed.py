"""Multi-GPU plumbing for the env-sharded update (one process per GPU, torch.distributed over NCCL;
`gloo` in CPU tests).  The path shards by environment: rank g owns envs [g*N, (g+1)*N) and its own
rollout storage, so the only exchanges are sums over samples:

  * per optimizer step: ONE all-reduce of the flat gradient buffer (pre-scaled by 1/world inside the
    loss kernel) and one of the 10 loss metrics -- issued from agents/ppo/ppo.py;
  * per rollout: the statistics merges below, so that an N-rank run on env shards equals a 1-rank run
    on the concatenated batch with minibatch i = union of the ranks' local minibatch i.

The reference has no such merge (its accelerate/DDP path only averages gradients and gathers metrics,
mineral/agents/ppo/ppo.py:116-123,292-311); `RunningMeanStd._update_from_moments`
(mineral/common/normalizers.py:44-55) applied across ranks is the definition used here.
Everything in this file is device-agnostic torch / torch.distributed host logic on tiny tensors.
"""

import torch
import torch.distributed as dist

POST_ROLLOUT_HEADER = 128   # bytes, csrc/post_rollout.cu kHeaderBytes
POST_ROLLOUT_PARTIAL = 72   # bytes, sizeof(TilePartial) = 3 x (count, mean, M2) doubles
ENVS_PER_PARTIAL = 32       # csrc/post_rollout.cu kEnvPerScalarCta


def world_size(group=None):
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


def merge_group_moments(mean, var, count, correction=1, group=None, comm=None):
    """In-place: per-rank (mean, var) of `count` rows each -> moments of the union of all ranks' rows.
    mean/var: (n_groups, cols) f32.  Exact sufficient statistics (sum x, sum x^2) in f64, one exchange:
    a peer-memory all-gather + rank-order sum when `comm` (mineral_b200.peer.PeerComm) is given, else one
    all-reduce on `group`."""
    w = comm.world if comm is not None else world_size(group)
    if w == 1:
        return mean, var
    cnt, corr = float(count), float(correction)
    m = mean.double()
    pack = torch.stack([m * cnt, var.double() * (cnt - corr) + cnt * m * m])
    if comm is not None:
        pack = comm.allgather(pack.contiguous()).sum(0)
    else:
        dist.all_reduce(pack, group=group)
    tot = cnt * w
    gmean = pack[0] / tot
    mean.copy_(gmean.float())
    var.copy_(((pack[1] - tot * gmean * gmean) / (tot - corr)).float())
    return mean, var


def n_partials(num_envs):
    return (num_envs + ENVS_PER_PARTIAL - 1) // ENVS_PER_PARTIAL


def combined_workspace(num_envs, world, device):
    """Workspace that holds every rank's post-rollout partial moments behind one header."""
    n = n_partials(num_envs)
    ws = torch.zeros(POST_ROLLOUT_HEADER + world * n * POST_ROLLOUT_PARTIAL, dtype=torch.uint8, device=device)
    ws[:4].view(torch.int32).fill_(world * n)
    return ws


def gather_partials(ws_local, ws_all, num_envs, group=None, comm=None):
    """All-gather the local (adv, values, returns) partial moments into `ws_all` (rank-major), which
    mnr_post_rollout_finalize then merges in index order: identical statistics on every rank."""
    n = n_partials(num_envs) * POST_ROLLOUT_PARTIAL
    if comm is not None and n % 16 == 0:
        comm.allgather(ws_local[POST_ROLLOUT_HEADER:POST_ROLLOUT_HEADER + n], out=ws_all[POST_ROLLOUT_HEADER:])
        return ws_all
    dist.all_gather_into_tensor(ws_all[POST_ROLLOUT_HEADER:], ws_local[POST_ROLLOUT_HEADER:POST_ROLLOUT_HEADER + n].contiguous(),
                                group=group)
    return ws_all
