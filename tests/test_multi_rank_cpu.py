"""world_size-2 `gloo` tests (CPU) of the host-side multi-GPU logic in mineral_b200/distributed.py:
the statistics merges that make an N-rank env-sharded update equal the 1-rank update on the
concatenated batch.  The kernels themselves are covered by the -m gpu suite."""

import os
import struct

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import scale_close


def _worker(rank, world, port, fn, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ret[rank] = fn(rank, world)
    finally:
        dist.destroy_process_group()


def _spawn(fn, world=2, port=29561):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, fn, ret), nprocs=world, join=True)
    return [ret[r] for r in range(world)]


def _moments_job(rank, world):
    from mineral_b200 import distributed as mdist

    g = torch.Generator().manual_seed(7)
    full = torch.randn(world * 4 * 64, 9, generator=g) * torch.linspace(0.2, 6, 9) + 1.0   # 4 minibatches of 64 rows per rank
    shard = full.view(world, 4, 64, 9)[rank]
    mean, var = shard.mean(1), shard.var(1, correction=1)
    mdist.merge_group_moments(mean, var, 64, 1)
    return mean, var


def test_group_moments_merge_equals_union():
    g = torch.Generator().manual_seed(7)
    full = torch.randn(2 * 4 * 64, 9, generator=g) * torch.linspace(0.2, 6, 9) + 1.0
    union = full.view(2, 4, 64, 9).transpose(0, 1).reshape(4, 128, 9)     # minibatch i = both ranks' slice i
    out = _spawn(_moments_job)
    for mean, var in out:
        scale_close(mean, union.mean(1), 1e-6, "merged mean")
        scale_close(var, union.var(1, correction=1), 1e-6, "merged var")
    assert torch.equal(out[0][0], out[1][0]) and torch.equal(out[0][1], out[1][1])   # identical on every rank


def _partials_job(rank, world):
    from mineral_b200 import distributed as mdist

    N = 70  # 3 partials per rank (32 + 32 + 6 envs)
    n = mdist.n_partials(N)
    local = torch.zeros(mdist.POST_ROLLOUT_HEADER + (n + 1) * mdist.POST_ROLLOUT_PARTIAL, dtype=torch.uint8)
    payload = bytes((rank * 100 + i) % 251 for i in range(n * mdist.POST_ROLLOUT_PARTIAL))
    local[mdist.POST_ROLLOUT_HEADER:mdist.POST_ROLLOUT_HEADER + len(payload)] = torch.tensor(list(payload), dtype=torch.uint8)
    ws_all = mdist.combined_workspace(N, world, "cpu")
    mdist.gather_partials(local, ws_all, N)
    return ws_all


def test_partials_gather_layout():
    from mineral_b200 import distributed as mdist

    out = _spawn(_partials_job, port=29562)
    assert torch.equal(out[0], out[1])
    ws = bytes(out[0].tolist())
    assert struct.unpack("<I", ws[:4])[0] == 2 * 3
    body = ws[mdist.POST_ROLLOUT_HEADER:]
    n = 3 * mdist.POST_ROLLOUT_PARTIAL
    for rank in range(2):
        assert body[rank * n:(rank + 1) * n] == bytes((rank * 100 + i) % 251 for i in range(n))


def test_struct_sizes_match_kernel_constants():
    """The byte constants mirrored in distributed.py must match csrc/post_rollout.cu."""
    import re

    from conftest import ROOT
    from mineral_b200 import distributed as mdist

    src = open(os.path.join(ROOT, "mineral_b200", "csrc", "post_rollout.cu")).read()
    assert int(re.search(r"kHeaderBytes = (\d+)", src).group(1)) == mdist.POST_ROLLOUT_HEADER
    assert int(re.search(r"kEnvPerScalarCta = (\d+)", src).group(1)) == mdist.ENVS_PER_PARTIAL
    assert mdist.POST_ROLLOUT_PARTIAL == 9 * 8
