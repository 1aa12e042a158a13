"""Peer-memory communicator for the env-sharded update (one process per GPU, NVLink 5 / NVSwitch).

Every rank allocates ONE peer buffer through the C ABI (`mnr_peer_alloc`: cudaMalloc + CUDA IPC handle),
the 64-byte handles travel over torch.distributed (`all_gather_object`, plumbing) and every rank maps
every other rank's buffer (`mnr_peer_open`).  After that the data path needs no NCCL call:

  * `reduce_clip_adam`  - csrc/peer.cu, one kernel per optimizer step: gradient + metrics all-reduce over
                          P2P loads, global-norm clip and Adam/AdamW (replaces DDP's all-reduce,
                          clip_grad_norm_ and optimizer.step of mineral/agents/ppo/ppo.py:291-311);
  * `allgather`         - small statistics payloads (post-rollout moment partials, per-minibatch obs
                          moments) once per rollout (the RunningMeanStd merge the reference lacks).

Both are graph-capturable (fixed pointers, device-resident barrier flags), so a multi-GPU PPO update is
replayed as one CUDA graph exactly like the single-GPU one.
"""

import ctypes as C

import torch
import torch.distributed as dist

from . import _native as nv
from ._native import check, lib, ptr, stream

SIGNAL_BYTES = 16384   # MNR_PEER_SIGNAL_BYTES
HANDLE_BYTES = 64      # MNR_PEER_HANDLE_BYTES
MAX_PEERS = 8          # MNR_MAX_PEERS
TAIL_FLOATS = 32       # metrics slot behind the gradient


class _RawCuda:
    """Exposes a raw device allocation to torch (zero-copy) through __cuda_array_interface__."""

    def __init__(self, addr, nbytes):
        self.__cuda_array_interface__ = dict(shape=(int(nbytes),), typestr="|u1", data=(int(addr), False), version=3)


def _align(x, a=256):
    return (x + a - 1) // a * a


class PeerComm:
    def __init__(self, regions, device, group=None):
        """regions: {name: bytes}.  Collective: every rank of `group` must call it with the same regions."""
        assert dist.is_initialized(), "PeerComm needs torch.distributed (torchrun) for the handle exchange"
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if self.world > MAX_PEERS:
            raise nv.MineralB200Error(f"peer communicator supports up to {MAX_PEERS} GPUs of one NVLink domain, got {self.world}")
        self.device = torch.device(device)
        self.offsets, off = {}, SIGNAL_BYTES
        for name, nbytes in regions.items():
            self.offsets[name] = (off, int(nbytes))
            off = _align(off + int(nbytes))
        self.nbytes = off
        with torch.cuda.device(self.device):
            p, h = C.c_void_p(), (C.c_ubyte * HANDLE_BYTES)()
            check(lib.mnr_peer_alloc(self.nbytes, C.byref(p), h), "mnr_peer_alloc", kernels=0)
            self._local = p.value
            handles = [None] * self.world
            dist.all_gather_object(handles, bytes(h), group=group)
            self._mapped = []
            comm = nv.PeerCommStruct()
            comm.rank, comm.world = self.rank, self.world
            for r in range(self.world):
                if r == self.rank:
                    comm.base[r] = self._local
                    continue
                q = C.c_void_p()
                buf = (C.c_ubyte * HANDLE_BYTES).from_buffer_copy(handles[r])
                check(lib.mnr_peer_open(buf, C.byref(q)), f"mnr_peer_open(rank {r})", kernels=0)
                comm.base[r] = q.value
                self._mapped.append(q.value)
            self.comm = comm
            self._bytes = torch.as_tensor(_RawCuda(self._local, self.nbytes), device=self.device)
        dist.barrier(group=group)  # every rank has mapped every buffer before anyone launches into it

    def tensor(self, name, dtype=torch.float32):
        """Local view of region `name`."""
        off, n = self.offsets[name]
        return self._bytes[off:off + n].view(dtype)

    def offset(self, name):
        return self.offsets[name][0]

    def allgather(self, src, out=None, stage="stage"):
        """src: contiguous local tensor (bytes % 16 == 0) -> (world, *src.shape) with every rank's payload."""
        off, cap = self.offsets[stage]
        nbytes = src.numel() * src.element_size()
        if nbytes % 16 != 0 or nbytes > cap:
            raise nv.MineralB200Error(f"allgather payload of {nbytes} B must be a multiple of 16 and fit the {cap} B staging region")
        if out is None:
            out = torch.empty((self.world,) + tuple(src.shape), dtype=src.dtype, device=src.device)
        check(lib.mnr_peer_allgather(C.byref(self.comm), off, ptr(src), nbytes, ptr(out), stream()), "mnr_peer_allgather")
        return out

    def reduce_clip_adam(self, n, reduced, param, exp_avg, exp_avg_sq, max_norm, beta1, beta2, eps, weight_decay, decoupled,
                         zero_grad, st, tail_out, kl_index, norm_out, ws, region="grad", zero_region=None, end_barrier=True, trace=None):
        """region: where this step's gradients sit; zero_region: the region of the previous step (ping-pong), cleared after
        the start barrier; end_barrier=False skips the second cross-GPU barrier (legal with ping-pong regions, except on the
        last step of an update); trace: optional (512,) int64 device tensor receiving globaltimer stamps of CTA 0."""
        check(lib.mnr_peer_reduce_clip_adam_pp(
            C.byref(self.comm), self.offsets[region][0], self.offsets[zero_region][0] if zero_region else 0, 1 if end_barrier else 0, n,
            TAIL_FLOATS, ptr(reduced, torch.float32), ptr(param, torch.float32), ptr(exp_avg, torch.float32), ptr(exp_avg_sq, torch.float32),
            max_norm, beta1, beta2, eps, weight_decay, 1 if decoupled else 0, 1 if (zero_grad and end_barrier) else 0, ptr(st, torch.uint8),
            ptr(tail_out, torch.float32, allow_none=True), kl_index, ptr(norm_out, torch.float32, allow_none=True),
            ptr(trace, torch.int64, allow_none=True), ptr(ws, torch.uint8), ws.numel(), stream()), "mnr_peer_reduce_clip_adam_pp")

    def error(self):
        """Synchronous host read: non-zero if a cross-GPU wait timed out (a peer died or diverged)."""
        e = C.c_int(0)
        check(lib.mnr_peer_error(C.byref(self.comm), C.byref(e)), "mnr_peer_error", kernels=0)
        return e.value

    def close(self):
        if getattr(self, "_local", None) is None:
            return
        torch.cuda.synchronize(self.device)
        try:
            dist.barrier(group=self.group)  # nobody is still reading my buffer
        except Exception:
            pass
        for q in self._mapped:
            lib.mnr_peer_close(C.c_void_p(q))
        self._bytes = None
        lib.mnr_peer_free(C.c_void_p(self._local))
        self._local, self._mapped = None, []
