"""LayerNorm + SiLU forward / backward alone (two towers, L2 flushed, CUDA events): time and achieved HBM GB/s on the
algorithmic bytes (8 / 12 B per element).  usage: ln_probe.py [rows]; MINERAL_B200_LIB=<variant .so> selects an A/B build."""
import sys

import torch

sys.path.insert(0, "/root/repo")
from mineral_b200 import ops  # noqa: E402

DEV = "cuda:0"
M = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)


def timed(fn, n=12):
    ts = []
    for _ in range(n):
        flush.sum()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


for N in (512, 256, 128):
    grp = []
    for _ in range(2):
        z, y = torch.randn(M, N, device=DEV), torch.empty(M, N, device=DEV)
        grp.append(dict(z=z, y=y, gamma=torch.rand(N, device=DEV) + 0.5, beta=torch.randn(N, device=DEV), mean=torch.empty(M, device=DEV),
                        rstd=torch.empty(M, device=DEV), dy=torch.randn(M, N, device=DEV), dz=torch.empty(M, N, device=DEV),
                        dgamma=torch.empty(N, device=DEV), dbeta=torch.empty(N, device=DEV)))
    ws = ops.ln_act_bwd_workspace(2, M, N, DEV)
    tf = timed(lambda: ops.ln_act_fwd(grp, M, N, 1e-5, 3))
    tb = timed(lambda: ops.ln_act_bwd(grp, M, N, 3, ws))
    print(f"rows={M} N={N} x2 towers: fwd {tf:.1f} us ({2 * 8 * M * N / tf / 1e3:.0f} GB/s)  bwd {tb:.1f} us ({2 * 12 * M * N / tb / 1e3:.0f} GB/s)")
