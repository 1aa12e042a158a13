"""CTA-0 timeline of tc_linear_kernel (debug aid, csrc/tc_linear.cu TC_TRACE): per k-block the TMA issue, converter
start / end, MMA start / issue-end; per tile the epilogue start / end.  usage: trace_tc.py [fwd|dx] M N K"""
import ctypes
import sys

import numpy as np
import torch

sys.path.insert(0, "/root/repo")
from mineral_b200 import ops  # noqa: E402
from mineral_b200._native import lib  # noqa: E402

DEV = "cuda:0"
mode = sys.argv[1] if len(sys.argv) > 1 else "fwd"
M, N, K = (int(v) for v in sys.argv[2:5]) if len(sys.argv) > 4 else (32768, 128, 256)
g = torch.Generator(device=DEV).manual_seed(0)
ld = (K + 3) // 4 * 4
groups = []
for t in range(2):
    w = torch.randn(N, K, device=DEV, generator=g) / 16
    hi, lo = torch.zeros(N, ld, device=DEV), torch.zeros(N, ld, device=DEV)
    ops.split_tf32([(w, hi, lo, N, K, K, ld)])
    if mode == "fwd":
        x = torch.zeros(M, ld, device=DEV)
        x[:, :K] = torch.randn(M, K, device=DEV, generator=g)
        groups.append(dict(a=x, b=hi, b_lo=lo, c=torch.empty(M, N, device=DEV), bias=torch.randn(N, device=DEV, generator=g), N=N, K=K, lda=ld, ldb=ld))
    elif mode == "dx":
        dy, y = torch.randn(M, N, device=DEV, generator=g), torch.randn(M, N, device=DEV, generator=g)
        groups.append(dict(a=dy, saved=y, b=hi, b_lo=lo, c=torch.empty(M, ld, device=DEV), N=N, K=ld, ldb=ld))
    else:   # dw: dW (N,K) = (dY * act'(Y))^T X
        dy, y = torch.randn(M, N, device=DEV, generator=g), torch.randn(M, N, device=DEV, generator=g)
        x = torch.zeros(M, ld, device=DEV)
        x[:, :K] = torch.randn(M, K, device=DEV, generator=g)
        groups.append(dict(a=dy, saved=y, b=x, c=torch.empty(N, K, device=DEV), bias_grad=torch.empty(N, device=DEV), N=N, K=K, ldb=ld))
if mode == "dw":
    ws = torch.zeros(ops.lib.mnr_tc_linear_bwd_weight_workspace_bytes(2, M, N, K), dtype=torch.uint8, device=DEV)
run = {"fwd": lambda: ops.tc_linear_fwd(groups, M, 1), "dx": lambda: ops.tc_linear_bwd_input(groups, M, 1),
       "dw": lambda: ops.tc_linear_bwd_weight_partials(groups, M, 1, ws)}[mode]
for _ in range(3):
    run()
torch.cuda.synchronize()
lib.mnr_tc_trace(1, None)
run()
torch.cuda.synchronize()
buf = (ctypes.c_longlong * (4 * 256 * 2))()
lib.mnr_tc_trace(0, buf)
tr = np.array(buf[:]).reshape(4, 256, 2)
t0 = tr[0, 0, 0]
rel = lambda x: (x - t0) / 1.9  # ~ns at 1.9 GHz
flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
ts = []
for _ in range(8):
    flush.sum()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); run(); b.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b) * 1e3)
print(f"{mode} M={M} N={N} K={K}: {sorted(ts)[len(ts) // 2]:.1f} us per launch (CUDA events, L2 flushed)")
print("kb : prod_issue | conv_start conv_done | mma_start mma_issued   (ns from first TMA issue)")
for kb in range(40):
    if tr[2, kb, 1] <= 0:
        break
    print(f"{kb:3d}: {rel(tr[0,kb,0]):8.0f} | {rel(tr[1,kb,0]):8.0f} {rel(tr[1,kb,1]):8.0f} | {rel(tr[2,kb,0]):8.0f} {rel(tr[2,kb,1]):8.0f}")
print("tile: epi_start epi_done")
for it in range(16):
    if tr[3, it, 1] <= 0:
        break
    print(it, f"{rel(tr[3,it,0]):8.0f} {rel(tr[3,it,1]):8.0f}")
