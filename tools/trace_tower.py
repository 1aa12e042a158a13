"""Fused tower forward (csrc/tc_tower.cu): time against three per-layer launches and print the CTA-0 timeline (TT_TRACE).
usage: trace_tower.py [M] [K0] [N1 N2 N3]"""
import ctypes
import sys

import numpy as np
import torch

sys.path.insert(0, "/root/repo")
from mineral_b200 import ops  # noqa: E402
from mineral_b200._native import lib  # noqa: E402

DEV = "cuda:0"
M = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
K0 = int(sys.argv[2]) if len(sys.argv) > 2 else 37
widths = [int(v) for v in sys.argv[3:6]] if len(sys.argv) > 5 else [256, 128, 64]
g = torch.Generator(device=DEV).manual_seed(0)
ldx = (K0 + 3) // 4 * 4
x = torch.zeros(M, ldx, device=DEV)
x[:, :K0] = torch.randn(M, K0, device=DEV, generator=g)
groups = []
for t in range(2):
    k, his, los, lds, bs, ys = K0, [], [], [], [], []
    for n in widths:
        ld = (k + 3) // 4 * 4
        w = torch.randn(n, k, device=DEV, generator=g) / k**0.5
        hi, lo = torch.zeros(n, ld, device=DEV), torch.zeros(n, ld, device=DEV)
        ops.split_tf32([(w, hi, lo, n, k, k, ld)])
        his.append(hi), los.append(lo), lds.append(ld), bs.append(torch.randn(n, device=DEV, generator=g)), ys.append(torch.empty(M, n, device=DEV))
        k = n
    groups.append(dict(x=x, ldx=ldx, w_hi=his, w_lo=los, ldw=lds, bias=bs, y=ys))


def fused():
    ops.tc_tower_fwd(groups, M, K0, widths, 1)


def layered():
    for i, n in enumerate(widths):
        k = K0 if i == 0 else widths[i - 1]
        ops.tc_linear_fwd([dict(a=x if i == 0 else G["y"][i - 1], b=G["w_hi"][i], b_lo=G["w_lo"][i], c=G["y"][i], bias=G["bias"][i], N=n, K=k,
                                lda=ldx if i == 0 else k, ldb=G["ldw"][i]) for G in groups], M, 1)


flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)


def timeit(fn, reps=10):
    ts = []
    for _ in range(reps):
        flush.sum()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return sum(ts[: max(1, reps * 3 // 4)]) / max(1, reps * 3 // 4)


for _ in range(3):
    fused(), layered()
torch.cuda.synchronize()
print(f"M={M} K0={K0} widths={widths}: fused {timeit(fused):.1f} us, three per-layer launches {timeit(layered):.1f} us (L2 flushed before each)")
lib.mnr_tc_tower_trace(1, None)
fused()
torch.cuda.synchronize()
buf = (ctypes.c_longlong * (4 * 256 * 2))()
lib.mnr_tc_tower_trace(0, buf)
tr = np.array(buf[:]).reshape(4, 256, 2)
t0 = tr[0, 0, 0]
rel = lambda v: (v - t0) / 1.9  # ~ns at 1.9 GHz
print("ns from the first layer-1 MMA issue of CTA 0 (issue times; the epilogue stamps are warp 8's)")
for it in range(8):
    if tr[2, it, 1] <= 0:
        break
    for h in range(widths[0] // 128):
        i = it * 4 + h
        print(f"tile {it} h{h}: L1 issue {rel(tr[0,i,0]):7.0f} {rel(tr[0,i,1]):7.0f} | epi1 {rel(tr[3,i,0]):7.0f} {rel(tr[3,i,1]):7.0f}")
    print(f"tile {it}   : L2 {rel(tr[1,it,0]):7.0f} {rel(tr[1,it,1]):7.0f} | epi2 {rel(tr[3,it*4+2,0]):7.0f} {rel(tr[3,it*4+2,1]):7.0f} | L3 {rel(tr[2,it,0]):7.0f} {rel(tr[2,it,1]):7.0f} | epi3 {rel(tr[3,it*4+3,0]):7.0f} {rel(tr[3,it*4+3,1]):7.0f}")

# fine-grained stamps of epilogue warp 0 (role-1 slots 64..): 8 per chunk
names = ["entry", "tmem ld done", "math done", "A slot free", "tcgen05.st issued", "box written + fence", "st done + a_full", "TMA store issued"]
fine = tr[1, 64:, :].reshape(-1)
n = int((fine > 0).sum()) // 8
print("epilogue warp 0, per chunk (ns since the chunk's entry; chunks in program order: L1h0 c0, c2, L1h1 c0, c2, L2 c0, c2, L3 c0, ...):")
for c in range(min(n, 28)):
    st = fine[8 * c: 8 * c + 8]
    print(f"chunk {c:2d} @ {rel(st[0]):7.0f}: " + "  ".join(f"{names[k]} +{(st[k] - st[0]) / 1.9:.0f}" for k in range(1, 8)))
