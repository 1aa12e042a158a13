"""tcgen05 3xTF32 linear kernels (csrc/tc_linear.cu) through the C ABI against float64 torch."""

import pytest
import torch

from conftest import scale_close

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _act(z, act):
    return {0: z, 1: torch.nn.functional.elu(z), 2: torch.relu(z)}[act]


def _dact(y, act):
    if act == 1:
        return torch.where(y > 0, torch.ones_like(y), y + 1)
    if act == 2:
        return (y > 0).to(y.dtype)
    return torch.ones_like(y)


def _split(w, ld):
    from mineral_b200 import ops

    rows, cols = w.shape
    hi, lo = torch.empty(rows, ld, device=DEV), torch.empty(rows, ld, device=DEV)
    ops.split_tf32([(w.contiguous(), hi, lo, rows, cols, cols, ld)])
    return hi, lo


def _pad(x, ld):
    out = torch.zeros(x.shape[0], ld, device=DEV)
    out[:, : x.shape[1]] = x
    return out


CASES = [(256, 128, 64, 1), (300, 256, 37, 1), (1024, 64, 128, 2), (128, 256, 256, 0), (4096, 128, 256, 1), (77, 16, 8, 1)]


@pytest.mark.parametrize("M,N,K,act", CASES)
def test_tc_split_and_fwd(M, N, K, act):
    from mineral_b200 import ops

    if not ops.tc_supported():
        pytest.skip("tensor-map driver entry point unavailable")
    g = torch.Generator(device=DEV).manual_seed(M + N + K)
    ld = (K + 3) // 4 * 4
    groups, refs = [], []
    for t in range(2):
        x = torch.randn(M, K, device=DEV, generator=g) * 2
        w = torch.randn(N, K, device=DEV, generator=g) / K**0.5
        b = torch.randn(N, device=DEV, generator=g)
        hi, lo = _split(w, ld)
        assert torch.equal((hi + lo)[:, :K], w) and (hi.view(torch.int32) & 0x1FFF).abs().max().item() == 0
        y = torch.full((M, N), float("nan"), device=DEV)
        groups.append(dict(a=_pad(x, ld), b=hi, b_lo=lo, c=y, bias=b, N=N, K=K, lda=ld, ldb=ld))
        refs.append(_act(x.double() @ w.double().T + b.double(), act))
    ops.tc_linear_fwd(groups, M, act)
    torch.cuda.synchronize()
    for grp, ref in zip(groups, refs):
        scale_close(grp["c"], ref, 5e-6, f"fwd {M}x{N}x{K}")


@pytest.mark.parametrize("M,N,K,act", CASES[:5])
def test_tc_bwd_input(M, N, K, act):
    from mineral_b200 import ops

    if not ops.tc_supported():
        pytest.skip("tensor-map driver entry point unavailable")
    if N % 4 or K % 4:
        K = (K + 3) // 4 * 4
    g = torch.Generator(device=DEV).manual_seed(7 + M + N + K)
    groups, refs = [], []
    for t in range(2):
        dy = torch.randn(M, N, device=DEV, generator=g)
        y = torch.randn(M, N, device=DEV, generator=g)
        w = torch.randn(N, K, device=DEV, generator=g) / N**0.5
        hi, lo = _split(w, K)
        dx = torch.full((M, K), float("nan"), device=DEV)
        groups.append(dict(a=dy, saved=y, b=hi, b_lo=lo, c=dx, N=N, K=K))
        refs.append((dy.double() * _dact(y.double(), act)) @ w.double())
    ops.tc_linear_bwd_input(groups, M, act)
    torch.cuda.synchronize()
    for grp, ref in zip(groups, refs):
        scale_close(grp["c"], ref, 5e-6, f"dx {M}x{N}x{K}")


# both forms of the weight-gradient kernel: N % 32 == 0 -> tc_dw2_kernel (X^T in tensor memory, one / two lane tiles, K and N
# tiling, ragged K, ragged M), else the first form
DW_CASES = CASES[:5] + [(32768, 128, 256, 1), (32768, 64, 128, 1), (32768, 256, 37, 1), (2048, 512, 88, 0), (2048, 256, 512, 0),
                        (1000, 48, 100, 1), (200, 96, 200, 1), (70, 32, 300, 2)]


@pytest.mark.parametrize("M,N,K,act", DW_CASES)
def test_tc_bwd_weight(M, N, K, act):
    from mineral_b200 import ops

    if not ops.tc_supported():
        pytest.skip("tensor-map driver entry point unavailable")
    g = torch.Generator(device=DEV).manual_seed(11 + M + N + K)
    ld = (K + 3) // 4 * 4
    groups, refs = [], []
    ws = torch.zeros(ops.lib.mnr_tc_linear_bwd_weight_workspace_bytes(2, M, N, K), dtype=torch.uint8, device=DEV)
    for t in range(2):
        dy = torch.randn(M, N, device=DEV, generator=g)
        y = torch.randn(M, N, device=DEV, generator=g)
        x = torch.randn(M, K, device=DEV, generator=g)
        dw = torch.full((N, K), float("nan"), device=DEV)
        db = torch.full((N,), float("nan"), device=DEV)
        groups.append(dict(a=dy, saved=y, b=_pad(x, ld), c=dw, bias_grad=db, N=N, K=K, ldb=ld))
        dz = dy.double() * _dact(y.double(), act)
        refs.append((dz.T @ x.double(), dz.sum(0)))
    ops.tc_linear_bwd_weight(groups, M, act, ws)
    torch.cuda.synchronize()
    for grp, (rw, rb) in zip(groups, refs):
        scale_close(grp["c"], rw, 5e-6, f"dw {M}x{N}x{K}")
        scale_close(grp["bias_grad"], rb, 1e-5, f"db {M}x{N}x{K}")


# ---- SIMT linear entry points (csrc/mlp.cu): the thin-layer kernels (N <= 32: heads) and the tile kernel, same API ----------
THIN = [(8192, 8, 128), (33, 1, 64), (1000, 21, 64), (257, 16, 128), (4096, 32, 256), (77, 4, 512), (512, 3, 16), (64, 12, 36)]
WIDE = [(300, 48, 37), (129, 64, 45)]   # not thin (N > 32 or K % 4 != 0): the 128x64 tile kernel


@pytest.mark.parametrize("M,N,K", THIN + WIDE)
def test_simt_linear_fwd_dx_dw_thin_and_tile(M, N, K):
    """mnr_linear_fwd / _bwd_input (+accumulate) / _bwd_weight for head-shaped layers against float64 torch; two groups per
    launch with different N in the forward (policy + value heads share a launch), padded input row stride."""
    from mineral_b200 import ops

    g = torch.Generator(device=DEV).manual_seed(M * 7 + N * 3 + K)
    ld = (K + 3) // 4 * 4 + 4
    x = _pad(torch.randn(M, K, device=DEV, generator=g), ld)
    xs = x[:, :K]
    w1, b1 = torch.randn(N, K, device=DEV, generator=g) / K**0.5, torch.randn(N, device=DEV, generator=g)
    w2, b2 = torch.randn(1, K, device=DEV, generator=g) / K**0.5, torch.randn(1, device=DEV, generator=g)
    y1, y2 = torch.full((M, N), float("nan"), device=DEV), torch.full((M, 1), float("nan"), device=DEV)
    ops.linear_fwd([dict(a=x, b=w1, c=y1, bias=b1, N=N, K=K, lda=ld), dict(a=x, b=w2, c=y2, bias=b2, N=1, K=K, lda=ld)], M, 0)
    scale_close(y1, xs.double() @ w1.double().T + b1.double(), 2e-6, "fwd head 1")
    scale_close(y2, xs.double() @ w2.double().T + b2.double(), 2e-6, "fwd head 2")
    # input gradient, then accumulate a second head's contribution onto it
    g1, g2 = torch.randn(M, N, device=DEV, generator=g), torch.randn(M, 1, device=DEV, generator=g)
    dx = torch.full((M, K), float("nan"), device=DEV)
    ops.linear_bwd_input([dict(a=g1, b=w1, c=dx, N=N, K=K)], M, 0)
    scale_close(dx, g1.double() @ w1.double(), 2e-6, "dx")
    ops.linear_bwd_input([dict(a=g2, b=w2, c=dx, N=1, K=K)], M, 0, accumulate=True)
    scale_close(dx, g1.double() @ w1.double() + g2.double() @ w2.double(), 3e-6, "dx accumulate")
    # weight / bias gradients (split over row slabs above 512 rows, fixed-order reduce)
    ws = torch.zeros(max(ops.lib.mnr_linear_bwd_weight_workspace_bytes(2, M, N, K), 256), dtype=torch.uint8, device=DEV)
    ga, gb = g1, torch.randn(M, N, device=DEV, generator=g)
    dwa, dba = torch.full((N, K), float("nan"), device=DEV), torch.full((N,), float("nan"), device=DEV)
    dwb, dbb = torch.full((N, K), float("nan"), device=DEV), torch.full((N,), float("nan"), device=DEV)
    ops.linear_bwd_weight([dict(a=ga, b=x, c=dwa, bias_grad=dba, N=N, K=K, ldb=ld), dict(a=gb, b=x, c=dwb, bias_grad=dbb, N=N, K=K, ldb=ld)], M, 0, ws)
    for gg, dw, db in ((ga, dwa, dba), (gb, dwb, dbb)):
        scale_close(dw, gg.double().T @ xs.double(), 3e-6, "dw")
        scale_close(db, gg.double().sum(0), 3e-6, "db")
    again = torch.empty_like(dwa)
    ops.linear_bwd_weight([dict(a=ga, b=x, c=again, bias_grad=dba, N=N, K=K, ldb=ld)], M, 0, ws)
    assert torch.equal(again, dwa), "weight gradient is not deterministic"


# (256, 128, *) and (128, 128, *): 16 epilogue warps (A-slot uses per tile % 4 == 0); (*, 64, *): the 8-warp form; ragged M; K0 = 64
TOWER_CASES = [(4096, 37, (256, 128, 64), 1), (300, 37, (256, 128, 64), 1), (1024, 64, (128, 64, 32), 2), (2048, 20, (256, 64, 128), 0),
               (32768, 37, (256, 128, 64), 1), (4096, 40, (256, 128, 128), 1), (1000, 37, (128, 128, 64), 2), (20000, 64, (256, 128, 32), 0)]


@pytest.mark.parametrize("M,K0,widths,act", TOWER_CASES)
def test_tc_tower_fwd_vs_float64_and_per_layer(M, K0, widths, act):
    """csrc/tc_tower.cu (whole tower forward in one launch, activations handed on through tensor memory) against float64
    torch and against three mnr_tc_linear_fwd launches on the same split weights."""
    from mineral_b200 import ops

    if not ops.tc_supported() or not ops.tc_tower_fwd_supported(K0, list(widths), act):
        pytest.skip("fused tower forward unavailable")
    g = torch.Generator(device=DEV).manual_seed(M + K0 + sum(widths))
    ldx = (K0 + 3) // 4 * 4
    x = torch.randn(M, K0, device=DEV, generator=g) * 2
    xp = _pad(x, ldx)
    groups, refs, layered = [], [], []
    for t in range(2):
        k, ws, bs, his, los, lds, ys, ys2 = K0, [], [], [], [], [], [], []
        ref = x.double()
        for n in widths:
            w = torch.randn(n, k, device=DEV, generator=g) / k**0.5
            b = torch.randn(n, device=DEV, generator=g) * 0.5
            ld = (k + 3) // 4 * 4
            hi, lo = _split(w, ld)
            ws.append(w), bs.append(b), his.append(hi), los.append(lo), lds.append(ld)
            ys.append(torch.full((M, n), float("nan"), device=DEV)), ys2.append(torch.full((M, n), float("nan"), device=DEV))
            ref = _act(ref @ w.double().T + b.double(), act)
            refs.append(ref)
            k = n
        groups.append(dict(x=xp, ldx=ldx, w_hi=his, w_lo=los, ldw=lds, bias=bs, y=ys))
        layered.append((his, los, lds, bs, ys2))
    ops.tc_tower_fwd(groups, M, K0, list(widths), act)
    for i, n in enumerate(widths):   # the per-layer kernels on the same operands
        k = K0 if i == 0 else widths[i - 1]
        ops.tc_linear_fwd([dict(a=xp if i == 0 else L[4][i - 1], b=L[0][i], b_lo=L[1][i], c=L[4][i], bias=L[3][i], N=n, K=k,
                                lda=ldx if i == 0 else k, ldb=L[2][i]) for L in layered], M, act)
    torch.cuda.synchronize()
    for t in range(2):
        for i in range(3):
            y = groups[t]["y"][i]
            assert torch.isfinite(y).all(), f"tower {t} layer {i}: non-finite output"
            scale_close(y, refs[3 * t + i], 5e-6, f"tower fwd {M}x{K0}->{widths} layer {i}")
            scale_close(y, layered[t][4][i], 2e-6, f"tower fwd vs per-layer, layer {i}")
