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


@pytest.mark.parametrize("M,N,K,act", CASES[:5] + [(32768, 128, 256, 1)])
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
        scale_close(grp["bias_grad"], rb, 3e-5, f"db {M}x{N}x{K}")
