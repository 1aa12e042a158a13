"""Fused heads + loss kernel (csrc/heads_fused.cu): time per launch and the CTA-0 phase timeline (HF_TRACE).
usage: trace_heads.py [rows] [A] [H] [fixed_sigma 0|1]"""
import ctypes
import sys

import torch

sys.path.insert(0, "/root/repo")
from mineral_b200 import ops  # noqa: E402
from mineral_b200._native import lib  # noqa: E402
from mineral_b200.agents.ppo.models import ActorCritic  # noqa: E402
from mineral_b200.nets.fused_towers import FusedTowers  # noqa: E402

DEV = "cuda:0"
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
A = int(sys.argv[2]) if len(sys.argv) > 2 else 8
H = int(sys.argv[3]) if len(sys.argv) > 3 else 64
fixed = bool(int(sys.argv[4])) if len(sys.argv) > 4 else True
torch.manual_seed(0)
O = 37
model = ActorCritic({"obs": (O,)}, A, mlp_kwargs=dict(units=[32, H], act_type="ELU"), fixed_sigma=fixed, actor_dist_kwargs=dict(dist_type="normal")).to(DEV)
for p in model.parameters():
    p.grad = torch.zeros_like(p)
ft = FusedTowers(model, rows, DEV, use_tc=False)
g = torch.Generator(device=DEV).manual_seed(1)
x = torch.randn(rows, O, device=DEV, generator=g)
d = dict(actions=torch.randn(rows, A, device=DEV, generator=g), neglogp=torch.randn(rows, device=DEV, generator=g) + 8,
         advantages=torch.randn(rows, device=DEV, generator=g), values=torch.randn(rows, 1, device=DEV, generator=g),
         returns=torch.randn(rows, 1, device=DEV, generator=g), mu=torch.randn(rows, A, device=DEV, generator=g),
         sigma=torch.rand(rows, A, device=DEV, generator=g) + 0.2)
cfg = ops.make_loss_cfg("normal", 0.1, 1.0, 0.2, False, True, 4.0, 0.0, 1e-4, "bound")
g_mu, g_ls, g_v = torch.zeros(rows, A, device=DEV), torch.zeros(rows, A, device=DEV), torch.zeros(rows, 1, device=DEV)
ft.forward_towers(x)
m, mm, sm = torch.zeros(10, device=DEV), torch.zeros(A, device=DEV), torch.zeros(A, device=DEV)
ws = ops.ppo_heads_workspace(rows, A, H, DEV)
run = lambda: ft.heads_loss_fwd_bwd(d, 0, rows, g_mu, g_ls, g_v.view(-1), m, cfg, ws, mm, sm)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
for warm in (False, True):
    ts = []
    for _ in range(10):
        if not warm:
            flush.sum()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); run(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    print(f"heads rows={rows} A={A} H={H} fixed_sigma={fixed}: {sorted(ts)[len(ts) // 2]:.1f} us per launch ({'L2 warm' if warm else 'L2 flushed'})")
lib.mnr_heads_trace(1, None)
flush.sum()
run()
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 32)()
lib.mnr_heads_trace(0, buf)
names = ["start", "Y3a tile staged", "mu dots done", "loss fwd loop done", "loss bwd loop done", "row pieces stored + sync", "dW outer products + sync",
         "dY3a in smem + sync", "dY3a out / Y3v in + sync", "value + g_v + sync", "dWv, column sums + sync", "tile loop done", "partials parked",
         "grid barrier passed", "output reduce done", "metrics written"]
names += ["(fold done, per-CTA partial starts)", "(metrics: partials loaded)", "(metrics: first reduce done)", "(metrics: pass 2 loaded)", "(metrics: second reduce done)"]
t0 = buf[0]
prev = t0
order = sorted(range(len(names)), key=lambda i: buf[i])
for i in order:
    print(f"{i:2d} {names[i]:40s} {buf[i] - t0:7d} ns  (+{buf[i] - prev})")
    prev = buf[i]
