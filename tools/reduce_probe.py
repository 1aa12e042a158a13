"""Time the split-reduce launch of one cfg2 optimizer step alone (CUDA events, L2 flushed).  usage: reduce_probe.py [rows]
MINERAL_B200_LIB=<variant .so> selects an A/B build."""
import sys

import torch

sys.path.insert(0, "/root/repo")
from mineral_b200 import ops  # noqa: E402
from mineral_b200.agents.ppo.models import ActorCritic  # noqa: E402
from mineral_b200.nets.fused_towers import FusedTowers  # noqa: E402

DEV = "cuda:0"
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
O, A = 37, 8
torch.manual_seed(0)
model = ActorCritic({"obs": (O,)}, A, mlp_kwargs=dict(units=[256, 128, 64], act_type="ELU"), fixed_sigma=True,
                    actor_dist_kwargs=dict(dist_type="normal")).to(DEV)
for p in model.parameters():
    p.grad = torch.zeros_like(p)
ft = FusedTowers(model, rows, DEV, use_tc=True)
x = torch.zeros(rows, 40, device=DEV)
x[:, :O] = torch.randn(rows, O, device=DEV)
ft.forward_towers(x, 40)
for t in ft.dY[-1]:
    t.normal_()
captured = []
real = ops.tc_reduce_partials
ops.tc_reduce_partials = lambda pending, folds=None: captured.append(pending)
ft.backward_towers(x)
ops.tc_reduce_partials = real
pending = captured[0]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
ts = []
for _ in range(20):
    flush.sum()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); real(pending); b.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b) * 1e3)
ts.sort()
print(f"split reduce rows={rows}: median {ts[len(ts) // 2]:.2f} us, min {ts[0]:.2f} us per launch (events, L2 flushed)")
