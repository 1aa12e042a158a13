"""cfg4 replay probe: a 4 M-transition Ant ring, uniform samples of `batch` rows with fused obs normalisation, L2 read-flushed
before every launch.  Prints CUDA-event times of the sample kernel alone (the statistics tables prepared once, indices given);
run under `ncu -k regex:sample_records` for the DRAM byte counters.  usage: replay_probe.py [batch]"""
import statistics
import sys

import torch

sys.path.insert(0, "/root/repo")
from mineral_b200 import ops  # noqa: E402
from mineral_b200._native import ROW_BOOL_TO_F32, ROW_F32, ROW_F32_NORMALIZE  # noqa: E402
from mineral_b200.buffers import ReplayBuffer  # noqa: E402
from mineral_b200.common.normalizers import RunningMeanStd  # noqa: E402

DEV = "cuda:0"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
O, A, C = 37, 8, 4_000_000
rb = ReplayBuffer({"obs": (O,)}, A, C, DEV)
rb.records.normal_()
rb.cur_capacity, rb.if_full = C, True
rms = RunningMeanStd((O,), eps=1e-4, with_clamp=True, initial_count="eps").to(DEV)
rms.update(rb.buf_obs["obs"][:65536])
m, d = rms._prepare()
f32 = dict(dtype=torch.float32, device=DEV)
obs, nobs, act, rew, done = torch.empty(B, O, **f32), torch.empty(B, O, **f32), torch.empty(B, A, **f32), torch.empty(B, 1, **f32), torch.empty(B, 1, **f32)
off, rec = rb._off, rb.records
fields = [(rec, obs, O, ROW_F32_NORMALIZE, m, d, -5.0, 5.0, 0), (rec, nobs, O, ROW_F32_NORMALIZE, m, d, -5.0, 5.0, 0),
          (rec, act, A, ROW_F32), (rec, rew, 1, ROW_F32), (rec, done, 1, ROW_BOOL_TO_F32)]
offsets = [off[("obs", "obs")], off[("next_obs", "obs")], off["action"], off["reward"], off["done"]]
flush = torch.zeros(64 << 20, device=DEV)
ts = []
for i in range(12):
    idx = torch.randint(C, (B,), device=DEV)
    flush.max()
    torch.cuda._sleep(400_000)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    ops.replay_sample_records(fields, offsets, rec, idx)
    b.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b) * 1e3)
t = statistics.mean(sorted(ts)[:9])
print(f"sample_records_kernel batch {B}: {t:.2f} us per launch (event pair incl. ~5 us launch/record latency); "
      f"{677 * B / t / 1e3:.1f} GB/s algorithmic (677 B/row), record bytes {384 * B / 1e6:.2f} MB")
