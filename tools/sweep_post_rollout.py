"""Tuning aid: time the fused post-rollout kernel at cfg2 for a few tile plans (env knobs read by make_plan)."""
import os, statistics, subprocess, sys
sys.path.insert(0, "/root/repo")
if len(sys.argv) > 1:
    import torch
    from mineral_b200.agents.ppo.experience import ExperienceBuffer
    T, N, O, A = 32, int(sys.argv[1]), 37, 8
    dev = "cuda:0"
    buf = ExperienceBuffer(N, T, N * T, N * T // 4, {"obs": (O,)}, A, dev)
    lv = torch.zeros(N, 1, device=dev)
    flush = torch.zeros(64 << 20, device=dev)
    ms = []
    for i in range(24):
        flush.max(); torch.cuda._sleep(400_000)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); buf._launch_fused_only(lv, 0.99, 0.95); b.record(); torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    m = statistics.mean(sorted(ms[4:])[:15])
    row = 4 * O + 12 * A + 17
    byts = row * N * T + 4 * N + (row + 8) * N * T
    print(f"N={N} tile_kb={os.environ.get('MNR_POST_ROLLOUT_TILE_KB')} min_copy={os.environ.get('MNR_POST_ROLLOUT_MIN_COPY')} {m*1e3:.2f} us {byts/m/1e6:.0f} GB/s frac {byts/m/1e6/6532.9:.3f}", flush=True)
else:
    for n in (4096, 8192, 16384):
        for skip in (0, 1, 2, 3):
            env = dict(os.environ, MNR_POST_ROLLOUT_SKIP=str(skip))
            print(f"skip={skip} (1: no scalar role, 2: no wide role, 3: empty kernel)", end=" ", flush=True)
            subprocess.run([sys.executable, __file__, str(n)], env=env)
