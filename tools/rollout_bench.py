#!/usr/bin/env python
"""Agent-side cost of a rollout: PPO.play_steps over a device-resident synthetic env (no simulator
time), fused (_RolloutEngine) vs the reference's torch op sequence.  Prints env-steps/s and launches per env step."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import WORKLOADS, make_cfg  # noqa: E402
from mineral_b200 import _native as nv  # noqa: E402
from mineral_b200.agents import PPO  # noqa: E402


class _Space:
    def __init__(self, shape):
        self.shape = tuple(shape)


class DeviceEnv:
    """Pre-drawn obs / reward / done streams already on the GPU: env.step costs one index, nothing else."""

    def __init__(self, N, O, A, T, dev):
        g = torch.Generator(device=dev).manual_seed(0)
        self.observation_space, self.action_space = _Space((O,)), _Space((A,))
        self.obs = torch.randn(T + 2, N, O, device=dev, generator=g)
        self.rew = torch.randn(T + 2, N, device=dev, generator=g)
        self.done = torch.rand(T + 2, N, device=dev, generator=g) < 0.02
        self.tout = torch.rand(T + 2, N, device=dev, generator=g) < 0.01
        self.t = 0

    def reset(self):
        self.t = 0
        return self.obs[0]

    def step(self, actions):
        self.t = self.t % (self.obs.shape[0] - 1) + 1
        t = self.t
        return self.obs[t], self.rew[t], self.done[t], {"time_outs": self.tout[t]}


def run(workload, fused, reps=5):
    w = WORKLOADS[workload]
    dev = "cuda:0"
    cfg = make_cfg(w, dev, False, cuda_graph=False)
    cfg.agent.ppo["fused_rollout"] = fused
    env = DeviceEnv(w["N"], w["O"], w["A"], w["T"], dev)
    torch.manual_seed(0)
    agent = PPO(cfg, env=env)
    # episode bookkeeping stays ON: the fused path keeps it on the device (mnr_episode_scores), the torch path pays the
    # reference's per-step `torch.where(dones)[0].tolist()` host sync
    agent.obs = agent._convert_obs(env.reset())
    agent.dones = torch.zeros(w["N"], dtype=torch.bool, device=dev)
    agent.set_eval()
    for _ in range(2):
        agent.play_steps()
    torch.cuda.synchronize()
    nv.LAUNCHES = 0
    t0 = time.perf_counter()
    for _ in range(reps):
        agent.play_steps()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    steps = w["T"] * w["N"]
    return dict(workload=workload, fused=fused, ms_per_rollout=dt * 1e3, env_steps_per_s=steps / dt,
                our_launches_per_env_step=nv.LAUNCHES / reps / w["T"])


if __name__ == "__main__":
    out = [run(wl, f) for wl in ("ppo_ant_4096x32", "ppo_allegro_16384x16") for f in (True, False)]
    print(json.dumps(out, indent=1))
