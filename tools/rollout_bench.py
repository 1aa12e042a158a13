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


def torch_play_steps(agent):
    """The reference's op sequence (mineral/agents/ppo/ppo.py:133-183) on the agent's torch modules: the comparison arm of
    this tool only - the product has no such path."""
    a = agent
    for n in range(a.horizon_len):
        obs_n = {k: a.obs_rms[k].normalize(v) for k, v in a.obs.items()} if a.normalize_input else a.obs
        model_out = a.model.act(obs_n)
        if a.normalize_value:
            model_out['values'] = a.value_rms.unnormalize(model_out['values'])
        a.storage.update_data('obses', n, a.obs)
        for k in ['actions', 'neglogp', 'values', 'mu', 'sigma']:
            a.storage.update_data(k, n, model_out[k])
        actions = torch.clamp(model_out['actions'], -1.0, 1.0)
        obs, r, a.dones, infos = a.env.step(actions)
        a.obs = a._convert_obs(obs)
        rewards = r.reshape(-1, 1)
        a.storage.update_data('dones', n, a.dones)
        shaped = a.reward_shaper(rewards.clone())
        if a.value_bootstrap and 'time_outs' in infos:
            shaped += a.gamma * model_out['values'] * infos['time_outs'].reshape(-1, 1).float()
        a.storage.update_data('rewards', n, shaped)
        done_indices = torch.where(a.dones)[0].tolist()   # the reference's per-step host round trip (ppo.py:165)
        a._host_metrics.update(a.epoch, a.env, a.obs, rewards.squeeze(-1), done_indices, infos)
    obs_n = {k: a.obs_rms[k].normalize(v) for k, v in a.obs.items()} if a.normalize_input else a.obs
    last = a.model.act(obs_n)['values']
    a.finish_rollout(a.value_rms.unnormalize(last) if a.normalize_value else last)


def run(workload, fused, reps=5):
    w = WORKLOADS[workload]
    dev = "cuda:0"
    cfg = make_cfg(w, dev, False, cuda_graph=False)
    env = DeviceEnv(w["N"], w["O"], w["A"], w["T"], dev)
    torch.manual_seed(0)
    agent = PPO(cfg, env=env)
    # episode bookkeeping stays ON: the fused path keeps it on the device (mnr_episode_scores), the torch path pays the
    # reference's per-step `torch.where(dones)[0].tolist()` host sync
    agent.obs = agent._convert_obs(env.reset())
    agent.dones = torch.zeros(w["N"], dtype=torch.bool, device=dev)
    agent.set_eval()
    from mineral_b200.agents.agent import Metrics
    agent._host_metrics = Metrics({'rewards': torch.zeros(w["N"], device=dev), 'lengths': torch.zeros(w["N"], dtype=torch.int64, device=dev)}, 100)
    play = agent.play_steps if fused else (lambda: torch_play_steps(agent))
    for _ in range(2):
        play()
    torch.cuda.synchronize()
    nv.LAUNCHES = 0
    t0 = time.perf_counter()
    for _ in range(reps):
        play()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    steps = w["T"] * w["N"]
    return dict(workload=workload, fused=fused, ms_per_rollout=dt * 1e3, env_steps_per_s=steps / dt,
                our_launches_per_env_step=nv.LAUNCHES / reps / w["T"])


if __name__ == "__main__":
    out = [run(wl, f) for wl in ("ppo_ant_4096x32", "ppo_allegro_16384x16") for f in (True, False)]
    print(json.dumps(out, indent=1))
