"""SAC / DDPG update_net on the B200 kernels at BASELINE cfg4 (Ant-shaped: obs 37, act 8; circular replay of 4 M transitions,
uniform sample batch 8192 with obs normalisation; twin-Q / actor towers [512,256,128] LayerNorm+SiLU as in
mineral/cfgs/agent/AllegroHandSAC.yaml; 8 mini-epochs per call), timed with CUDA events:

  fused_graph   the product path: one CUDA graph of our kernels per update_net call
  fused_eager   the same launches without the graph
  torch_ops     the reference's op sequence (sac.py:251-345: torch modules, autograd, clip_grad_norm_, torch.optim.AdamW,
                torch index gathers) on the same device - what the unmodified reference costs on a B200

    python tools/offpolicy_bench.py [--algo sac|ddpg] [--iters 20] [--capacity 4000000] [--no-torch] [--json out.json]
"""
import argparse
import copy
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class AD(dict):
    __getattr__ = dict.__getitem__

    def __setattr__(self, k, v):
        self[k] = v


def to_attr(x):
    return AD({k: to_attr(v) for k, v in x.items()}) if isinstance(x, dict) else x


class _Space:
    def __init__(self, shape):
        self.shape = tuple(shape)


class NullEnv:
    def __init__(self, O, A):
        self.observation_space, self.action_space = _Space((O,)), _Space((A,))


def make_cfg(algo, N, O, A, B, capacity, mini_epochs, graph):
    mlp = dict(units=[512, 256, 128], norm_type="LayerNorm", act_type="SiLU")
    if algo == "sac":
        network = dict(normalize_input=True, encoder=None, actor="Actor", critic="EnsembleQ",
                       actor_kwargs=dict(mlp_kwargs=mlp, weight_init="orthogonal", tanh_policy=False, fixed_sigma=False,
                                         dist_kwargs=dict(dist_type="squashed_normal", minlogstd=-5.0, maxlogstd=2.0)),
                       critic_kwargs=dict(n_critics=2, mlp_kwargs=mlp, weight_init="orthogonal"))
        extra = dict(init_alpha=1.0, alpha=None, alpha_optim_kwargs=dict(lr=5e-3), backup_entropy=True)
    else:
        network = dict(normalize_input=True, encoder=None, actor="Actor", critic="EnsembleQ",
                       actor_kwargs=dict(tanh_policy=True, mlp_kwargs=mlp), critic_kwargs=dict(n_critics=2, mlp_kwargs=mlp))
        extra = dict(exp_decay_rate=0.99, noise=dict(type="mixed", decay=None, exp_decay_rate=0.99, lin_decay_iters=10000, std_max=0.8,
                                                     std_min=0.05, tgt_pol_std=0.8, tgt_pol_noise_bound=0.2))
    algo_cfg = dict(multi_gpu=False, num_actors=N, reward_shaper=dict(fn="scale", scale=0.01), max_agent_steps=1e9, horizon_len=1,
                    memory_size=capacity, batch_size=B, nstep=3, tau=0.05, gamma=0.99, warm_up=32, no_tgt_actor=True, handle_timeout=True,
                    update_actor_interval=1, update_targets_interval=1, mini_epochs=mini_epochs, optim_type="AdamW",
                    actor_optim_kwargs=dict(lr=5e-4), critic_optim_kwargs=dict(lr=5e-4), max_grad_norm=0.5, cuda_graph=graph, **extra)
    return to_attr(dict(rl_device="cuda:0", multi_gpu=False, env_render=False, task=dict(env_autoresets=True),
                        agent={"tracker_len": 100, "print_every": -1, "ckpt_every": -1, "network": network, algo: algo_cfg}))


def fill_memory(agent, capacity, O, A):
    """Pre-fill the ring with N(0,1) rows without going through add_to_buffer."""
    m = agent.memory
    g = torch.Generator(device="cuda").manual_seed(0)
    step = 500_000
    for s in range(0, capacity, step):
        e = min(capacity, s + step)
        m.buf_obs["obs"][s:e].normal_(generator=g)
        m.buf_next_obs["obs"][s:e].normal_(generator=g)
        m.buf_action[s:e].uniform_(-1, 1, generator=g)
        m.buf_reward[s:e].normal_(generator=g)
        m.buf_done[s:e] = torch.rand(e - s, 1, device="cuda", generator=g) < 0.02
    m.next_p, m.if_full, m.cur_capacity = 0, True, capacity
    agent.obs_rms["obs"].update(m.buf_obs["obs"][:65536])


def time_calls(fn, iters, warm):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(iters):
        fn()
    t1.record()
    torch.cuda.synchronize()
    return t0.elapsed_time(t1) / iters


def torch_sac_update(agent, st):
    """The reference's SAC.update_net (sac.py:251-345) as torch ops on copies of the same modules."""
    import torch.nn.functional as F
    cfg, m = agent.sac_config, agent.memory
    rms = agent.obs_rms["obs"]

    def norm(x):
        y = (x - rms.running_mean.float()) / torch.sqrt(rms.running_var.float() + rms.eps)
        return torch.clamp(y, -5.0, 5.0)

    def opt_step(opt, loss):
        opt.zero_grad(set_to_none=True)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(opt.param_groups[0]["params"], cfg.max_grad_norm)
        opt.step()

    for _ in range(cfg.mini_epochs):
        idx = torch.randint(m.cur_capacity, (agent.batch_size,), device="cuda")
        obs, action, reward = m.buf_obs["obs"][idx], m.buf_action[idx], m.buf_reward[idx]
        next_obs, done = m.buf_next_obs["obs"][idx], m.buf_done[idx].float()
        with torch.no_grad():
            nz = norm(next_obs)
            _, _, d = st["actor"](nz)
            na = d.rsample()
            lp = d.log_prob(na).sum(-1, keepdim=True)
            tq = st["critic_target"].get_q_min(nz, na) - st["log_alpha"].exp().detach() * lp
            tq = reward + (1 - done) * (cfg.gamma ** cfg.nstep) * tq
        z = norm(obs)
        qs = st["critic"].get_q_values(z, action)
        opt_step(st["critic_opt"], torch.sum(torch.stack([F.mse_loss(q, tq) for q in qs])))
        st["critic"].requires_grad_(False)
        z = norm(obs)
        _, _, d = st["actor"](z)
        a = d.rsample()
        lp = d.log_prob(a).sum(-1, keepdim=True)
        q = st["critic"].get_q_min(z, a)
        opt_step(st["actor_opt"], (st["log_alpha"].exp().detach() * lp - q).mean())
        st["critic"].requires_grad_(True)
        opt_step(st["alpha_opt"], (st["log_alpha"].exp() * (-lp - agent.target_entropy).detach()).mean())
        with torch.no_grad():
            for t, c in zip(st["critic_target"].parameters(), st["critic"].parameters()):
                t.mul_(1.0 - cfg.tau).add_(c * cfg.tau)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--algo", default="sac")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--capacity", type=int, default=4_000_000)
    ap.add_argument("--batch", type=int, default=8192)
    ap.add_argument("--mini-epochs", type=int, default=8)
    ap.add_argument("--envs", type=int, default=4096)
    ap.add_argument("--no-torch", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--json", default=None)
    a = ap.parse_args()
    from mineral_b200 import _native as nv
    from mineral_b200.agents import DDPG, SAC
    O, A = 37, 8
    out = dict(workload=f"{a.algo}_update_net_ant_replay{a.capacity}_batch{a.batch}x{a.mini_epochs}", obs=O, act=A, mlp=[512, 256, 128],
               transitions_per_call=a.batch * a.mini_epochs)
    for mode in (["fused_eager"] if a.no_graph else ["fused_graph", "fused_eager"]):
        torch.manual_seed(0)
        agent = (SAC if a.algo == "sac" else DDPG)(make_cfg(a.algo, a.envs, O, A, a.batch, a.capacity, a.mini_epochs, mode == "fused_graph"),
                                                   env=NullEnv(O, A))
        fill_memory(agent, a.capacity, O, A)
        agent.set_train()
        l0 = nv.LAUNCHES
        agent.update_net(agent.memory)
        launches = nv.LAUNCHES - l0
        ms = time_calls(lambda: agent.update_net(agent.memory), a.iters, a.warmup)
        t0 = time.perf_counter()
        for _ in range(a.iters):
            agent.update_net(agent.memory)
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) / a.iters * 1e3
        out[mode] = dict(ms_per_update_net=ms, wall_ms=wall, transitions_per_s=a.batch * a.mini_epochs / ms * 1e3, kernels_per_call=launches)
        print(mode, out[mode], flush=True)
        if mode == "fused_eager" and a.algo == "sac" and not a.no_torch:
            st = dict(actor=copy.deepcopy(agent.actor), critic=copy.deepcopy(agent.critic), critic_target=copy.deepcopy(agent.critic_target),
                      log_alpha=torch.nn.Parameter(torch.zeros((), device="cuda")))
            for k in ("actor", "critic"):
                for p in st[k].parameters():
                    p.grad = None
            st["actor_opt"] = torch.optim.AdamW(st["actor"].parameters(), lr=5e-4)
            st["critic_opt"] = torch.optim.AdamW(st["critic"].parameters(), lr=5e-4)
            st["alpha_opt"] = torch.optim.AdamW([st["log_alpha"]], lr=5e-3)
            ms_t = time_calls(lambda: torch_sac_update(agent, st), max(3, a.iters // 4), 2)
            out["torch_ops"] = dict(ms_per_update_net=ms_t, transitions_per_s=a.batch * a.mini_epochs / ms_t * 1e3)
            print("torch_ops", out["torch_ops"], flush=True)
        del agent
        torch.cuda.empty_cache()
    if "torch_ops" in out and "fused_graph" in out:
        out["speedup_vs_torch_ops"] = out["torch_ops"]["ms_per_update_net"] / out["fused_graph"]["ms_per_update_net"]
    print(json.dumps(out))
    if a.json:
        json.dump(out, open(a.json, "w"), indent=1)


if __name__ == "__main__":
    main()
