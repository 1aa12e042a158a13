"""Generate tests/golden/*.pt by running the UNMODIFIED reference on CPU.

Run in the build container only (needs /root/reference):

    python oracle/make_golden.py

Each fixture stores seeded inputs and the reference's own outputs for one row of
SURVEY.md section 8(a).  The fixtures are small (KBs) and are committed; the GPU box
has no /root/reference, so tests read only these files.
"""

import copy
import os
import re
import sys
import tempfile
import types
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import reference_shim as rs  # noqa: E402

warnings.filterwarnings("ignore")
rs.install()

from mineral.agents.diffrl.shac import SHAC  # noqa: E402
from mineral.agents.diffrl.utils import CriticDataset  # noqa: E402
from mineral.agents.ppo.experience import ExperienceBuffer  # noqa: E402
from mineral.agents.ppo.ppo import PPO, actor_loss, bounds_loss, critic_loss, policy_kl  # noqa: E402
from mineral.buffers import NStepReplay, ReplayBuffer  # noqa: E402
from mineral.common.normalizers import RunningMeanStd  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
os.makedirs(OUT, exist_ok=True)


def save(name, obj):
    path = os.path.join(OUT, name + ".pt")
    torch.save(obj, path)
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KB")


def clone_tree(x):
    if isinstance(x, dict):
        return {k: clone_tree(v) for k, v in x.items()}
    if torch.is_tensor(x):
        return x.detach().clone()
    return copy.deepcopy(x)


# ---- K1a + K3a: ExperienceBuffer.compute_return / prepare_training -------------------------
def gen_experience():
    cases = []
    for seed, (T, N, O, A), done_p in [
        (0, (8, 24, 5, 3), 0.1),
        (1, (16, 12, 37, 8), 0.0),
        (2, (4, 20, 7, 2), 1.0),
        (3, (32, 8, 3, 1), 0.3),
    ]:
        g = torch.Generator().manual_seed(seed)
        buf = ExperienceBuffer(N, T, N * T, N * T // 4, {"obs": (O,)}, A, "cpu", re.compile("$^"))
        sd = buf.storage_dict
        sd["rewards"].copy_(torch.randn(T, N, 1, generator=g))
        sd["values"].copy_(3 * torch.randn(T, N, 1, generator=g))
        sd["neglogp"].copy_(torch.randn(T, N, generator=g))
        sd["dones"].copy_((torch.rand(T, N, generator=g) < done_p).to(torch.uint8))
        for k in ("actions", "mu", "sigma"):
            sd[k].copy_(torch.randn(T, N, A, generator=g))
        sd["obses"]["obs"].copy_(torch.randn(T, N, O, generator=g) * 4 + 1)
        last_values = 3 * torch.randn(N, 1, generator=g)
        inputs = clone_tree({k: v for k, v in sd.items() if k != "returns"})
        buf.compute_return(last_values, 0.99, 0.95)
        returns_tm = sd["returns"].clone()
        data = buf.prepare_training()
        mb1 = buf[1]
        cases.append(
            dict(
                shape=(T, N, O, A), gamma=0.99, tau=0.95, inputs=inputs, last_values=last_values,
                returns_time_major=returns_tm, data=clone_tree(data), minibatch_size=N * T // 4,
                minibatch1=clone_tree({k: v for k, v in zip(("values", "neglogp", "advantages", "mu", "sigma", "returns", "actions", "obses"), mb1)}),
                length=len(buf),
            )
        )
    save("experience", cases)


# ---- K2: RunningMeanStd ---------------------------------------------------------------------
def gen_rms():
    cases = []
    cfgs = {
        "ppo_f64": dict(eps=1e-5, with_clamp=True, initial_count=1, dtype=torch.float64),
        "sac_f32": dict(eps=1e-4, with_clamp=True, initial_count="eps"),
        "shac_legacy": dict(eps=1e-5, correction=0, initial_count=1e-4, dtype=torch.float32),
        "shac_f64_noclamp": dict(eps=1e-5, initial_count=1, dtype=torch.float64),
    }
    for ci, (name, kw) in enumerate(cfgs.items()):
        g = torch.Generator().manual_seed(10 + ci)
        for shape in [(37,), (1,), (4, 3)]:
            rms = RunningMeanStd(shape, **kw)
            steps = []
            for rows in (32, 7, 90):
                x = torch.randn(rows, *shape, generator=g) * torch.linspace(0.1, 9.0, int(np.prod(shape))).view(shape) + 2.5
                rms.update(x)
                y = rms.normalize(x)
                z = rms.unnormalize(y * 1.3)
                steps.append(dict(x=x, mean=rms.running_mean.clone(), var=rms.running_var.clone(), count=rms.running_count.clone(), normalized=y, unnormalized=z))
            cases.append(dict(name=name, kwargs={k: (str(v) if isinstance(v, torch.dtype) else v) for k, v in kw.items()}, shape=shape, steps=steps, state_dict_keys=list(rms.state_dict().keys())))
    save("rms", cases)


# ---- K1b: SHAC target values ------------------------------------------------------------------
def gen_tdlambda():
    cases = []
    for seed, (T, N), done_p in [(0, (8, 10), 0.1), (1, (32, 6), 0.0), (2, (5, 9), 0.5)]:
        g = torch.Generator().manual_seed(seed)
        rew = torch.randn(T, N, generator=g)
        dm = (torch.rand(T, N, generator=g) < done_p).float()
        dm[T - 1] = 1.0  # shac.py:789-792
        nv = 2 * torch.randn(T, N, generator=g)
        ent = torch.randn(T, N, generator=g) + 3
        lp = torch.randn(T, N, generator=g) - 3
        for method in ("td-lambda", "one-step"):
            ns = types.SimpleNamespace(critic_method=method, num_envs=N, device="cpu", horizon_len=T, lam=0.95, gamma=0.99,
                                       done_mask=dm, rew_buf=rew, next_values=nv, target_values=torch.zeros(T, N))
            SHAC.compute_target_values(ns)
            plain = ns.target_values.clone()
            variants = []
            for use_distr_ent, offset, scale, alpha in [(True, False, False, 0.2), (False, True, True, 0.05), (True, True, False, 0.7)]:
                ns2 = types.SimpleNamespace(critic_method=method, num_envs=N, device="cpu", horizon_len=T, lam=0.95, gamma=0.99,
                                            done_mask=dm, rew_buf=rew, next_values=nv, target_values=torch.zeros(T, N),
                                            distr_ent=ent, logprobs=lp, use_distr_ent=use_distr_ent, offset_by_target_entropy=offset,
                                            scale_by_target_entropy=scale, target_entropy=-4.0, with_autoent=False, entropy_coef=alpha)
                SHAC.compute_target_values_with_entropy(ns2)
                variants.append(dict(use_distr_ent=use_distr_ent, offset=offset, scale=scale, alpha=alpha, target=ns2.target_values.clone()))
            cases.append(dict(method=method, gamma=0.99, lam=0.95, rew=rew, done_mask=dm, next_values=nv, distr_ent=ent, logprobs=lp,
                              target_entropy=-4.0, target=plain, entropy_variants=variants))
    save("tdlambda", cases)


# ---- K3f: CriticDataset -------------------------------------------------------------------------
def gen_critic_dataset():
    cases = []
    g = torch.Generator().manual_seed(5)
    T, B, O = 6, 7, 4
    obs = {"obs": torch.randn(T, B, O, generator=g)}
    tv = torch.randn(T, B, generator=g)
    for bs, shuffle, drop_last in [(10, False, False), (10, False, True), (8, True, False)]:
        if shuffle:
            np.random.seed(123)
        ds = CriticDataset(bs, obs, tv, shuffle=shuffle, drop_last=drop_last)
        batches = [clone_tree(dict(obs=ds[i][0], target=ds[i][1])) for i in range(len(ds))]
        cases.append(dict(batch_size=bs, shuffle=shuffle, drop_last=drop_last, np_seed=123 if shuffle else None, obs=obs, target_values=tv, length=len(ds), batches=batches))
    save("critic_dataset", cases)


# ---- a16: SHAC.update_critic on the reference's Critic / EnsembleCritic (Humanoid shape, cfg3) ----------------
def gen_shac_critic():
    from mineral.agents.diffrl import models as ref_models

    cases = []
    T, B, O = 32, 64, 76  # BASELINE.json configs[2]
    for seed, n_critics, iters, n_batches in [(0, 1, 16, 4), (1, 2, 4, 4), (2, 1, 3, 3)]:
        torch.manual_seed(seed)
        mlp_kwargs = dict(units=[64, 64], norm_type="LayerNorm", act_type="ELU")
        if n_critics == 1:
            critic = ref_models.Critic(O, 21, mlp_kwargs=mlp_kwargs)
        else:
            critic = ref_models.EnsembleCritic(O, 21, n_critics=n_critics, mlp_kwargs=mlp_kwargs)
        from oracle.mineral_oracle import shac_critic_inputs
        o_, tv = shac_critic_inputs(seed, T, B, O, wild=(seed == 2))
        obs = {"obs": o_}
        bs = T * B // n_batches if seed != 2 else 700  # seed 2: ragged last batch (2048 = 700 + 700 + 648)
        init = clone_tree(critic.state_dict())
        optim = torch.optim.Adam(critic.parameters(), lr=2e-3, betas=(0.7, 0.95))  # DFlexAntSHAC.yaml:46-47
        cfg = rs.to_attr(dict(truncate_grads=True, max_grad_norm=1.0))
        ns = types.SimpleNamespace(critic_iterations=iters, critic=critic, critic_optim=optim, shac_config=cfg, encoder=lambda o: o["obs"])
        ns.compute_critic_loss = lambda o, t: SHAC.compute_critic_loss(ns, o, t)
        ds = CriticDataset(bs, obs, tv, shuffle=False, drop_last=False)
        res = SHAC.update_critic(ns, ds)
        cases.append(dict(n_critics=n_critics, iterations=iters, batch_size=bs, shape=(T, B, O), seed=seed, wild=(seed == 2),
                          init_state_dict=init, final_state_dict=clone_tree(critic.state_dict()),
                          results={k: torch.stack([x.detach() for x in v]) for k, v in res.items()}, lr=2e-3, betas=(0.7, 0.95),
                          max_grad_norm=1.0))
    save("shac_critic", cases)


# ---- 8(f)-2: SAC / DDPG update math (soft_update, distl_projection, the SAC critic target) ------------------------
def gen_offpolicy():
    import torch.nn as nn
    import torch.nn.functional as F_
    from mineral.agents.ddpg.utils import distl_projection, soft_update
    from mineral.agents.sac import sac as sac_mod

    g = torch.Generator().manual_seed(77)
    out = {}
    # Polyak: two small nets
    tar, cur = nn.Linear(7, 5), nn.Linear(7, 5)
    before = clone_tree(tar.state_dict())
    soft_update(tar, cur, 0.05)
    out["soft_update"] = dict(tau=0.05, target_before=before, current=clone_tree(cur.state_dict()), target_after=clone_tree(tar.state_dict()))
    # C51 projection
    B, N = 37, 51
    nd = torch.softmax(torch.randn(B, N, generator=g), dim=1)
    rew = 3 * torch.randn(B, 1, generator=g)
    rew[0, 0], rew[1, 0] = 40.0, -40.0                      # clamped at both ends
    done = (torch.rand(B, 1, generator=g) < 0.3).float()
    support = torch.linspace(-10, 10, N)
    rew[2, 0], done[2, 0] = 0.0, 1.0                        # lands exactly on an atom (l == u fix-ups)
    out["distl"] = dict(next_dist=nd, reward=rew, done=done, gamma=0.99 ** 3, v_min=-10, v_max=10, num_atoms=N, support=support,
                        proj=distl_projection(nd, rew, done, 0.99 ** 3, -10, 10, N, support))
    # SAC critic target: the real SAC.update_critic on a namespace, target_Q captured at F.mse_loss
    B = 64
    q1, q2 = torch.randn(B, 1, generator=g) * 5, torch.randn(B, 1, generator=g) * 5
    logp = torch.randn(B, 1, generator=g) - 4
    rew, done = torch.randn(B, 1, generator=g), (torch.rand(B, 1, generator=g) < 0.2).float()
    cases = []
    for backup_entropy, alpha in [(True, 0.2), (False, 0.2), (True, 0.013)]:
        seen = {}
        real_mse = F_.mse_loss

        def spy_mse(a, b, *args, **kw):
            seen["target"] = b.clone()
            return real_mse(a, b, *args, **kw)

        ns = types.SimpleNamespace(
            normalize_input=False, encoder_target=lambda o: o, encoder=lambda o: o,
            get_actions=lambda z=None, logprob=False: (None, None, logp),
            critic_target=types.SimpleNamespace(get_q_min=lambda z, a: torch.min(torch.stack([q1, q2]), dim=0).values),
            critic=types.SimpleNamespace(get_q_values=lambda z, a: [torch.zeros(B, 1), torch.zeros(B, 1)]),
            sac_config=rs.to_attr(dict(backup_entropy=backup_entropy, gamma=0.99, nstep=3, alpha=alpha)),
            critic_optim=None, optimizer_update=lambda opt, loss: torch.zeros(()))
        ns.get_alpha = lambda detach=True, scalar=False: sac_mod.SAC.get_alpha(ns, detach, scalar)
        sac_mod.F.mse_loss = spy_mse   # (sac.py's `F` IS torch.nn.functional: restored right below)
        try:
            sac_mod.SAC.update_critic(ns, {"obs": torch.zeros(B, 3)}, torch.zeros(B, 2), rew, {"obs": torch.zeros(B, 3)}, done)
        finally:
            sac_mod.F.mse_loss = real_mse
        cases.append(dict(backup_entropy=backup_entropy, alpha=alpha, gamma=0.99, nstep=3, target=seen["target"]))
    out["sac_target"] = dict(q=[q1, q2], log_prob=logp, reward=rew, done=done, cases=cases)
    save("offpolicy", out)


# ---- K3c/K3d: ReplayBuffer --------------------------------------------------------------------------
def gen_replay():
    g = torch.Generator().manual_seed(7)
    O, A, C = 5, 2, 50
    rb = ReplayBuffer({"obs": (O,)}, A, C, device="cpu")
    for k in rb.buf_obs:
        rb.buf_obs[k].zero_(); rb.buf_next_obs[k].zero_()
    rb.buf_action.zero_(); rb.buf_reward.zero_(); rb.buf_done.zero_()
    events = []
    for rows in (12, 30, 20, 50, 3, 49):
        tr = ({"obs": torch.randn(rows, O, generator=g)}, torch.randn(rows, A, generator=g), torch.randn(rows, 1, generator=g),
              {"obs": torch.randn(rows, O, generator=g)}, (torch.rand(rows, 1, generator=g) < 0.3))
        rb.add_to_buffer(tr)
        torch.manual_seed(1000 + rows)
        idx = torch.randint(rb.cur_capacity, size=(16,), device="cpu")
        torch.manual_seed(1000 + rows)
        obs, act, rew, nobs, done = rb.sample_batch(16, device="cpu")
        events.append(dict(insert=clone_tree(dict(obs=tr[0]["obs"], act=tr[1], rew=tr[2], nobs=tr[3]["obs"], done=tr[4])),
                           next_p=rb.next_p, if_full=rb.if_full, cur_capacity=rb.cur_capacity,
                           buf=clone_tree(dict(obs=rb.buf_obs["obs"], act=rb.buf_action, rew=rb.buf_reward, nobs=rb.buf_next_obs["obs"], done=rb.buf_done)),
                           idx=idx, sample=clone_tree(dict(obs=obs["obs"], act=act, rew=rew, nobs=nobs["obs"], done=done))))
    save("replay", dict(obs_dim=O, act_dim=A, capacity=C, events=events))


# ---- K3e: NStepReplay -----------------------------------------------------------------------------------
def gen_nstep():
    cases = []
    for seed, (N, O, A, n, S_list, done_p) in enumerate([(6, 4, 2, 3, (4, 1, 5), 0.25), (5, 3, 1, 2, (1, 1, 3), 0.5), (4, 2, 2, 4, (2, 8), 0.1), (3, 2, 1, 1, (3,), 0.2)]):
        g = torch.Generator().manual_seed(20 + seed)
        ns = NStepReplay({"obs": (O,)}, A, N, n, device="cpu")
        ns.nstep_buf_obs["obs"].zero_(); ns.nstep_buf_next_obs["obs"].zero_()
        ns.nstep_buf_action.zero_(); ns.nstep_buf_reward.zero_(); ns.nstep_buf_done.zero_()
        calls = []
        for S in S_list:
            obs = torch.randn(N, S, O, generator=g); act = torch.randn(N, S, A, generator=g)
            rew = torch.randn(N, S, 1, generator=g); nobs = torch.randn(N, S, O, generator=g)
            done = (torch.rand(N, S, 1, generator=g) < done_p).float()  # sac.py:150 stores dones in a float tensor
            try:
                o, a, r, no, d = ns.add_to_buffer({"obs": obs}, act, rew, {"obs": nobs}, done)
                out = clone_tree(dict(obs=o["obs"], act=a, rew=r, nobs=no["obs"], done=d))
            except (RuntimeError, ValueError):  # reference raises (torch.cat of an empty list) while the window warms up
                out = None
            calls.append(dict(inp=dict(obs=obs, act=act, rew=rew, nobs=nobs, done=done), out=out))
        cases.append(dict(N=N, O=O, A=A, nstep=n, gamma=0.99, calls=calls))
    save("nstep", cases)


# ---- K4: loss functions + autograd gradients ------------------------------------------------------------
def gen_losses():
    cases = []
    for seed, (mb, A) in enumerate([(64, 8), (33, 3), (128, 16)]):
        g = torch.Generator().manual_seed(30 + seed)
        for e_clip, smooth, clip_v, btype in [(0.2, False, False, "bound"), (0.2, False, True, "bound"), (0.1, True, False, "reg")]:
            old_nlp = torch.randn(mb, generator=g) + 8
            nlp = (old_nlp + 0.3 * torch.randn(mb, generator=g)).requires_grad_(True)
            adv = torch.randn(mb, generator=g)
            vold = torch.randn(mb, 1, generator=g)
            v = (vold + 0.4 * torch.randn(mb, 1, generator=g)).requires_grad_(True)
            ret = torch.randn(mb, 1, generator=g)
            mu = (1.5 * torch.randn(mb, A, generator=g)).requires_grad_(True)
            sigma = torch.rand(mb, A, generator=g) + 0.1
            old_mu = mu.detach() + 0.1 * torch.randn(mb, A, generator=g)
            old_sigma = sigma * (1 + 0.1 * torch.rand(mb, A, generator=g))
            a, frac = actor_loss(old_nlp, nlp, adv, e_clip, smooth)
            c, ev = critic_loss(vold, v, e_clip, ret, clip_v)
            b = bounds_loss(mu, btype)
            kl = policy_kl(mu.detach(), sigma, old_mu, old_sigma)
            total = a.mean() + 0.5 * c.mean() * 4 + b.mean() * 1e-4
            total.backward()
            cases.append(dict(e_clip=e_clip, smooth=smooth, clip_value=clip_v, bounds_type=btype,
                              old_nlp=old_nlp, nlp=nlp.detach(), adv=adv, vold=vold, v=v.detach(), ret=ret, mu=mu.detach(), sigma=sigma, old_mu=old_mu, old_sigma=old_sigma,
                              a_loss=a.detach(), clip_frac=frac, c_loss=c.detach(), explained_var=ev, b_loss=b.detach(), kl=kl, total=total.detach(),
                              g_nlp=nlp.grad.clone(), g_v=v.grad.clone(), g_mu=mu.grad.clone()))
    save("losses", cases)


# ---- a8 + a9: the real PPO agent, play_steps tail + train_epoch -------------------------------------------
def ppo_cfg(N, T, mb, units, variant, mini_epochs):
    if variant == "elu_normal_fixed":
        ack = dict(separate_value_mlp=True, mlp_kwargs=dict(units=units, act_type="ELU"), fixed_sigma=True, actor_dist_kwargs=dict(dist_type="normal"))
        optim_type, mgn = "Adam", 1.0
    else:  # DFlexAntPPO.yaml:14-26,59-66
        ack = dict(separate_value_mlp=True, mlp_kwargs=dict(units=units, norm_type="LayerNorm", act_type="SiLU"), fixed_sigma=False,
                   actor_dist_kwargs=dict(dist_type="dreamerv3_normal", minstd=0.1, maxstd=1.0))
        optim_type, mgn = "AdamW", 0.5
    return rs.to_attr(dict(
        rl_device="cpu", multi_gpu=False, env_render=False, run="train", seed=0, task=dict(env_autoresets=True),
        agent=dict(tracker_len=100, metrics_kwargs={}, print_every=-1, ckpt_every=-1,
                   network=dict(normalize_input=True, actor_critic="ActorCritic", actor_critic_kwargs=ack),
                   ppo=dict(multi_gpu=False, num_actors=N, reward_shaper=dict(fn="scale", scale=1.0), max_agent_steps=1e6, horizon_len=T, minibatch_size=mb,
                            normalize_value=True, value_bootstrap=True, clip_value_loss=False, normalize_advantage=True, gamma=0.99, tau=0.95,
                            entropy_coef=0.0 if variant == "elu_normal_fixed" else 0.001, e_clip=0.2, use_smooth_clamp=False, critic_coef=4,
                            bounds_loss_coef=0.0001, bounds_type="bound", mini_epochs=mini_epochs, optim_type=optim_type, optim_kwargs=dict(lr=5e-4, eps=1e-5),
                            lr_schedule="kl", kl_threshold=0.008, max_grad_norm=mgn, truncate_grads=True))))


def gen_ppo_e2e():
    cases = []
    for variant in ("elu_normal_fixed", "ln_silu_dreamer_learned"):
        N, T, O, A, mb, units, epochs = 16, 8, 11, 3, 32, [32, 16], 3
        torch.manual_seed(40)
        env = rs.SynthEnv(N, O, A, T, seed=41, done_p=0.08, timeout_p=0.05)
        cfg = ppo_cfg(N, T, mb, units, variant, epochs)
        with open(os.devnull, "w") as devnull:
            stdout, sys.stdout = sys.stdout, devnull
            try:
                agent = PPO(cfg, logdir=tempfile.mkdtemp(), env=env)
            finally:
                sys.stdout = stdout
        agent.obs = agent._convert_obs(env.reset())
        agent.dones = torch.zeros(N, dtype=torch.bool)
        agent.epoch = 0
        init_sd = clone_tree(agent.model.state_dict())
        rec = {}
        orig = agent.storage.compute_return

        def spy(last_values, gamma, tau):
            rec["last_values"] = last_values.clone()
            rec["storage"] = clone_tree({k: v for k, v in agent.storage.storage_dict.items() if k != "returns"})
            rec["value_rms_before"] = clone_tree(agent.value_rms.state_dict())
            return orig(last_values, gamma, tau)

        agent.storage.compute_return = spy
        agent.set_eval()
        agent.play_steps()
        rec["data_after_play"] = clone_tree(agent.storage.data_dict)
        rec["value_rms_after_play"] = clone_tree(agent.value_rms.state_dict())
        rec["obs_rms_before"] = clone_tree(agent.obs_rms.state_dict())
        agent.set_train()
        res = agent.train_epoch()
        cases.append(dict(
            variant=variant, shape=(T, N, O, A), minibatch_size=mb, units=units, mini_epochs=epochs, cfg=rs.to_plain(cfg),
            init_state_dict=init_sd, **rec,
            results={k: torch.stack(v) for k, v in res.items() if v[0].dim() == 0},
            results_mu_mean=torch.cat(res["mu"]).mean(0), results_sigma_mean=torch.cat(res["sigma"]).mean(0),
            final_state_dict=clone_tree(agent.model.state_dict()), obs_rms_after=clone_tree(agent.obs_rms.state_dict()),
            value_rms_after=clone_tree(agent.value_rms.state_dict()), data_after_train=clone_tree({k: agent.storage.data_dict[k] for k in ("mu", "sigma")}),
            last_lr=agent.last_lr, optim_state={i: clone_tree(s) for i, s in enumerate(agent.optim.state_dict()["state"].values())},
        ))
    save("ppo_e2e", cases)


# ---- 8(f)-2 agents: the real SAC / DDPG agents (explore_env + update_net), random draws recorded -----------------------
def offpolicy_cfg(algo, N, units, batch_size, mini_epochs, variant):
    if algo == "sac":
        mlp = dict(units=units, norm_type="LayerNorm", act_type="SiLU") if variant == "ln_silu" else dict(units=units, act_type="ELU")
        network = dict(normalize_input=True, encoder=None, actor="Actor", critic="EnsembleQ",
                       actor_kwargs=dict(mlp_kwargs=mlp, weight_init="orthogonal", tanh_policy=False, fixed_sigma=False,
                                         dist_kwargs=dict(dist_type="squashed_normal", minlogstd=-5.0, maxlogstd=2.0)),
                       critic_kwargs=dict(n_critics=2, mlp_kwargs=mlp, weight_init="orthogonal"))
        algo_cfg = dict(multi_gpu=False, num_actors=N, reward_shaper=dict(fn="scale", scale=0.5), max_agent_steps=1e6, horizon_len=4,
                        memory_size=1000, batch_size=batch_size, nstep=3, tau=0.05, gamma=0.99, warm_up=8, no_tgt_actor=True, handle_timeout=True,
                        update_actor_interval=1, update_targets_interval=1, mini_epochs=mini_epochs, init_alpha=1.0, alpha=None,
                        alpha_optim_kwargs=dict(lr=5e-3), backup_entropy=True, optim_type="AdamW", actor_optim_kwargs=dict(lr=5e-4),
                        critic_optim_kwargs=dict(lr=5e-4), max_grad_norm=0.5)
        if variant == "fixed":   # the other branches: sigma parameter, dreamerv3 std, no entropy back-up, raw
            # observations, a target actor, one-step returns, Adam, no time-out handling, a single critic
            network["normalize_input"] = False
            network["actor_kwargs"].update(fixed_sigma=True, dist_kwargs=dict(dist_type="dreamerv3_squashed_normal", minstd=0.1, maxstd=1.0))
            network["critic_kwargs"]["n_critics"] = 1
            # (a fixed temperature `alpha: 0.2` crashes the reference itself: sac.py:332 returns an unbound `alpha_loss`)
            algo_cfg.update(backup_entropy=False, no_tgt_actor=False, nstep=1, optim_type="Adam", handle_timeout=False,
                            update_targets_interval=2, target_entropy_scalar=0.5)
    else:  # AllegroHandDDPG.yaml (TD3-style intervals in the second variant)
        mlp = dict(units=units, act_type="ELU")
        td3 = variant in ("td3", "distl_td3")
        network = dict(normalize_input=True, encoder=None, actor="Actor", critic="EnsembleQ",
                       actor_kwargs=dict(tanh_policy=True, mlp_kwargs=mlp), critic_kwargs=dict(n_critics=2, mlp_kwargs=mlp))
        if variant.startswith("distl"):   # AllegroHandDDPG.yaml:34 (commented alternative): categorical critics on a fixed support
            network.update(critic="DistributionalEnsembleQ",
                           critic_kwargs=dict(v_min=-4.0, v_max=4.0, num_atoms=21 if td3 else 51, n_critics=2, mlp_kwargs=mlp))
        algo_cfg = dict(multi_gpu=False, num_actors=N, reward_shaper=dict(fn="scale", scale=0.5), max_agent_steps=1e6, horizon_len=4,
                        memory_size=1000, batch_size=batch_size, nstep=3, tau=0.05, gamma=0.99, warm_up=8, no_tgt_actor=not td3,
                        handle_timeout=True, update_actor_interval=2 if td3 else 1, update_targets_interval=2 if td3 else 1,
                        mini_epochs=mini_epochs, optim_type="AdamW", actor_optim_kwargs=dict(lr=5e-4), critic_optim_kwargs=dict(lr=5e-4),
                        max_grad_norm=0.5, exp_decay_rate=0.99,
                        noise=dict(type="fixed" if td3 else "mixed", decay=None, exp_decay_rate=0.99, lin_decay_iters=10000, std_max=0.8,
                                   std_min=0.05, tgt_pol_std=0.8, tgt_pol_noise_bound=0.2))
    return rs.to_attr(dict(rl_device="cpu", multi_gpu=False, env_render=False, run="train", seed=0, task=dict(env_autoresets=True),
                           agent={"tracker_len": 100, "metrics_kwargs": {}, "print_every": -1, "ckpt_every": -1, "network": network, algo: algo_cfg}))


class _Recorder:
    """Wraps torch's random entry points the agents use and records every draw, in call order per stream."""

    def __init__(self):
        import torch.distributions.normal as tdn
        self.tdn = tdn
        self.streams = dict(rand=[], randint=[], std_normal=[], normal=[])
        self.orig = dict(rand=torch.rand, randint=torch.randint, std_normal=tdn._standard_normal, normal=torch.normal)

    def __enter__(self):
        def wrap(name):
            def f(*a, **kw):
                if kw.get("device") == "cuda":   # DDPG samples with sample_batch's default device='cuda' (ddpg.py:244): no GPU here
                    kw["device"] = "cpu"
                out = self.orig[name](*a, **kw)
                self.streams[name].append(out.detach().clone())
                return out
            return f
        torch.rand, torch.randint, torch.normal = wrap("rand"), wrap("randint"), wrap("normal")
        self.tdn._standard_normal = wrap("std_normal")
        return self

    def __exit__(self, *exc):
        torch.rand, torch.randint, torch.normal = self.orig["rand"], self.orig["randint"], self.orig["normal"]
        self.tdn._standard_normal = self.orig["std_normal"]


def gen_offpolicy_e2e(spec=(("sac", "ln_silu"), ("sac", "elu"), ("ddpg", "mixed"), ("ddpg", "td3"), ("sac", "fixed")), name="offpolicy_e2e"):
    from mineral.agents.ddpg.ddpg import DDPG
    from mineral.agents.sac.sac import SAC
    cases = []
    for algo, variant in spec:
        N, O, A, units, B, mini_epochs, epochs = 16, 11, 3, [32, 16], 64, 3, 2
        torch.manual_seed(50)
        env = rs.SynthEnv(N, O, A, 64, seed=51, done_p=0.08, timeout_p=0.05)
        cfg = offpolicy_cfg(algo, N, units, B, mini_epochs, variant)
        Cls = SAC if algo == "sac" else DDPG
        with open(os.devnull, "w") as devnull:
            stdout, sys.stdout = sys.stdout, devnull
            try:
                agent = Cls(cfg, logdir=tempfile.mkdtemp(), env=env)
            finally:
                sys.stdout = stdout
        acfg = cfg.agent[algo]
        if algo == "ddpg":   # ddpg.py:244 samples with sample_batch's default device='cuda'; this container has no GPU
            sample_cpu = agent.memory.sample_batch
            agent.memory.sample_batch = lambda bs, device="cpu": sample_cpu(bs, device="cpu")
        init = dict(actor=clone_tree(agent.actor.state_dict()), critic=clone_tree(agent.critic.state_dict()))
        per_epoch = []
        with _Recorder() as rec:
            agent.obs = agent._convert_obs(env.reset())
            agent.dones = torch.ones(N, dtype=torch.bool)
            agent.epoch = 0
            agent.set_eval()
            traj, steps = agent.explore_env(env, acfg.warm_up, random=True)
            agent.memory.add_to_buffer(traj)
            warm = dict(trajectory=clone_tree(traj), next_p=agent.memory.next_p, cur_capacity=agent.memory.cur_capacity)
            for _ in range(epochs):
                agent.epoch += 1
                agent.set_eval()
                traj, steps = agent.explore_env(env, acfg.horizon_len, sample=True)
                agent.memory.add_to_buffer(traj)
                agent.set_train()
                res = agent.update_net(agent.memory)
                per_epoch.append(dict(
                    trajectory=clone_tree(traj), results={k: torch.stack([x.detach() for x in v]) for k, v in res.items()},
                    actor=clone_tree(agent.actor.state_dict()), critic=clone_tree(agent.critic.state_dict()),
                    critic_target=clone_tree(agent.critic_target.state_dict()),
                    actor_target=clone_tree(agent.actor_target.state_dict()),
                    log_alpha=agent.log_alpha.detach().clone() if algo == "sac" else None,
                    obs_rms=clone_tree(agent.obs_rms.state_dict()) if agent.obs_rms is not None else {}, next_p=agent.memory.next_p, cur_capacity=agent.memory.cur_capacity))
            streams = {k: [x.clone() for x in v] for k, v in rec.streams.items()}
        cases.append(dict(algo=algo, variant=variant, shape=(N, O, A), cfg=rs.to_plain(cfg), init=init, warm=warm, epochs=per_epoch,
                          streams=streams, env=dict(horizon=64, seed=51, done_p=0.08, timeout_p=0.05)))
    save(name, cases)


def gen_offpolicy_distl():
    """DDPG with the categorical critic (DistributionalEnsembleQ, ddpg.py:267-290): mixed-noise and TD3-style variants."""
    gen_offpolicy_e2e(spec=(("ddpg", "distl"), ("ddpg", "distl_td3")), name="offpolicy_distl")


# ---- SAC policy head: Dist + SquashedNormal rsample / log_prob and autograd's gradients (the oracle restatement is pinned on these)
def gen_sac_policy():
    from mineral.nets.dist import Dist
    cases = []
    g = torch.Generator().manual_seed(60)
    for dist_kwargs in (dict(dist_type="squashed_normal", minlogstd=-5.0, maxlogstd=2.0), dict(dist_type="squashed_normal"),
                        dict(dist_type="dreamerv3_squashed_normal", minstd=0.1, maxstd=1.0)):
        for B, A in ((64, 8), (33, 21)):
            mu = (torch.randn(B, A, generator=g) * 1.5).requires_grad_()
            ls = (torch.randn(B, A, generator=g) * 2.5).requires_grad_()   # wide enough to hit both clamp bounds
            eps = torch.randn(B, A, generator=g)
            g_act = torch.randn(B, A, generator=g) / B
            alpha = 0.37
            dist = Dist(**dist_kwargs)
            _, sigma, distr = dist(mu, ls)
            import torch.distributions.normal as tdn
            orig = tdn._standard_normal
            tdn._standard_normal = lambda shape, dtype, device: eps.clone()
            try:
                actions = distr.rsample()
            finally:
                tdn._standard_normal = orig
            log_prob = distr.log_prob(actions).sum(-1, keepdim=True)
            objective = (alpha * log_prob).mean() + (g_act * actions).sum()
            objective.backward()
            cases.append(dict(dist_kwargs=dist_kwargs, mu=mu.detach().clone(), logstd=ls.detach().clone(), eps=eps, g_act=g_act, alpha=alpha,
                              sigma=sigma.detach().clone(), actions=actions.detach().clone(), log_prob=log_prob.detach().clone(),
                              g_mu=mu.grad.clone(), g_logstd=ls.grad.clone()))
    save("sac_policy", cases)


# ---- DDPG action noise (mineral/agents/ddpg/noise.py:22-55) and the actor objectives' seeds (sac.py:320-331, ddpg.py:305-307) --------
def gen_offpolicy_small():
    from mineral.agents.ddpg import noise as noise_mod
    g = torch.Generator().manual_seed(80)
    out = dict(noise=[], objective=[])
    B, A = 96, 5
    mu = torch.randn(B, A, generator=g) * 2
    for kind, kw in (("fixed", dict(std=0.8, out_bounds=[-1.0, 1.0])),
                     ("fixed", dict(std=0.8, noise_bounds=[-0.2, 0.2], out_bounds=[-1.0, 1.0])),
                     ("mixed", dict(std_min=0.05, std_max=0.8, out_bounds=[-1.0, 1.0]))):
        drawn = {}
        real_normal = torch.normal

        def spy(mean, std, *a, **k):
            drawn["noise"] = real_normal(mean, std, *a, **k)
            return drawn["noise"]

        noise_mod.torch.normal = spy   # (noise.py's `torch` IS torch: restored right below)
        try:
            fn = noise_mod.add_normal_noise if kind == "fixed" else noise_mod.add_mixed_normal_noise
            res = fn(mu.tanh(), **kw)
        finally:
            torch.normal = real_normal
        out["noise"].append(dict(kind=kind, kwargs=kw, mu=mu.clone(), noise=drawn["noise"].clone(), actions=res.clone()))
    # actor objectives through the reference's own expressions + autograd
    for use_lp in (True, False):
        qs = [(5 * torch.randn(B, 1, generator=g)).requires_grad_() for _ in range(2)]
        with torch.no_grad():
            qs[1][:10] = qs[0][:10]   # ties
        lp = torch.randn(B, 1, generator=g) - 3
        alpha, te = 0.37, -float(A)
        q_min = torch.min(torch.stack(qs), dim=0).values                      # EnsembleQ.get_q_min (ddpg/models.py:108-110)
        actor_loss = (alpha * lp - q_min).mean() if use_lp else -q_min.mean()  # sac.py:323 / ddpg.py:306
        actor_loss.backward()
        entropy = -lp
        alpha_loss = (alpha * (entropy - te)).mean()                           # sac.py:329-330 (value; d/dlog_alpha equals it)
        out["objective"].append(dict(use_lp=use_lp, q=[q.detach().clone() for q in qs], log_prob=lp, alpha=alpha, target_entropy=te,
                                     actor_loss=actor_loss.detach().clone(), entropy=entropy.mean(), alpha_loss=alpha_loss,
                                     seeds=[q.grad.clone() for q in qs]))
    save("offpolicy_small", out)


# ---- f4: a checkpoint WRITTEN BY THE REFERENCE (ppo.py:416-427) and what the reference does after it -------------------
def gen_ppo_ckpt():
    """Reference PPO: rollout + update #1 -> agent.save() -> rollout + update #2.  The fixture holds the checkpoint file's
    content, rollout #2's storage and update #2's results: loading the checkpoint into another implementation and feeding
    it rollout #2 must reproduce update #2 (model, Adam moments / step / lr, obs and value statistics all matter)."""
    cases = []
    for variant in ("elu_normal_fixed", "ln_silu_dreamer_learned"):
        N, T, O, A, mb, units, epochs = 16, 8, 11, 3, 32, [32, 16], 2
        torch.manual_seed(70)
        env = rs.SynthEnv(N, O, A, 2 * T + 2, seed=71, done_p=0.08, timeout_p=0.05)
        cfg = ppo_cfg(N, T, mb, units, variant, epochs)
        with open(os.devnull, "w") as devnull:
            stdout, sys.stdout = sys.stdout, devnull
            try:
                agent = PPO(cfg, logdir=tempfile.mkdtemp(), env=env)
            finally:
                sys.stdout = stdout
        agent.obs = agent._convert_obs(env.reset())
        agent.dones = torch.zeros(N, dtype=torch.bool)
        agent.epoch, agent.mini_epoch, agent.agent_steps = 0, -1, 0
        agent.set_eval(); agent.play_steps(); agent.agent_steps += agent.batch_size
        agent.set_train(); agent.train_epoch(); agent.storage.data_dict = None
        agent.epoch = 1
        f = os.path.join(tempfile.mkdtemp(), "ref.pth")
        agent.save(f)
        ckpt = torch.load(f, weights_only=False)
        rec = {}
        orig = agent.storage.compute_return

        def spy(last_values, gamma, tau):
            rec["last_values"] = last_values.clone()
            rec["storage"] = clone_tree({k: v for k, v in agent.storage.storage_dict.items() if k != "returns"})
            return orig(last_values, gamma, tau)

        agent.storage.compute_return = spy
        agent.set_eval(); agent.play_steps()
        agent.set_train()
        res = agent.train_epoch()
        cases.append(dict(variant=variant, shape=(T, N, O, A), cfg=rs.to_plain(cfg), ckpt=clone_tree(ckpt), **rec,
                          results={k: torch.stack(v) for k, v in res.items() if v[0].dim() == 0},
                          final_state_dict=clone_tree(agent.model.state_dict()), last_lr=agent.last_lr,
                          optim_after=clone_tree(agent.optim.state_dict()), obs_rms_after=clone_tree(agent.obs_rms.state_dict()),
                          value_rms_after=clone_tree(agent.value_rms.state_dict())))
    save("ppo_ckpt", cases)


# ---- a17 + rollout bookkeeping: the REAL SHAC.compute_actor_loss on a synthetic differentiable env -----------------------
class SynthDiffEnv:
    """Differentiable toy simulator with the interface shac.py:38-43,345,645,688,714 consumes: linear dynamics in torch,
    pre-drawn early terminations, time-outs at `max_episode_length`, `info['obs_before_reset']`, one NaN terminal observation."""

    def __init__(self, N, O, A, max_episode_length, seed, T):
        g = torch.Generator().manual_seed(seed)
        self.num_envs, self.num_obs, self.num_actions, self.max_episode_length = N, O, A, max_episode_length
        self.observation_space, self.action_space = rs._Space((O,)), rs._Space((A,))
        self.scale = torch.empty(O).uniform_(0.1, 4.0, generator=g)
        self.shift = torch.empty(O).uniform_(-2.0, 2.0, generator=g)
        self.M = 0.3 * torch.randn(A, O, generator=g)
        self.fresh = torch.randn(T + 2, N, O, generator=g) * self.scale + self.shift
        self.noise = 0.05 * torch.randn(T + 2, N, O, generator=g)
        self.term = torch.rand(T + 2, N, generator=g) < 0.12
        self.state = self.fresh[0].clone()
        self.progress = torch.zeros(N, dtype=torch.long)
        self.t = 0
        self.log = []

    def clear_grad(self):
        self.state = self.state.detach()

    def reset(self):
        return self.state.clone()

    def initialize_trajectory(self):
        self.state = self.state.detach()
        return self.state.clone()

    def step(self, actions):
        self.t += 1
        t = self.t
        self.state = 0.9 * self.state + actions @ self.M + self.noise[t]
        rew = -0.01 * (self.state ** 2).mean(1) + 0.1 * actions.sum(1)
        self.progress += 1
        done = self.term[t] | (self.progress >= self.max_episode_length)
        before = self.state.detach().clone()
        if t == 3 and done.any():   # a non-finite terminal observation (shac.py:720-731)
            before[done.nonzero()[0, 0], 0] = float("nan")
        self.state = torch.where(done[:, None], self.fresh[t], self.state)
        self.progress[done] = 0
        obs = self.state.clone()
        self.log.append(dict(actions=actions.detach().clone(), obs=obs.detach().clone(), rew=rew.detach().clone(), done=done.clone(),
                             obs_before_reset=before.clone()))
        return obs, rew, done, {"obs_before_reset": before}


def gen_shac_rollout():
    cases = []
    for variant in ("legacy_tanh_clamp", "f64_ensemble"):
        N, T, O, A, max_len = 12, 8, 7, 3, 5
        legacy = variant == "legacy_tanh_clamp"
        mlp = dict(units=[16, 8], norm_type="LayerNorm", act_type="ELU")
        critic = dict(critic="Critic", critic_kwargs=dict(mlp_kwargs=mlp)) if legacy else \
            dict(critic="EnsembleCritic", critic_kwargs=dict(n_critics=2, mlp_kwargs=mlp))
        cfg = rs.to_attr(dict(
            rl_device="cpu", multi_gpu=False, env_render=False, run="train", seed=0, task=dict(env_autoresets=True),
            agent=dict(tracker_len=100, metrics_kwargs={}, print_every=-1, ckpt_every=-1,
                       network=dict(normalize_input=True, encoder=None, tanh_clamp=legacy, actor="Actor",
                                    actor_kwargs=dict(mlp_kwargs=dict(units=[16, 8], norm_type="LayerNorm", act_type="ELU")), **critic),
                       shac=dict(multi_gpu=False, num_actors=N, reward_shaper=dict(fn="scale", scale=0.5), max_epochs=10, max_agent_steps=1e6,
                                 horizon_len=T, num_critic_batches=4, gamma=0.99, critic_method="td-lambda", **{"lambda": 0.95},
                                 critic_iterations=2, normalize_ret=False, optim_type="Adam", actor_optim_kwargs=dict(lr=2e-3, betas=[0.7, 0.95]),
                                 critic_optim_kwargs=dict(lr=2e-3, betas=[0.7, 0.95]), lr_schedule="linear", target_critic_alpha=0.2,
                                 no_target_critic=False, max_grad_norm=1.0, truncate_grads=True, with_logprobs=not legacy))))
        torch.manual_seed(90)
        env = SynthDiffEnv(N, O, A, max_len, seed=91, T=2 * T)
        with open(os.devnull, "w") as devnull:
            stdout, sys.stdout = sys.stdout, devnull
            try:
                agent = SHAC(cfg, logdir=tempfile.mkdtemp(), env=env)
                rollouts = []
                for _ in range(2):   # two consecutive rollouts: the second starts from non-trivial statistics
                    env.log = []
                    rms_before = clone_tree(agent.obs_rms.state_dict())
                    first_obs = env.state.detach().clone()
                    agent.compute_actor_loss()
                    rollouts.append(dict(
                        first_obs=first_obs, log=env.log, obs_rms_before=rms_before, obs_rms_after=clone_tree(agent.obs_rms.state_dict()),
                        obs_buf=clone_tree(agent.obs_buf), action_buf=agent.action_buf.detach().clone(), rew_buf=agent.rew_buf.clone(),
                        done_mask=agent.done_mask.clone(), next_values=agent.next_values.detach().clone(),
                        avg_next_values=agent.avg_next_values.detach().clone(), episode_lengths=agent.episode_lengths.clone(),
                        episode_rewards=agent.episode_rewards.clone(), num_episodes=int(agent.num_episodes), agent_steps=agent.agent_steps))
            finally:
                sys.stdout = stdout
        cases.append(dict(variant=variant, shape=(T, N, O, A), max_episode_length=max_len, cfg=rs.to_plain(cfg),
                          critic_target=clone_tree(agent.critic_target.state_dict()), critic=clone_tree(agent.critic.state_dict()),
                          actor=clone_tree(agent.actor.state_dict()), rollouts=rollouts))
    save("shac_rollout", cases)


if __name__ == "__main__":
    gen_shac_rollout() if "shac_rollout" in sys.argv[1:] else None
    gen_ppo_ckpt() if "ppo_ckpt" in sys.argv[1:] else None
    gen_offpolicy_distl() if "offpolicy_distl" in sys.argv[1:] else None
    if len(sys.argv) > 1:
        sys.exit(0)
    gen_experience()
    gen_rms()
    gen_tdlambda()
    gen_critic_dataset()
    gen_shac_critic()
    gen_offpolicy()
    gen_replay()
    gen_nstep()
    gen_losses()
    gen_ppo_e2e()
    gen_offpolicy_e2e()
    gen_sac_policy()
    gen_offpolicy_small()
    gen_shac_rollout()
    gen_ppo_ckpt()
    gen_offpolicy_distl()
