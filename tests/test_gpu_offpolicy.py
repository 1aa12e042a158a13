"""GPU parity of the SAC / DDPG agents (SURVEY.md section 8(b) agent plug-ins, 8(f) row 2): the row-wise kernels of
csrc/offpolicy.cu against the oracle and the reference-generated goldens, then the whole agents (explore_env + update_net,
every network pass on the tcgen05 / SIMT kernels) against the REAL reference agents' recorded runs with their own random
draws replayed (tests/golden/offpolicy_e2e.pt, oracle/make_golden.py:gen_offpolicy_e2e)."""

import pytest
import torch

from conftest import load_golden, scale_close
from oracle import mineral_oracle as oc

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def cu(x):
    if isinstance(x, dict):
        return {k: cu(v) for k, v in x.items()}
    return x.to(DEV).contiguous() if torch.is_tensor(x) else x


def eq(a, b, msg=""):
    a = a.cpu()
    assert a.shape == b.shape and a.dtype == b.dtype, f"{msg}: {a.shape}/{a.dtype} vs {b.shape}/{b.dtype}"
    assert torch.equal(a, b), f"{msg}: max|diff| = {(a.double() - b.double()).abs().max().item():.3e}"


def _squash_cfg(kw):
    from mineral_b200 import ops
    return ops.make_squash_cfg(kw.get("dist_type"), kw.get("minlogstd"), kw.get("maxlogstd"), kw.get("minstd", 1.0), kw.get("maxstd", 1.0))


def test_squashed_normal_kernels_match_reference_golden_and_oracle():
    from mineral_b200 import ops

    def run(mu, ls, eps, g_act, alpha, kw, log_alpha_on_device):
        B, A = mu.shape
        cfg = _squash_cfg(kw)
        ld = A + 5
        act_buf = torch.zeros(B, ld, device=DEV)           # strided destination: a column block of a wider buffer
        lp = torch.empty(B, device=DEV)
        ops.squashed_sample(cu(mu), cu(ls), cu(eps), cfg, act_buf[:, 2:2 + A], lp)
        g_buf = torch.zeros(B, ld, device=DEV)
        g_buf[:, 3:3 + A] = cu(g_act)
        g_mu, g_ls = torch.empty(B, A, device=DEV), torch.empty(B, A, device=DEV)
        la = torch.tensor([alpha], device=DEV).log() if log_alpha_on_device else None
        ops.squashed_bwd(cu(mu), cu(ls), cu(eps), g_buf[:, 3:3 + A], la, alpha, cfg, g_mu, g_ls)
        assert act_buf[:, :2].abs().max().item() == 0 and act_buf[:, 2 + A:].abs().max().item() == 0, "wrote outside its block"
        return act_buf[:, 2:2 + A].contiguous(), lp.view(B, 1), g_mu, g_ls

    for c in load_golden("sac_policy"):
        kw = c["dist_kwargs"]
        B = c["mu"].shape[0]
        act, lp, g_mu, g_ls = run(c["mu"], c["logstd"], c["eps"], c["g_act"], c["alpha"], kw, False)
        scale_close(act, c["actions"], 2e-6, "actions")
        scale_close(lp, c["log_prob"], 2e-6, "log_prob")
        sigma, _ = oc.squashed_sigma(c["logstd"], **kw)
        floor = (c["alpha"] / B * c["eps"] / sigma).abs().max().item() * 2.0 ** -24   # see tests/test_oracle_golden.py
        for name, mine, ref, extra in (("g_mu", g_mu, c["g_mu"], 2 * floor), ("g_logstd", g_ls, c["g_logstd"], 2 * floor * sigma.max().item())):
            err = (mine.cpu() - ref).abs().max().item()
            assert err <= 1e-5 * ref.abs().max().item() + extra, f"{name}: {err:.3e}"
        o_mu, o_ls = oc.squashed_bwd(c["mu"], c["logstd"], c["eps"], c["g_act"], c["alpha"] / B, **kw)
        scale_close(g_mu, o_mu, 2e-6, "g_mu vs oracle")
        scale_close(g_ls, o_ls, 2e-6, "g_logstd vs oracle")
    # cfg4 batch (8192 x 8), temperature read from the device
    g = torch.Generator().manual_seed(70)
    B, A = 8192, 8
    mu, ls, eps, g_act = torch.randn(B, A, generator=g), torch.randn(B, A, generator=g) - 1, torch.randn(B, A, generator=g), torch.randn(B, A, generator=g) / B
    kw = dict(dist_type="squashed_normal", minlogstd=-5.0, maxlogstd=2.0)
    act, lp, g_mu, g_ls = run(mu, ls, eps, g_act, 0.2, kw, True)
    o_act, o_lp = oc.squashed_sample(mu, ls, eps, **kw)
    o_mu, o_ls = oc.squashed_bwd(mu, ls, eps, g_act, 0.2 / B, **kw)
    scale_close(act, o_act, 2e-6, "actions 8192"), scale_close(lp, o_lp, 2e-6, "log_prob 8192")
    scale_close(g_mu, o_mu, 3e-6, "g_mu 8192"), scale_close(g_ls, o_ls, 3e-6, "g_logstd 8192")
    # the action gradient given as two addends (one per critic tower)
    from mineral_b200 import ops as _ops
    part = torch.randn(B, A, generator=g) / B
    g2_mu, g2_ls = torch.empty(B, A, device=DEV), torch.empty(B, A, device=DEV)
    _ops.squashed_bwd(cu(mu), cu(ls), cu(eps), cu(g_act - part), torch.tensor([0.2], device=DEV).log(), 0.0, _squash_cfg(kw), g2_mu, g2_ls,
                      g_actions2=cu(part))
    scale_close(g2_mu, o_mu, 3e-6, "g_mu, two towers"), scale_close(g2_ls, o_ls, 3e-6, "g_logstd, two towers")
    # fixed sigma: the (A,) parameter is broadcast; deterministic branch returns the un-squashed mean
    ls1 = torch.randn(A, generator=g) * 0.3
    cfg = _squash_cfg(kw)
    a2, lp2 = torch.empty(B, A, device=DEV), torch.empty(B, device=DEV)
    ops.squashed_sample(cu(mu), cu(ls1), cu(eps), cfg, a2, lp2)
    o_a2, o_lp2 = oc.squashed_sample(mu, ls1.expand(B, A), eps, **kw)
    scale_close(a2, o_a2, 2e-6, "broadcast actions"), scale_close(lp2.view(B, 1), o_lp2, 2e-6, "broadcast log_prob")
    a3 = torch.empty(B, A, device=DEV)
    ops.squashed_sample(cu(mu), None, None, cfg, a3, None)
    eq(a3, mu, "deterministic actions are the raw mean (sac.py:129-130)")


def test_offpolicy_row_kernels_vs_oracle():
    from mineral_b200 import ops

    g = torch.Generator().manual_seed(71)
    B, A, O = 8192, 8, 37
    # DDPG / TD3 actor output
    mu, noise, ga = torch.randn(B, A, generator=g) * 2, torch.randn(B, A, generator=g) * 0.8, torch.randn(B, A, generator=g)
    for nz, nb, ob in ((None, None, None), (noise, None, (-1.0, 1.0)), (noise, (-0.2, 0.2), (-1.0, 1.0))):
        out = torch.empty(B, A, device=DEV)
        ops.tanh_action(cu(mu), None if nz is None else cu(nz), out, nb, ob)
        scale_close(out, oc.tanh_action(mu, nz, nb, ob), 1e-6, f"tanh_action {nb} {ob}")
    g_mu = torch.empty(B, A, device=DEV)
    ops.tanh_action_bwd(cu(mu), cu(ga), g_mu)
    scale_close(g_mu, ga * (1 - mu.tanh() ** 2), 2e-6, "tanh_action_bwd")
    gb = torch.randn(B, A, generator=g)   # two addends (one input gradient per critic tower)
    ops.tanh_action_bwd(cu(mu), cu(ga), g_mu, g_actions2=cu(gb))
    scale_close(g_mu, (ga + gb) * (1 - mu.tanh() ** 2), 2e-6, "tanh_action_bwd, two towers")
    # actor objective + seeds (SAC with a device temperature, DDPG without log-prob)
    q = [5 * torch.randn(B, generator=g) for _ in range(2)]
    q[1][:100] = q[0][:100]                              # ties: torch.min picks the first critic
    lp = torch.randn(B, generator=g) - 3
    alpha, te = 0.37, -8.0
    for use_lp in (True, False):
        seeds = [torch.empty(B, device=DEV) for _ in q]
        out4, la_grad = torch.zeros(4, device=DEV), torch.zeros(1, device=DEV)
        la = torch.tensor([alpha], device=DEV).log()
        ops.actor_seed([cu(x) for x in q], seeds, cu(lp) if use_lp else None, la if use_lp else None, 0.0, te, out4, la_grad if use_lp else None)
        a_loss, ent, al_loss, o_seeds = oc.actor_objective(q, lp if use_lp else None, alpha if use_lp else 0.0, te)
        scale_close(out4[0], a_loss, 1e-5, "actor_loss")
        for k in range(2):
            eq(seeds[k], o_seeds[k], f"seed {k}")
        if use_lp:
            scale_close(out4[1], ent, 1e-5, "entropy"), scale_close(out4[2], al_loss, 1e-5, "alpha_loss")
            scale_close(out4[3], torch.tensor(alpha), 1e-6, "alpha"), scale_close(la_grad[0], al_loss, 1e-5, "d alpha_loss / d log_alpha")
    # td target with a device temperature + its statistics
    rew, dn = torch.randn(B, generator=g), (torch.rand(B, generator=g) < 0.1).float()
    out, stats = torch.empty(B, device=DEV), torch.empty(4, device=DEV)
    ops.td_target_stats([cu(x) for x in q], cu(lp), cu(rew), cu(dn), torch.tensor([alpha], device=DEV).log(), 0.0, 0.99 ** 3, out, stats)
    ref = oc.td_target([x.view(B, 1) for x in q], rew.view(B, 1), dn.view(B, 1), 0.99, 3, alpha, lp.view(B, 1)).view(B)
    scale_close(out, ref, 1e-6, "td target (device alpha)")
    scale_close(stats, torch.stack([ref.mean(), ref.std(), ref.max(), ref.min()]), 1e-5, "target stats")
    ops.td_target_stats([cu(x) for x in q], None, cu(rew), cu(dn), None, 0.0, 0.99 ** 3, out, None)
    eq(out, oc.td_target([x.view(B, 1) for x in q], rew.view(B, 1), dn.view(B, 1), 0.99, 3).view(B), "ddpg td target")
    # sum-over-critics MSE
    preds = [torch.randn(B, generator=g) for _ in range(2)]
    tgt = torch.randn(B, generator=g)
    grads, loss = [torch.empty(B, device=DEV) for _ in preds], torch.zeros(1, device=DEV)
    ws = torch.zeros(ops.lib.mnr_critic_mse_workspace_bytes(B), dtype=torch.uint8, device=DEV)
    ops.critic_mse_sum_fwd_bwd([cu(p) for p in preds], cu(tgt), grads, loss, ws)
    scale_close(loss[0], sum(((p - tgt) ** 2).mean() for p in preds), 1e-6, "mse sum")
    for p, g_ in zip(preds, grads):
        scale_close(g_, 2 * (p - tgt) / B, 1e-6, "mse grad")
    # trajectory store: (N, T, .) slices, dones masked by time-outs, rewards scaled
    N, T = 256, 4
    tr = dict(obs=torch.zeros(N, T, O, device=DEV), nobs=torch.zeros(N, T, O, device=DEV), act=torch.zeros(N, T, A, device=DEV),
              rew=torch.zeros(N, T, device=DEV), done=torch.zeros(N, T, device=DEV))
    ref = {k: v.cpu().clone() for k, v in tr.items()}
    for t in range(T):
        o, no, a = torch.randn(N, O, generator=g), torch.randn(N, O, generator=g), torch.randn(N, A, generator=g)
        r, d, to = torch.randn(N, generator=g), torch.rand(N, generator=g) < 0.3, torch.rand(N, generator=g) < 0.3
        ops.traj_store_step(cu(o), cu(no), cu(a), cu(r), cu(d), cu(to) if t % 2 == 0 else None, t, 0.01, tr["obs"], tr["nobs"], tr["act"], tr["rew"], tr["done"])
        ref["obs"][:, t], ref["nobs"][:, t], ref["act"][:, t] = o, no, a
        ref["rew"][:, t] = r * 0.01
        ref["done"][:, t] = (d * (~to) if t % 2 == 0 else d).float()
    for k in tr:
        eq(tr[k], ref[k], f"traj {k}")
    # strided gather destination (replay sample straight into the padded critic input)
    from mineral_b200.buffers import ReplayBuffer
    from mineral_b200.common.normalizers import RunningMeanStd
    buf = ReplayBuffer({"obs": (O,)}, A, capacity=1000, device=DEV)
    rows = 700
    traj = ({"obs": torch.randn(rows, O, generator=g)}, torch.randn(rows, A, generator=g), torch.randn(rows, 1, generator=g),
            {"obs": torch.randn(rows, O, generator=g)}, torch.rand(rows, 1, generator=g) < 0.2)
    buf.add_to_buffer(tuple(cu(x) for x in traj))
    rms = torch.nn.ModuleDict({"obs": RunningMeanStd((O,), eps=1e-4, with_clamp=True, initial_count="eps")}).to(DEV)
    rms["obs"].update(cu(traj[0]["obs"]))
    idx = torch.randint(rows, (512,), generator=g)
    plain = buf.sample_batch(512, device=DEV, indices=cu(idx), obs_rms=rms)
    ld = (O + A + 3) // 4 * 4
    xc, xn = torch.zeros(512, ld, device=DEV), torch.zeros(512, ld, device=DEV)
    out = dict(obs={"obs": xc[:, :O]}, next_obs={"obs": xn[:, :O]}, action=xc[:, O:O + A])
    buf.sample_batch(512, device=DEV, indices=cu(idx), obs_rms=rms, out=out)
    eq(xc[:, :O].contiguous(), plain[0]["obs"].cpu(), "strided obs"), eq(xc[:, O:O + A].contiguous(), plain[1].cpu(), "strided action")
    eq(xn[:, :O].contiguous(), plain[3]["obs"].cpu(), "strided next_obs")
    assert xc[:, O + A:].abs().max().item() == 0 and xn[:, O:].abs().max().item() == 0


class _Replay:
    """Feeds an agent's random hooks from the reference's recorded draws, in order, checking shapes."""

    def __init__(self, streams):
        self.s = {k: list(v) for k, v in streams.items()}

    def pop(self, name, shape=None):
        t = self.s[name].pop(0)
        if shape is not None:
            assert tuple(t.shape) == tuple(shape), f"{name}: recorded draw {tuple(t.shape)} vs requested {tuple(shape)}"
        return cu(t)


@pytest.mark.parametrize("cuda_graph", [False, True])
@pytest.mark.parametrize("case_index", [0, 1])
def test_ddpg_categorical_critic_matches_reference_run(case_index, cuda_graph):
    """DDPG with `DistributionalEnsembleQ` (mineral/agents/ddpg/models.py:117-170, ddpg.py:267-290; 51 atoms with mixed noise,
    21 atoms TD3-style): softmax + C51 projection + ensemble min, BCE regression of every critic, actor through the minimal
    critic's expected value - against the real reference agent's recorded run (golden offpolicy_distl.pt)."""
    test_offpolicy_agents_match_reference_run(case_index, cuda_graph, golden="offpolicy_distl")


@pytest.mark.parametrize("cuda_graph", [False, True])
@pytest.mark.parametrize("case_index", [0, 1, 2, 3, 4])
def test_offpolicy_agents_match_reference_run(case_index, cuda_graph, golden="offpolicy_e2e"):
    """SAC (LayerNorm+SiLU and ELU towers; case 4: sigma parameter, dreamerv3 std, single critic, target actor, raw observations,
    one-step returns, Adam) and DDPG (mixed exploration noise; TD3 with target actor, delayed updates) through
    warm-up, two explore_env + update_net epochs: trajectories, every logged scalar and every parameter against the real
    reference agents' run on the same env streams and random draws."""
    from mineral_b200.agents import DDPG, SAC
    from oracle.reference_shim import SynthEnv, to_attr

    c = load_golden(golden)[case_index]
    algo = c["algo"]
    N, O, A = c["shape"]
    cfg = to_attr(c["cfg"])
    cfg.rl_device = DEV
    cfg.agent[algo]["cuda_graph"] = cuda_graph   # True: update_net warms up eagerly once, then replays as ONE CUDA graph
    cpu_env = SynthEnv(N, O, A, c["env"]["horizon"], seed=c["env"]["seed"], done_p=c["env"]["done_p"], timeout_p=c["env"]["timeout_p"])

    class OnDevice:
        observation_space, action_space = cpu_env.observation_space, cpu_env.action_space

        def reset(self):
            return cu(cpu_env.reset())

        def step(self, actions):
            o, r, d, info = cpu_env.step(actions)
            return cu(o), cu(r), cu(d), cu(info)

    env = OnDevice()
    agent = (SAC if algo == "sac" else DDPG)(cfg, env=env)
    acfg = cfg.agent[algo]
    agent.actor.load_state_dict(c["init"]["actor"])
    agent.critic.load_state_dict(c["init"]["critic"])
    agent.critic_target.load_state_dict(c["init"]["critic"])
    if not acfg.no_tgt_actor:
        agent.actor_target.load_state_dict(c["init"]["actor"])
    rp = _Replay(c["streams"])
    agent._rand = lambda shape: rp.pop("rand", shape)
    agent._sample_indices = lambda memory, bs: rp.pop("randint", (bs,))
    if algo == "sac":
        agent._randn = lambda shape: rp.pop("std_normal", shape)
    else:  # torch.normal draws are recorded already scaled by their std (noise.py:30-33,48-49)
        agent._exploration_noise = lambda rows: rp.pop("normal", (rows, A))
        if not cuda_graph:   # graph mode pre-draws the target-smoothing noise of all mini-epochs through _draw_noise
            agent._target_noise = lambda rows: rp.pop("normal", (rows, A))

    B = acfg.batch_size

    def draw_indices(memory, buf):   # graph mode draws all mini-epochs of one update_net call up front
        for i in range(buf.shape[0]):
            buf[i].copy_(rp.pop("randint", (B,)))

    def draw_noise(buf):
        for i in range(buf.shape[0]):
            for j in range(buf.shape[1]):
                buf[i, j].copy_(rp.pop("std_normal" if algo == "sac" else "normal", (B, A)))

    agent._draw_indices, agent._draw_noise = draw_indices, draw_noise

    def check_traj(got, ref, tag):
        eq(got[0]["obs"], ref[0]["obs"], f"{tag} obs")
        scale_close(got[1], ref[1], 1e-5, f"{tag} actions")
        scale_close(got[2], ref[2], 1e-6, f"{tag} rewards")
        eq(got[3]["obs"], ref[3]["obs"], f"{tag} next_obs")
        eq(got[4].float(), ref[4].float(), f"{tag} dones")

    agent.obs = agent._convert_obs(env.reset())
    agent.dones = torch.ones(N, dtype=torch.bool, device=DEV)
    agent.epoch = 0
    agent.set_eval()
    traj, steps = agent.explore_env(env, acfg.warm_up, random=True)
    agent.memory.add_to_buffer(traj)
    check_traj(traj, c["warm"]["trajectory"], "warm-up")
    assert (agent.memory.next_p, agent.memory.cur_capacity) == (c["warm"]["next_p"], c["warm"]["cur_capacity"])
    for e, ref in enumerate(c["epochs"]):
        tag = f"{algo}/{c['variant']} epoch {e}"
        agent.epoch += 1
        agent.set_eval()
        traj, steps = agent.explore_env(env, acfg.horizon_len, sample=True)
        agent.memory.add_to_buffer(traj)
        check_traj(traj, ref["trajectory"], tag)
        assert (agent.memory.next_p, agent.memory.cur_capacity) == (ref["next_p"], ref["cur_capacity"])
        agent.set_train()
        res = agent.update_net(agent.memory)
        assert set(res) == set(ref["results"]), (sorted(res), sorted(ref["results"]))
        for k, v in ref["results"].items():
            scale_close(torch.stack(res[k]), v, 1e-5, f"{tag} {k}")
        for name in ("actor", "critic", "critic_target") + (() if acfg.no_tgt_actor else ("actor_target",)):
            sd = getattr(agent, name).state_dict()
            for k, v in ref[name].items():
                scale_close(sd[k], v, 1e-5, f"{tag} {name}.{k}")
        if algo == "sac":
            scale_close(agent.log_alpha.data, ref["log_alpha"], 1e-5, f"{tag} log_alpha")
        for k, v in ref["obs_rms"].items():
            scale_close(agent.obs_rms.state_dict()[k], v, 2e-6, f"{tag} obs_rms {k}")
    assert all(len(v) == 0 for v in rp.s.values()), {k: len(v) for k, v in rp.s.items()}   # every recorded draw was consumed
    if algo == "sac":   # checkpoint round trip in the reference's format (sac.py:414-453)
        import os
        f = os.path.join(agent.ckpt_dir, "t.pth")
        agent.save(f)
        ck = torch.load(f, map_location="cpu", weights_only=False)
        assert set(ck) == {"epoch", "mini_epoch", "agent_steps", "obs_rms", "encoder", "actor", "critic", "encoder_target", "actor_target",
                           "critic_target", "log_alpha"}
        before = {k: v.clone() for k, v in agent.actor.state_dict().items()}
        for p in agent.actor.parameters():
            p.data.zero_()
        agent.load(f)
        for k, v in before.items():
            eq(agent.actor.state_dict()[k], v.cpu(), f"reloaded actor.{k}")


@pytest.mark.parametrize("algo", ["sac", "ddpg"])
def test_train_loop_runs_like_the_reference(algo):
    """`train()` end to end (sac.py:190-249, ddpg.py:201-237): warm-up, explore / insert / update_net epochs until
    max_agent_steps, metrics rows with the reference's keys, the final checkpoint; SAC also `eval()`."""
    import os

    from mineral_b200.agents import DDPG, SAC
    from oracle.reference_shim import SynthEnv, to_attr

    c = load_golden("offpolicy_e2e")[0 if algo == "sac" else 2]
    N, O, A = c["shape"]
    cfg = to_attr(c["cfg"])
    cfg.rl_device = DEV
    acfg = cfg.agent[algo]
    epochs = 4
    acfg["max_agent_steps"] = N * (acfg.warm_up + epochs * acfg.horizon_len)
    cfg.agent["ckpt_every"] = 2
    cpu_env = SynthEnv(N, O, A, 4096, seed=3, done_p=0.1, timeout_p=0.05)

    class OnDevice:
        observation_space, action_space = cpu_env.observation_space, cpu_env.action_space

        def reset(self):
            return cu(cpu_env.reset())

        def step(self, actions):
            o, r, d, info = cpu_env.step(actions)
            return cu(o), cu(r), cu(d), cu(info)

    torch.manual_seed(0)
    agent = (SAC if algo == "sac" else DDPG)(cfg, env=OnDevice())
    agent.train()
    assert agent.epoch == epochs - 1 and agent.agent_steps == acfg.max_agent_steps and agent.mini_epoch == epochs * acfg.mini_epochs - 1
    rows = agent.writer.rows
    assert len(rows) == epochs
    last = rows[-1][1]
    need = {"train_stats/loss/critic", "train_stats/loss/actor", "train_stats/grad_norm/critic", "train_stats/grad_norm/actor", "train_stats/epoch",
            "train_scores/episode_rewards", "train_scores/episode_lengths", "train_scores/num_episodes"}
    if algo == "sac":
        need |= {"train_stats/alpha", "train_stats/entropy", "train_stats/loss/alpha", "train_stats/target_values/mean",
                 "train_timings/agent.update_net/avg", "train_timings/agent.explore_env/avg", "train_timings/totalrate"}
    assert need <= set(last), sorted(need - set(last))
    assert all(v == v for k, v in last.items() if isinstance(v, float)), "NaN in the metrics"
    assert last["train_scores/num_episodes"] > 0
    if algo == "sac":
        assert os.path.exists(os.path.join(agent.ckpt_dir, "final.pth"))
        ev = agent.eval()
        assert ev["eval_scores/num_episodes"] >= 2 * N
