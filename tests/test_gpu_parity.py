"""GPU parity: every kernel, called through the C ABI (ctypes), against (1) the committed golden
vectors produced by the unmodified reference and (2) the CPU oracle on the same seeded inputs.

Bar (BASELINE.json north_star): indexing, permutation, replay sampling and the scans are bit-exact;
fp32 statistics, losses and gradients agree within 1e-5 "relative" = max|a-b| <= 1e-5 * max|b|
(element-wise rtol is meaningless next to zero)."""

import numpy as np
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


# ------------------------------------------------------------------------------------------------
# K1a + K3a
# ------------------------------------------------------------------------------------------------
def _fill(buf, st):
    for k, v in st.items():
        if k == "returns":
            continue
        if isinstance(v, dict):
            for kk, vv in v.items():
                buf.storage_dict[k][kk].copy_(vv)
        else:
            buf.storage_dict[k].copy_(v)


def test_experience_buffer_matches_reference_golden():
    from mineral_b200.agents.ppo.experience import ExperienceBuffer

    for c in load_golden("experience"):
        T, N, O, A = c["shape"]
        for fused in (False, True):
            buf = ExperienceBuffer(N, T, N * T, c["minibatch_size"], {"obs": (O,)}, A, DEV)
            _fill(buf, c["inputs"])
            if fused:
                data = buf.compute_return_and_prepare(cu(c["last_values"]), c["gamma"], c["tau"])
            else:
                buf.compute_return(cu(c["last_values"]), c["gamma"], c["tau"])
                data = buf.prepare_training()
            eq(buf.storage_dict["returns"], c["returns_time_major"], "returns (time-major)")
            for k, v in c["data"].items():
                if k == "obses":
                    eq(data[k]["obs"], v["obs"], "obs")
                elif k == "advantages":
                    scale_close(data[k], v, 1e-5, "advantages")
                else:
                    eq(data[k], v, k)
            assert len(buf) == c["length"]
            mb = buf[1]
            names = ("values", "neglogp", "advantages", "mu", "sigma", "returns", "actions")
            for name, t in zip(names, mb[:7]):
                if name == "advantages":
                    scale_close(t, c["minibatch1"][name], 1e-5, name)
                else:
                    eq(t, c["minibatch1"][name], f"minibatch1 {name}")
            eq(mb[7]["obs"], c["minibatch1"]["obses"]["obs"], "minibatch1 obs")
            # update_mu_sigma writes back into the same rows (experience.py:83-87)
            new_mu = torch.randn_like(mb[3])
            new_sigma = torch.rand_like(mb[4])
            buf.update_mu_sigma(new_mu, new_sigma)
            s, e = buf.last_range
            eq(buf.data_dict["mu"][s:e], new_mu.cpu())
            eq(buf.data_dict["sigma"][s:e], new_sigma.cpu())
            eq(buf.data_dict["mu"][:s], c["data"]["mu"][:s])


@pytest.mark.parametrize("T,N,O,A,done_p", [(32, 64, 37, 8, 0.02), (32, 4096, 37, 8, 0.03), (16, 16384, 88, 16, 0.0),
                                            (8, 6, 5, 3, 1.0), (5, 10, 3, 1, 0.3), (32, 4096, 37, 8, 1.0)])
def test_gae_and_fused_pack_vs_oracle(T, N, O, A, done_p):
    from mineral_b200 import ops
    from mineral_b200.agents.ppo.experience import ExperienceBuffer
    from mineral_b200.common.normalizers import RunningMeanStd

    st, last = oc.synth_rollout(T, N, O, A, seed=1, done_p=done_p)
    ref_ret, ref_adv = oc.gae_scan(st["rewards"], st["values"], st["dones"], last, 0.99, 0.95)
    adv_dev = torch.empty(T, N, 1, device=DEV)
    ret_dev = ops.gae(cu(st["rewards"]), cu(st["values"]), cu(st["dones"]), cu(last), 0.99, 0.95, advantages=adv_dev)
    eq(ret_dev, ref_ret, "gae returns")
    eq(adv_dev, ref_adv, "gae advantages")

    mb = T * N // 4 if (T * N) % 4 == 0 else T * N
    buf = ExperienceBuffer(N, T, N * T, mb, {"obs": (O,)}, A, DEV)
    _fill(buf, st)
    vr = RunningMeanStd((1,), eps=1e-5, with_clamp=True, initial_count=1, dtype=torch.float64).to(DEV)
    data = buf.compute_return_and_prepare(cu(last), 0.99, 0.95, vr)

    st["returns"] = ref_ret
    ref = oc.pack_rollout(st)
    raw_adv = (ref["returns"] - ref["values"]).squeeze(1)
    ref_state = oc.rms_init((1,), 1e-5, 1, torch.float64)
    ref = oc.ppo_value_block(ref_state, ref)
    for k in ("rewards", "neglogp", "dones", "actions", "mu", "sigma"):
        eq(data[k], ref[k], k)
    eq(data["obses"]["obs"], ref["obses"]["obs"], "obs")
    scale_close(data["advantages"], ref["advantages"], 1e-5, "advantages")
    scale_close(data["values"], ref["values"], 1e-5, "normalised values")
    scale_close(data["returns"], ref["returns"], 1e-5, "normalised returns")
    scale_close(vr.running_mean, ref_state[0], 1e-6, "value_rms mean")
    scale_close(vr.running_var, ref_state[1], 1e-6, "value_rms var")
    scale_close(vr.running_count, ref_state[2], 0, "value_rms count")
    stats = buf.stats.cpu()
    scale_close(stats[0], raw_adv.double().mean(), 1e-5, "adv mean")
    scale_close(stats[1], raw_adv.double().std(), 1e-6, "adv std")


def test_pack_properties_at_full_size():
    """Size-independent properties at BASELINE cfg2 size: the pack is a pure permutation (multiset
    of every field preserved, row n*T+t == [t,n]) and normalised advantages have mean 0 / std 1."""
    from mineral_b200.agents.ppo.experience import ExperienceBuffer

    T, N, O, A = 32, 4096, 37, 8
    g = torch.Generator(device=DEV).manual_seed(0)
    buf = ExperienceBuffer(N, T, N * T, 32768, {"obs": (O,)}, A, DEV)
    sd = buf.storage_dict
    for k in ("rewards", "values", "neglogp", "actions", "mu", "sigma"):
        sd[k].copy_(torch.randn(sd[k].shape, device=DEV, generator=g))
    sd["obses"]["obs"].copy_(torch.randn(T, N, O, device=DEV, generator=g))
    sd["dones"].copy_((torch.rand(T, N, device=DEV, generator=g) < 0.02).to(torch.uint8))
    data = buf.compute_return_and_prepare(torch.randn(N, 1, device=DEV, generator=g), 0.99, 0.95)
    for k in ("actions", "mu", "sigma"):
        assert torch.equal(data[k].view(N, T, A), sd[k].transpose(0, 1))
    assert torch.equal(data["obses"]["obs"].view(N, T, O), sd["obses"]["obs"].transpose(0, 1))
    assert torch.equal(data["dones"].view(N, T), sd["dones"].transpose(0, 1))
    assert torch.equal(data["returns"].view(N, T), sd["returns"].squeeze(-1).transpose(0, 1))
    adv = data["advantages"].double()
    assert abs(adv.mean().item()) < 1e-5 and abs(adv.std().item() - 1.0) < 1e-5


# ------------------------------------------------------------------------------------------------
# K2
# ------------------------------------------------------------------------------------------------
def test_running_mean_std_matches_reference_golden():
    from mineral_b200.common.normalizers import RunningMeanStd

    for c in load_golden("rms"):
        kw = dict(c["kwargs"])
        if "dtype" in kw:
            kw["dtype"] = torch.float64 if kw["dtype"] == "torch.float64" else torch.float32
        rms = RunningMeanStd(c["shape"], **kw).to(DEV)
        assert list(rms.state_dict().keys()) == c["state_dict_keys"]
        for s in c["steps"]:
            x = cu(s["x"])
            rms.update(x)
            assert rms.running_mean.dtype == s["mean"].dtype and rms.running_mean.shape == s["mean"].shape
            scale_close(rms.running_mean, s["mean"], 2e-6, f"{c['name']} mean")
            scale_close(rms.running_var, s["var"], 2e-6, f"{c['name']} var")
            scale_close(rms.running_count, s["count"], 1e-7, f"{c['name']} count")
            # pin the state to the reference's so normalise / unnormalise are compared on equal stats
            rms.running_mean.copy_(s["mean"]), rms.running_var.copy_(s["var"]), rms.running_count.copy_(s["count"])
            # not bit-compared: torch's CPU vectorised sqrt is 1 ulp off IEEE in some lanes (the goldens
            # come from the CPU reference); CUDA sqrt.rn is exact.  1e-6 of scale instead.
            y = rms.normalize(x)
            assert y.shape == s["normalized"].shape and y.dtype == torch.float32
            scale_close(y, s["normalized"], 1e-6, f"{c['name']} normalize")
            scale_close(rms.unnormalize(cu(s["normalized"]) * 1.3), s["unnormalized"], 1e-6, f"{c['name']} unnormalize")


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_rms_merge_bit_exact_and_schedule(dtype):
    from mineral_b200 import ops

    g = torch.Generator().manual_seed(3)
    cols, rows, mb = 37, 4096, 512
    x = torch.randn(rows, cols, generator=g) * torch.linspace(0.1, 9, cols) + 1.5
    st = oc.rms_init((cols,), 1e-5, 1, dtype)
    dst = [t.clone().to(DEV) for t in st]
    # (a) one merge from torch's own batch moments: op-for-op identical => bit exact
    bm, bv = x[:mb].mean(0), x[:mb].var(0, correction=1)
    ops.rms_merge(dst[0], dst[1], dst[2], cu(bm), cu(bv), mb)
    ref = oc.rms_merge(st[0], st[1], st[2], bm, bv, mb)
    for a, b in zip(dst, ref):
        eq(a, b, "merge")
    # (b) batch moments per group vs torch
    gm, gv = ops.rms_batch_moments(cu(x), mb, 1)
    ref_gm = torch.stack([x[i * mb:(i + 1) * mb].mean(0) for i in range(rows // mb)])
    ref_gv = torch.stack([x[i * mb:(i + 1) * mb].var(0, correction=1) for i in range(rows // mb)])
    scale_close(gm, ref_gm, 1e-6, "group mean")
    scale_close(gv, ref_gv, 1e-6, "group var")
    # (c) the schedule == the sequence of merges (bit exact given the same group moments)
    n_steps = 3 * (rows // mb) + 2
    mt, dt = torch.empty(n_steps, cols, device=DEV), torch.empty(n_steps, cols, device=DEV)
    state = [t.clone().to(DEV) for t in oc.rms_init((cols,), 1e-5, 1, dtype)]
    ops.rms_schedule(state[0], state[1], state[2], gm, gv, mb, rows, n_steps, 1e-5, mt, dt)
    r = oc.rms_init((cols,), 1e-5, 1, dtype)
    gm_c, gv_c = gm.cpu(), gv.cpu()
    for s in range(n_steps):
        i = s % (rows // mb)
        r = list(oc.rms_merge(r[0], r[1], r[2], gm_c[i], gv_c[i], mb))
        eq(mt[s], r[0].float(), f"mean_tab[{s}]")
        eq(dt[s], torch.from_numpy(np.sqrt((r[1].float() + 1e-5).numpy())), f"denom_tab[{s}]")  # numpy sqrt is IEEE-exact
    for a, b in zip(state, r):
        eq(a, b, "final state")


def test_rms_ragged_groups_and_wide_columns():
    from mineral_b200 import ops

    g = torch.Generator().manual_seed(4)
    for rows, cols, grp in [(1000, 5, 300), (257, 600, 257), (33, 1, 33), (131072, 1, 131072), (64, 76, 64)]:
        x = torch.randn(rows, cols, generator=g) * 3 + 0.7
        gm, gv = ops.rms_batch_moments(cu(x), grp, 1)
        ng = (rows + grp - 1) // grp
        ref_m = torch.stack([x[i * grp:(i + 1) * grp].double().mean(0) for i in range(ng)])
        ref_v = torch.stack([x[i * grp:(i + 1) * grp].double().var(0, correction=1) for i in range(ng)])
        scale_close(gm, ref_m, 1e-6, f"mean {rows}x{cols}")
        scale_close(gv, ref_v, 1e-6, f"var {rows}x{cols}")


# ------------------------------------------------------------------------------------------------
# K1b
# ------------------------------------------------------------------------------------------------
def test_shac_targets_match_reference_golden():
    from mineral_b200.agents.diffrl.shac import SHAC

    for c in load_golden("tdlambda"):
        T, N = c["rew"].shape
        ag = SHAC.from_shapes(N, 4, 4, T, DEV, gamma=c["gamma"], lam=c["lam"], critic_method=c["method"], with_logprobs=True)
        ag.rew_buf.copy_(c["rew"]), ag.done_mask.copy_(c["done_mask"]), ag.next_values.copy_(c["next_values"])
        ag.distr_ent.copy_(c["distr_ent"]), ag.logprobs.copy_(c["logprobs"])
        ag.compute_target_values()
        eq(ag.target_values, c["target"], f"{c['method']} targets")
        for v in c["entropy_variants"]:
            ag.use_distr_ent, ag.offset_by_target_entropy, ag.scale_by_target_entropy = v["use_distr_ent"], v["offset"], v["scale"]
            ag.entropy_coef, ag.target_entropy = v["alpha"], c["target_entropy"]
            ag.compute_target_values_with_entropy()
            eq(ag.target_values, v["target"], f"{c['method']} entropy targets {v}")


def test_td_lambda_humanoid_shape_vs_oracle():
    from mineral_b200 import ops

    g = torch.Generator().manual_seed(9)
    for T, N in [(32, 64), (32, 4096), (7, 33)]:
        rew, nv = torch.randn(T, N, generator=g), 2 * torch.randn(T, N, generator=g)
        dm = (torch.rand(T, N, generator=g) < 0.05).float()
        dm[-1] = 1.0
        out = ops.critic_targets(cu(rew), cu(dm), cu(nv), 0.99, 0.95, "td-lambda")
        eq(out, oc.td_lambda_targets(rew, dm, nv, 0.99, 0.95), f"td-lambda {T}x{N}")


@pytest.mark.parametrize("cuda_graph", [False, True])
def test_shac_update_critic_matches_reference_golden(cuda_graph):
    """a16: SHAC.update_critic on our kernels (critic MLP fwd/bwd, MSE + gradient, nan_to_num + clip + Adam; optionally one
    CUDA graph) against the reference's own run on its Critic / EnsembleCritic, Humanoid shape (cfg3)."""
    from mineral_b200.agents.diffrl.shac import SHAC

    for c in load_golden("shac_critic"):
        T, B, O = c["shape"]
        obs, tv = oc.shac_critic_inputs(c["seed"], T, B, O, c["wild"])
        ag = SHAC.from_shapes(B, O, 21, T, DEV, critic_iterations=c["iterations"], num_critics=c["n_critics"], critic_lr=c["lr"],
                              critic_betas=c["betas"], max_grad_norm=c["max_grad_norm"], cuda_graph=cuda_graph)
        ag.critic_batch_size = c["batch_size"]
        ag.obs_buf["obs"].copy_(obs), ag.target_values.copy_(tv)
        snap = ag.critic_optim.state_dict()
        for rep in range(3 if cuda_graph else 1):  # eager warm-up, capture + replay, replay: each from the same state
            ag.critic.load_state_dict(c["init_state_dict"])
            ag.critic_optim.load_state_dict(snap)
            res = ag.update_critic()
            tag = f"critics={c['n_critics']} graph={cuda_graph} rep={rep}"
            for k, v in c["results"].items():
                scale_close(torch.stack(res[k]), v, 1e-5, f"{tag} {k}")
            for k, v in c["final_state_dict"].items():
                scale_close(ag.critic.state_dict()[k], v, 1e-5, f"{tag} param {k}")
        assert (ag._critic_engine.graph is not None) == cuda_graph


def test_critic_dataset_matches_reference_golden():
    from mineral_b200.agents.diffrl.utils import CriticDataset

    for c in load_golden("critic_dataset"):
        if c["shuffle"]:
            np.random.seed(c["np_seed"])
        ds = CriticDataset(c["batch_size"], cu(c["obs"]), cu(c["target_values"]), shuffle=c["shuffle"], drop_last=c["drop_last"])
        assert len(ds) == c["length"]
        for i, b in enumerate(c["batches"]):
            obs, tv = ds[i]
            eq(obs["obs"], b["obs"]["obs"], "critic batch obs")
            eq(tv, b["target"], "critic batch target")


# ------------------------------------------------------------------------------------------------
# K3c/d/e
# ------------------------------------------------------------------------------------------------
def test_replay_buffer_matches_reference_golden():
    from mineral_b200.buffers import ReplayBuffer

    g = load_golden("replay")
    rb = ReplayBuffer({"obs": (g["obs_dim"],)}, g["act_dim"], g["capacity"], device=DEV)
    rb.buf_obs["obs"].zero_(), rb.buf_next_obs["obs"].zero_(), rb.buf_action.zero_(), rb.buf_reward.zero_(), rb.buf_done.zero_()
    for ev in g["events"]:
        i = cu(ev["insert"])
        rb.add_to_buffer(({"obs": i["obs"]}, i["act"], i["rew"], {"obs": i["nobs"]}, i["done"]))
        assert (rb.next_p, rb.if_full, rb.cur_capacity) == (ev["next_p"], ev["if_full"], ev["cur_capacity"])
        eq(rb.buf_obs["obs"], ev["buf"]["obs"], "ring obs")
        eq(rb.buf_action, ev["buf"]["act"], "ring act")
        eq(rb.buf_reward, ev["buf"]["rew"], "ring rew")
        eq(rb.buf_next_obs["obs"], ev["buf"]["nobs"], "ring nobs")
        eq(rb.buf_done, ev["buf"]["done"], "ring done")
        obs, act, rew, nobs, done = rb.sample_batch(16, device=DEV, indices=cu(ev["idx"]))
        for name, t in (("obs", obs["obs"]), ("act", act), ("rew", rew), ("nobs", nobs["obs"]), ("done", done)):
            eq(t, ev["sample"][name], f"sample {name}")
    # float dones (as SAC passes them) and sampling through torch.randint on the device
    rb.add_to_buffer(({"obs": torch.ones(3, g["obs_dim"], device=DEV)}, torch.ones(3, g["act_dim"], device=DEV),
                      torch.ones(3, 1, device=DEV), {"obs": torch.ones(3, g["obs_dim"], device=DEV)},
                      torch.tensor([[0.0], [1.0], [0.5]], device=DEV)))
    torch.manual_seed(5)
    idx = torch.randint(rb.cur_capacity, size=(32,), device=DEV)
    torch.manual_seed(5)
    obs, act, rew, nobs, done = rb.sample_batch(32, device=DEV)
    eq(obs["obs"], rb.buf_obs["obs"][idx].cpu(), "randint sample")
    eq(done, rb.buf_done[idx].float().cpu(), "randint done")


def test_replay_sample_fused_normalize_cfg4_shape():
    from mineral_b200.buffers import ReplayBuffer
    from mineral_b200.common.normalizers import RunningMeanStd

    O, A, C = 37, 8, 200_000
    rb = ReplayBuffer({"obs": (O,)}, A, C, device=DEV)
    g = torch.Generator(device=DEV).manual_seed(0)
    rows = 64 * 32
    for _ in range(3):
        rb.next_p = C - rows // 2 if _ == 1 else rb.next_p  # force a wrap on the second insert
        tr = ({"obs": torch.randn(rows, O, device=DEV, generator=g) * 3 + 1}, torch.randn(rows, A, device=DEV, generator=g),
              torch.randn(rows, 1, device=DEV, generator=g), {"obs": torch.randn(rows, O, device=DEV, generator=g)},
              torch.rand(rows, 1, device=DEV, generator=g) < 0.1)
        before = rb.next_p
        rb.add_to_buffer(tr)
        slots = (before + torch.arange(rows, device=DEV)) % C
        assert torch.equal(rb.buf_obs["obs"][slots], tr[0]["obs"]) and torch.equal(rb.buf_done[slots], tr[4])
    rms = torch.nn.ModuleDict({"obs": RunningMeanStd((O,), eps=1e-4, with_clamp=True, initial_count="eps")}).to(DEV)
    rms["obs"].update(rb.buf_obs["obs"][:rows])
    idx = torch.randint(rb.cur_capacity, (8192,), device=DEV)
    obs, act, rew, nobs, done = rb.sample_batch(8192, device=DEV, indices=idx, obs_rms=rms)
    eq(obs["obs"], rms["obs"].normalize(rb.buf_obs["obs"][idx]).cpu(), "fused normalised obs")
    eq(nobs["obs"], rms["obs"].normalize(rb.buf_next_obs["obs"][idx]).cpu(), "fused normalised next_obs")
    eq(act, rb.buf_action[idx].cpu())


def test_nstep_matches_reference_golden():
    from mineral_b200.buffers import NStepReplay

    for c in load_golden("nstep"):
        ns = NStepReplay({"obs": (c["O"],)}, c["A"], c["N"], c["nstep"], device=DEV)
        for call in c["calls"]:
            i = cu(call["inp"])
            out = ns.add_to_buffer({"obs": i["obs"]}, i["act"], i["rew"], {"obs": i["nobs"]}, i["done"])
            if call["out"] is None:  # reference raised while warming up; we return empty tensors
                assert out[1].shape[0] == 0
                continue
            got = dict(obs=out[0]["obs"], act=out[1], rew=out[2], nobs=out[3]["obs"], done=out[4])
            for name, t in got.items():
                eq(t, call["out"][name], f"nstep{c['nstep']} {name}")


def test_nstep_float_dones_vs_oracle():
    from mineral_b200.buffers import NStepReplay

    N, O, A, n = 64, 37, 8, 3
    g = torch.Generator().manual_seed(11)
    ns, ref = NStepReplay({"obs": (O,)}, A, N, n, device=DEV), oc.NStepOracle(N, O, A, n, 0.99)
    for S in (32, 1, 5):
        obs, act = torch.randn(N, S, O, generator=g), torch.randn(N, S, A, generator=g)
        rew, nobs = torch.randn(N, S, 1, generator=g), torch.randn(N, S, O, generator=g)
        done = (torch.rand(N, S, 1, generator=g) < 0.1).float()
        out = ns.add_to_buffer({"obs": cu(obs)}, cu(act), cu(rew), {"obs": cu(nobs)}, cu(done))
        r = ref.push(obs, act, rew, nobs, done)
        for name, a, b in zip(("obs", "act", "rew", "nobs", "done"), (out[0]["obs"], out[1], out[2], out[3]["obs"], out[4]), r):
            eq(a, b, f"nstep S={S} {name}")


# ------------------------------------------------------------------------------------------------
# K4
# ------------------------------------------------------------------------------------------------
def test_module_level_losses_match_reference_golden():
    from mineral_b200.agents.ppo.ppo import actor_loss, bounds_loss, critic_loss, policy_kl

    for c in load_golden("losses"):
        d = cu({k: v for k, v in c.items() if torch.is_tensor(v)})
        nlp, v, mu = d["nlp"].requires_grad_(True), d["v"].requires_grad_(True), d["mu"].requires_grad_(True)
        a, frac = actor_loss(d["old_nlp"], nlp, d["adv"], c["e_clip"], c["smooth"])
        cl, ev = critic_loss(d["vold"], v, c["e_clip"], d["ret"], c["clip_value"])
        b = bounds_loss(mu, c["bounds_type"])
        kl = policy_kl(mu.detach(), d["sigma"], d["old_mu"], d["old_sigma"])
        scale_close(a, c["a_loss"], 1e-5, "a_loss"), scale_close(frac, c["clip_frac"], 1e-6, "clip_frac")
        scale_close(cl, c["c_loss"], 1e-5, "c_loss"), scale_close(ev, c["explained_var"], 1e-5, "explained_var")
        scale_close(b, c["b_loss"], 1e-5, "b_loss"), scale_close(kl, c["kl"], 1e-5, "kl")
        assert a.shape == c["a_loss"].shape and cl.shape == c["c_loss"].shape and b.shape == c["b_loss"].shape
        (a.mean() + 0.5 * cl.mean() * 4 + b.mean() * 1e-4).backward()
        scale_close(nlp.grad, c["g_nlp"], 1e-5, "g_nlp"), scale_close(v.grad, c["g_v"], 1e-5, "g_v")
        scale_close(mu.grad, c["g_mu"], 1e-5, "g_mu")


@pytest.mark.parametrize("dist_type,fixed_sigma,mb,A", [("normal", True, 512, 8), ("dreamerv3_normal", False, 2048, 8),
                                                        ("normal", False, 333, 3), ("dreamerv3_normal", False, 32768, 16),
                                                        ("normal", True, 100, 21)])
@pytest.mark.parametrize("clip_value,smooth,btype", [(False, False, "bound"), (True, True, "reg")])
def test_fused_loss_vs_oracle_autograd(dist_type, fixed_sigma, mb, A, clip_value, smooth, btype):
    from mineral_b200 import ops

    g = torch.Generator().manual_seed(mb + A)
    mu_raw = (0.8 * torch.randn(mb, A, generator=g)).requires_grad_(True)
    ls = ((0.3 * torch.randn(A, generator=g)) if fixed_sigma else 0.3 * torch.randn(mb, A, generator=g)).requires_grad_(True)
    value = torch.randn(mb, 1, generator=g).requires_grad_(True)
    old_mu, old_sigma = torch.randn(mb, A, generator=g) * 0.7, torch.rand(mb, A, generator=g) * 0.8 + 0.2
    actions = old_mu + old_sigma * torch.randn(mb, A, generator=g)
    old_nlp = oc.normal_neglogp(old_mu, old_sigma, actions) + 0.2 * torch.randn(mb, generator=g)
    adv, vold, ret = torch.randn(mb, generator=g), torch.randn(mb, 1, generator=g), torch.randn(mb, 1, generator=g)
    hp = dict(e_clip=0.2, critic_coef=4.0, entropy_coef=0.003, bounds_loss_coef=1e-3)

    logstd = mu_raw * 0 + ls
    mu, sigma = oc.dist_params(mu_raw, logstd, dist_type, 0.1, 1.0)
    nlp, ent = oc.normal_neglogp(mu, sigma, actions), oc.normal_entropy(sigma)
    a, frac = oc.actor_loss(old_nlp, nlp, adv, hp["e_clip"], smooth)
    c, ev = oc.critic_loss(vold, value, hp["e_clip"], ret, clip_value)
    b = oc.bounds_loss(mu, btype)
    total = a.mean() + 0.5 * c.mean() * hp["critic_coef"] - ent.mean() * hp["entropy_coef"] + b.mean() * hp["bounds_loss_coef"]
    total.backward()
    kl = oc.policy_kl(mu.detach(), sigma.detach(), old_mu, old_sigma)

    cfg = ops.make_loss_cfg(dist_type, 0.1, 1.0, hp["e_clip"], smooth, clip_value, hp["critic_coef"], hp["entropy_coef"],
                            hp["bounds_loss_coef"], btype)
    d_mu, d_sig = cu(old_mu), cu(old_sigma)
    g_mu, g_ls, g_v = torch.empty(mb, A, device=DEV), torch.empty(mb, A, device=DEV), torch.empty(mb, 1, device=DEV)
    metrics, mm, sm = torch.zeros(10, device=DEV), torch.zeros(A, device=DEV), torch.zeros(A, device=DEV)
    ws = ops.ppo_loss_workspace(mb, A, DEV)
    for _ in range(2):  # twice: the ticket counter must come back to zero
        d_mu.copy_(old_mu), d_sig.copy_(old_sigma)
        ops.ppo_loss_fwd_bwd(cu(mu_raw.detach()), cu(ls.detach()), cu(value.detach()), cu(actions), cu(old_nlp), cu(adv), cu(vold),
                             cu(ret), d_mu, d_sig, g_mu, g_ls, g_v, metrics, cfg, ws, True, mm, sm)
    ref_metrics = torch.stack([total.detach(), a.mean().detach(), c.mean().detach(), b.mean().detach(), ent.mean().detach(), frac, ev, kl])
    for i, name in enumerate(("total", "actor", "critic", "bounds", "entropy", "clip_frac", "explained_var", "kl")):
        scale_close(metrics[i], ref_metrics[i], 1e-5, name)
    scale_close(g_mu, mu_raw.grad, 1e-5, "d/d mu_raw")
    scale_close(g_v, value.grad, 1e-5, "d/d value")
    scale_close(g_ls.sum(0) if fixed_sigma else g_ls, ls.grad, 1e-5, "d/d logstd")
    scale_close(d_mu, mu.detach(), 1e-6, "mu write-back"), scale_close(d_sig, sigma.detach(), 1e-6, "sigma write-back")
    scale_close(mm, mu.detach().mean(0), 1e-5, "mu column mean"), scale_close(sm, sigma.detach().mean(0), 1e-5, "sigma mean")


# ------------------------------------------------------------------------------------------------
# a12
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("decoupled", [False, True])
@pytest.mark.parametrize("one_launch,n", [(False, 10_007), (True, 10_008), (True, 427_432)])
def test_fused_adam_vs_torch(decoupled, one_launch, n):
    from mineral_b200 import ops

    g = torch.Generator().manual_seed(2)
    p0 = torch.randn(n, generator=g)
    ref_p = p0.clone().requires_grad_(True)
    Opt = torch.optim.AdamW if decoupled else torch.optim.Adam
    ref_opt = Opt([ref_p], lr=5e-4, eps=1e-5)
    p, m, v = cu(p0.clone()), torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    st = ops.optim_state_tensor(5e-4, DEV)
    ws = torch.zeros(ops.lib.mnr_grad_norm_workspace_bytes(n), dtype=torch.uint8, device=DEV)
    fws = torch.zeros(ops.lib.mnr_fused_adam_workspace_bytes(), dtype=torch.uint8, device=DEV)
    kl = torch.full((1,), 0.25, device=DEV)
    norm_out = torch.zeros(1, device=DEV)
    for step in range(6):
        grad = torch.randn(n, generator=g) * (0.01 if step % 2 else 3.0)
        ref_p.grad = grad.clone()
        ref_norm = torch.nn.utils.clip_grad_norm_([ref_p], 1.0)
        ref_opt.step()
        gd = cu(grad)
        if one_launch:  # csrc/peer.cu, world == 1 instance: norm + clip + Adam + zero_grad behind one grid barrier
            ops.fused_clip_adam(p, gd, m, v, 1.0, 0.9, 0.999, 1e-5, 0.01 if decoupled else 0.0, decoupled, True, st, kl, norm_out, fws)
        else:
            ops.grad_norm_prepare(gd, 1.0, 0.9, 0.999, st, kl, norm_out, ws)
            ops.adam_step(p, gd, m, v, 0.9, 0.999, 1e-5, 0.01 if decoupled else 0.0, decoupled, True, st)
        scale_close(norm_out[0], ref_norm, 5e-6, "grad norm")  # torch accumulates its norm in f32
        assert gd.abs().max().item() == 0.0  # zero_grad fused
        scale_close(p, ref_p.detach(), 1e-6, f"params after step {step}")
    s = ops.read_optim_state(st)
    assert s.step == 6 and abs(s.lr - 5e-4) < 1e-18
    assert s.kl_count == 6 and abs(s.kl_sum - 1.5) < 1e-12


def test_kl_lr_rule_on_device():
    from mineral_b200 import ops

    st = ops.optim_state_tensor(5e-4, DEV)
    ws = torch.zeros(ops.lib.mnr_grad_norm_workspace_bytes(8), dtype=torch.uint8, device=DEV)
    lr = 5e-4
    avg, lr_out = torch.zeros(1, device=DEV), torch.zeros(1, dtype=torch.float64, device=DEV)
    for kls in ([0.001, 0.002], [0.05, 0.03, 0.04], [0.008, 0.009], [1e-5] * 12, [0.5] * 30):
        for k in kls:
            ops.grad_norm_prepare(torch.ones(8, device=DEV), 0.0, 0.9, 0.999, st, torch.tensor([k], device=DEV), None, ws)
        ops.kl_lr_update(st, 0.008, 1e-6, 1e-2, avg, lr_out)
        mean_kl = torch.tensor(kls, dtype=torch.float32).mean().item()
        lr = oc.kl_lr_update(lr, mean_kl, 0.008, 1e-6, 1e-2)
        assert lr_out.item() == lr, (kls, lr_out.item(), lr)
        scale_close(avg[0], torch.tensor(mean_kl), 1e-6)
    assert ops.read_optim_state(st).kl_count == 0


# ------------------------------------------------------------------------------------------------
# a8 + a9: the whole agent against the reference agent's recorded run
# ------------------------------------------------------------------------------------------------
def _agent_from_golden(c, cuda_graph=False):
    from mineral_b200.agents import PPO
    from oracle.reference_shim import SynthEnv, to_attr

    T, N, O, A = c["shape"]
    cfg = to_attr(c["cfg"])
    cfg.rl_device = DEV
    cfg.agent.ppo["cuda_graph"] = cuda_graph
    agent = PPO(cfg, env=SynthEnv(N, O, A, T))
    agent.model.load_state_dict(c["init_state_dict"])
    agent.value_rms.load_state_dict(c["value_rms_before"])
    agent.obs_rms.load_state_dict(c["obs_rms_before"])
    for k, v in c["storage"].items():
        if isinstance(v, dict):
            for kk, vv in v.items():
                agent.storage.storage_dict[k][kk].copy_(vv)
        else:
            agent.storage.storage_dict[k].copy_(v)
    return agent


@pytest.mark.parametrize("cuda_graph", [False, True])
def test_ppo_update_matches_reference_agent(cuda_graph):
    for c in load_golden("ppo_e2e"):
        agent = _agent_from_golden(c, cuda_graph)
        reps = 3 if cuda_graph else 1  # eager warm-up, capture+replay, replay: all from identical state
        snap = None
        for rep in range(reps):
            if rep > 0:
                agent.model.load_state_dict(c["init_state_dict"])
                agent.value_rms.load_state_dict(c["value_rms_before"])
                agent.obs_rms.load_state_dict(c["obs_rms_before"])
                agent.optim.load_state_dict(snap)
            else:
                snap = agent.optim.state_dict()
            agent.finish_rollout(cu(c["last_values"]))
            data = agent.storage.data_dict
            ref_play = c["data_after_play"]
            scale_close(data["advantages"], ref_play["advantages"], 1e-5, "advantages")
            scale_close(data["values"], ref_play["values"], 1e-5, "values")
            scale_close(data["returns"], ref_play["returns"], 1e-5, "returns")
            eq(data["obses"]["obs"], ref_play["obses"]["obs"], "obs")
            agent.set_train()
            res = agent.train_epoch()
            ref = c["results"]
            tag = f"{c['variant']} graph={cuda_graph} rep={rep}"
            for k in ("loss/total", "loss/actor", "loss/critic", "loss/bounds", "loss/entropy", "explained_var", "grad_norm/all"):
                scale_close(torch.stack(res[k]), ref[k], 1e-5, f"{tag} {k}")
            # KL is a sum of O(1) terms that cancel to ~1e-3 (log ratio + quadratic - 1/2 per action dim): its
            # rounding floor is ~1e-7 * A absolute in ANY f32 evaluation order, so it gets an absolute tolerance
            kl_err = (torch.stack(res["avg_kl"]).cpu() - ref["avg_kl"]).abs().max().item()
            assert kl_err <= 1e-5 * ref["avg_kl"].abs().max().item() + 2e-7, f"{tag} avg_kl {kl_err:.3e}"
            scale_close(torch.stack(res["clip_frac"]), ref["clip_frac"], 1.01 / c["minibatch_size"], f"{tag} clip_frac")
            assert agent.last_lr == c["last_lr"], (agent.last_lr, c["last_lr"])
            scale_close(torch.cat(res["mu"]).mean(0), c["results_mu_mean"], 1e-5, "mu mean")
            for k, v in c["final_state_dict"].items():
                scale_close(agent.model.state_dict()[k], v, 1e-5, f"{tag} param {k}")
            for k, v in c["obs_rms_after"].items():
                scale_close(agent.obs_rms.state_dict()[k], v, 2e-6, f"{tag} obs_rms {k}")
            for k, v in c["value_rms_after"].items():
                scale_close(agent.value_rms.state_dict()[k], v, 2e-6, f"{tag} value_rms {k}")
            scale_close(data["mu"], c["data_after_train"]["mu"], 1e-5, "mu write-back")
            scale_close(data["sigma"], c["data_after_train"]["sigma"], 1e-5, "sigma write-back")


def test_play_steps_matches_reference_rollout():
    """PPO.play_steps on our kernels (obs normalise -> tower GEMMs -> mnr_policy_step -> env.step ->
    mnr_post_step, then the fused post-rollout pass) against the REAL reference agent's rollout recorded in the ppo_e2e
    golden: same synthetic env streams, the reference's own sampling noise (recovered from its stored actions / mu /
    sigma) injected step by step."""
    from mineral_b200.agents import PPO
    from mineral_b200.agents.ppo.ppo import _RolloutEngine
    from oracle.reference_shim import SynthEnv, to_attr

    fused = True

    for c in load_golden("ppo_e2e"):
        T, N, O, A = c["shape"]
        cfg = to_attr(c["cfg"])
        cfg.rl_device = DEV
        cpu_env = SynthEnv(N, O, A, T, seed=41, done_p=0.08, timeout_p=0.05)  # oracle/make_golden.py:gen_ppo_e2e

        class OnDevice:  # the agents expect the simulator to hand out device tensors (mineral/agents/agent.py:144-156)
            observation_space, action_space = cpu_env.observation_space, cpu_env.action_space

            def reset(self):
                return cu(cpu_env.reset())

            def step(self, actions):
                o, r, d, info = cpu_env.step(actions)
                return cu(o), cu(r), cu(d), cu(info)

        env = OnDevice()
        agent = PPO(cfg, env=env)
        agent.model.load_state_dict(c["init_state_dict"])
        agent.value_rms.load_state_dict(c["value_rms_before"])
        agent.obs_rms.load_state_dict(c["obs_rms_before"])
        ref = c["storage"]
        eps = (ref["actions"] - ref["mu"]) / ref["sigma"]
        assert _RolloutEngine.supported(agent)
        agent._rollout = _RolloutEngine(agent)
        agent._rollout.eps_source = lambda t: cu(eps[t])
        seen = {}
        orig_finish = agent.finish_rollout

        def finish(last_values):
            seen["last_values"] = last_values.clone()
            return orig_finish(last_values)

        agent.finish_rollout = finish
        agent.obs = agent._convert_obs(env.reset())
        agent.dones = torch.zeros(N, dtype=torch.bool, device=DEV)
        agent.epoch = 0
        agent.set_eval()
        agent.play_steps()
        sd = agent.storage.storage_dict
        tag = f"{c['variant']} fused={fused}"
        eq(sd["obses"]["obs"], ref["obses"]["obs"], f"{tag} obs")
        eq(sd["dones"], ref["dones"], f"{tag} dones")
        for k in ("actions", "mu", "sigma", "values", "rewards"):
            scale_close(sd[k], ref[k], 1e-5, f"{tag} storage {k}")
        scale_close(sd["neglogp"], ref["neglogp"], 1e-5, f"{tag} storage neglogp")
        scale_close(seen["last_values"], c["last_values"], 1e-5, f"{tag} last_values")
        data = agent.storage.data_dict
        for k in ("advantages", "values", "returns"):
            scale_close(data[k], c["data_after_play"][k], 1e-5, f"{tag} data {k}")
        assert (agent._rollout is not None) == fused


@pytest.mark.parametrize("N,L,done_p", [(300, 20, 0.03), (4096, 100, 0.02), (64, 8, 0.6), (2500, 100, 0.0)])
def test_episode_accounting_on_device(N, L, done_p):
    """8(f)-4: mnr_episode_scores == Metrics._update_scores + Tracker (mineral/common/metrics.py:59-67, tracker.py:9-22),
    including steps in which more episodes finish than the window holds."""
    import collections

    from mineral_b200 import ops

    g = torch.Generator().manual_seed(N + L)
    cur_r, cur_l = torch.zeros(N), torch.zeros(N, dtype=torch.int64)
    win_r, win_l = collections.deque([0] * L, maxlen=L), collections.deque([0] * L, maxlen=L)
    d_cur_r, d_cur_l = cu(cur_r.clone()), cu(cur_l.clone())
    d_win_r, d_win_l, cnt = torch.zeros(L, device=DEV), torch.zeros(L, device=DEV), torch.zeros(1, dtype=torch.int64, device=DEV)
    total = 0
    for step in range(40):
        r = torch.randn(N, generator=g)
        dones = torch.rand(N, generator=g) < done_p
        # reference semantics on the CPU
        cur_r += r
        cur_l += 1
        idx = torch.where(dones)[0].tolist()
        if idx:
            total += len(idx)
            win_r.extend(cur_r[idx].tolist()), win_l.extend(cur_l[idx].tolist())
            cur_r[idx] = 0
            cur_l[idx] = 0
        ops.episode_scores(cu(r), cu(dones), d_cur_r, d_cur_l, d_win_r, d_win_l, cnt)
    eq(d_cur_r, cur_r, "current rewards"), eq(d_cur_l, cur_l, "current lengths")
    assert cnt.item() == total
    ring_r, ring_l = d_win_r.cpu().tolist(), d_win_l.cpu().tolist()
    order = [(total + i) % L for i in range(L)]  # deque order = ring rotated by the running episode count
    assert [ring_r[i] for i in order] == [float(torch.tensor(v, dtype=torch.float32)) for v in win_r]
    assert [ring_l[i] for i in order] == [float(v) for v in win_l]


def test_offpolicy_math_matches_reference_golden():
    """8(f)-2: Polyak update, SAC / DDPG TD target and the C51 projection on the B200 kernels, BIT-exact against the
    reference's own functions (goldens), plus larger seeded cases against the oracle."""
    import torch.nn as nn

    from mineral_b200.agents.ddpg import utils as du

    c = load_golden("offpolicy")
    su = c["soft_update"]
    tar, cur = nn.Linear(7, 5).to(DEV), nn.Linear(7, 5).to(DEV)
    tar.load_state_dict(su["target_before"]), cur.load_state_dict(su["current"])
    du.soft_update(tar, cur, su["tau"])
    for k, v in su["target_after"].items():
        eq(tar.state_dict()[k], v, f"soft_update {k}")
    d = c["distl"]
    got = du.distl_projection(cu(d["next_dist"]), cu(d["reward"]), cu(d["done"]), d["gamma"], d["v_min"], d["v_max"], d["num_atoms"], d["support"])
    eq(got, d["proj"], "distl_projection")
    s = c["sac_target"]
    for case in s["cases"]:
        got = du.td_target([cu(q) for q in s["q"]], cu(s["reward"]), cu(s["done"]), case["gamma"], case["nstep"],
                           case["alpha"] if case["backup_entropy"] else None, cu(s["log_prob"]))
        eq(got, case["target"], f"sac target {case['backup_entropy']} {case['alpha']}")
    # larger seeded cases vs the oracle (SAC batch 8192 of cfg4; flat Polyak over a 400k-parameter buffer)
    g = torch.Generator().manual_seed(9)
    B = 8192
    qs = [5 * torch.randn(B, 1, generator=g) for _ in range(3)]
    lp, rew, dn = torch.randn(B, 1, generator=g) - 3, torch.randn(B, 1, generator=g), (torch.rand(B, 1, generator=g) < 0.1).float()
    eq(du.td_target([cu(q) for q in qs], cu(rew), cu(dn), 0.99, 3, 0.07, cu(lp)), oc.td_target(qs, rew, dn, 0.99, 3, 0.07, lp), "td_target 8192")
    eq(du.td_target([cu(q) for q in qs[:1]], cu(rew), cu(dn), 0.99, 1), oc.td_target(qs[:1], rew, dn, 0.99, 1), "ddpg target")
    t, cflat = torch.randn(400_003, generator=g), torch.randn(400_003, generator=g)
    eq(du.soft_update_flat(cu(t.clone()), cu(cflat), 0.005), oc.soft_update(t, cflat, 0.005), "flat polyak")
    nd = torch.softmax(torch.randn(300, 51, generator=g), dim=1)
    rew2, dn2 = 4 * torch.randn(300, 1, generator=g), (torch.rand(300, 1, generator=g) < 0.2).float()
    sup = torch.linspace(-10, 10, 51)
    eq(du.distl_projection(cu(nd), cu(rew2), cu(dn2), 0.99, -10, 10, 51, sup), oc.distl_projection(nd, rew2, dn2, 0.99, -10, 10, 51, sup), "distl 300")


def test_ppo_first_step_gradients_vs_oracle():
    """Gradients of the first optimizer step (before any parameter moved) within 1e-5 of scale."""
    for c in load_golden("ppo_e2e"):
        from test_oracle_golden import _hp

        hp, net, p = _hp(c)
        hp = dict(hp, mini_epochs=1)
        params = {k: v.clone().requires_grad_(True) for k, v in c["init_state_dict"].items()}
        optim = getattr(torch.optim, p["optim_type"])(list(params.values()), **p["optim_kwargs"])
        st = {k: (dict(v) if isinstance(v, dict) else v.clone()) for k, v in c["storage"].items()}
        vr = [c["value_rms_before"][k].clone() for k in ("running_mean", "running_var", "running_count")]
        orms = [c["obs_rms_before"][f"obs.{k}"].clone() for k in ("running_mean", "running_var", "running_count")]
        rec = {}
        oc.ppo_update(st, c["last_values"], params, optim, orms, vr, hp, net, p["optim_kwargs"]["lr"], record=rec)

        agent = _agent_from_golden(c)
        agent.finish_rollout(cu(c["last_values"]))
        # one step by hand through the same engine pieces, without the optimizer
        from mineral_b200.agents.ppo.ppo import _UpdateEngine
        eng = _UpdateEngine(agent)
        a, d, mb = agent, agent.storage.data_dict, agent.minibatch_size
        n = eng._n_obs
        from mineral_b200 import ops
        ops.rms_batch_moments(d["obses"]["obs"], mb, 1, n["ws"], n["gm"], n["gv"])
        rms = a.obs_rms["obs"]
        ops.rms_schedule(rms.running_mean, rms.running_var, rms.running_count, n["gm"], n["gv"], mb, eng.B, eng.steps, rms.eps,
                         n["mean_tab"], n["denom_tab"])
        x = ops.rms_apply(d["obses"]["obs"][:mb], n["mean_tab"][0], n["denom_tab"][0], (-5.0, 5.0))
        scale_close(x, rec["first_x"], 1e-5, "normalised obs of step 0")
        mu_raw, logstd, value = a.model._actor_critic({"obs": x})
        ops.ppo_loss_fwd_bwd(mu_raw.detach(), logstd.detach().contiguous(), value.detach(), d["actions"][:mb], d["neglogp"][:mb],
                             d["advantages"][:mb], d["values"][:mb], d["returns"][:mb], d["mu"][:mb], d["sigma"][:mb], eng.g_mu,
                             eng.g_ls, eng.g_v, eng.metrics[0], eng.loss_cfg, eng.loss_ws, False)
        torch.autograd.backward([mu_raw, logstd, value], [eng.g_mu, eng.g_ls, eng.g_v])
        for k, prm in a.model.named_parameters():
            scale_close(prm.grad, rec["grads"][k], 1e-5, f"{c['variant']} grad {k}")


# ------------------------------------------------------------------------------------------------
# G: fused MLP towers vs torch autograd on the same module
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("rows,O,A,units,act,fixed_sigma,norm", [
    (512, 37, 8, [256, 128, 64], "ELU", True, None), (333, 11, 3, [32, 16], "ELU", False, None),
    (2048, 88, 16, [64, 64, 32, 16], "ReLU", True, None), (130, 5, 2, [8], "ELU", True, None),
    # the shipped PPO recipes: Linear -> LayerNorm -> SiLU (DFlexAntPPO.yaml / AllegroHandPPO.yaml), plus odd widths
    (1024, 37, 8, [512, 256, 128], "SiLU", False, "LayerNorm"), (700, 88, 16, [512, 256, 128], "SiLU", False, "LayerNorm"),
    (333, 11, 3, [48, 16], "ELU", True, "LayerNorm"), (257, 9, 4, [160, 32], None, True, "LayerNorm")])
@pytest.mark.parametrize("use_tc", [False, True])
def test_fused_towers_vs_autograd(rows, O, A, units, act, fixed_sigma, norm, use_tc):
    from mineral_b200.agents.ppo.models import ActorCritic
    from mineral_b200.nets.fused_towers import FusedTowers

    torch.manual_seed(rows)
    model = ActorCritic({"obs": (O,)}, A, mlp_kwargs=dict(units=units, act_type=act, norm_type=norm), fixed_sigma=fixed_sigma).to(DEV)
    with torch.no_grad():
        for p in model.parameters():
            p.add_(0.05 * torch.randn_like(p))  # biases / sigma away from zero
    assert FusedTowers.supported(model)
    x = torch.randn(rows, O, device=DEV)
    g_mu, g_ls, g_v = torch.randn(rows, A, device=DEV), torch.randn(rows, A, device=DEV), torch.randn(rows, 1, device=DEV)
    mu, ls, v = model._actor_critic({"obs": x})
    torch.autograd.backward([mu, ls, v], [g_mu, g_ls, g_v])
    ref = {k: p.grad.clone() for k, p in model.named_parameters()}
    for p in model.parameters():
        p.grad = torch.full_like(p, float("nan"))  # every gradient must be (over)written
    ft = FusedTowers(model, rows, DEV, use_tc=use_tc)
    if use_tc and ft.use_tc:  # first-layer input through a padded row stride, as the PPO engine provides it
        x_ld = (O + 3) // 4 * 4
        xp = torch.zeros(rows, x_ld, device=DEV)
        xp[:, :O] = x
        x_in = xp
    else:
        x_in, x_ld = x, None
    mu2, ls2, v2 = ft.forward(x_in, x_ld)
    scale_close(mu2, mu.detach(), 1e-5, "mu_raw"), scale_close(v2, v.detach(), 1e-5, "value")
    scale_close(ls2.expand(rows, A) if fixed_sigma else ls2, ls.detach(), 1e-5, "logstd")
    ft.backward(x_in, g_mu, g_ls, g_v)
    for k, p in model.named_parameters():
        # bias gradients here are sums of `rows` N(0,1) seeds that cancel to O(sqrt(rows)): both f32 summation orders
        # carry ~rows * 2^-24 absolute error, so they get a looser bound than the weights
        # A bias gradient is the plain sum of 32768 dZ values: torch's f32 autograd and the kernels (per-thread partial sums, fixed
        # tree) add them in different orders, and the sum cancels to ~1e-2 of its largest terms.  Measured 1.1e-5 at worst
        # (profiles/r02_tolerances.tsv); everything else in this test sits below 1e-5.
        scale_close(p.grad, ref[k], 2e-5 if k.endswith("bias") or k == "sigma" else 1e-5, f"grad {k}")


@pytest.mark.parametrize("rows,A,H,fixed_sigma,dist_type", [(512, 8, 64, True, "normal"), (2048, 8, 64, False, "dreamerv3_normal"),
                                                             (333, 3, 16, True, "normal"), (1000, 16, 128, False, "dreamerv3_normal"),
                                                             (32768, 8, 64, True, "normal"), (300, 21, 64, True, "normal"),
                                                             (500, 6, 64, False, "dreamerv3_normal"), (260, 30, 128, False, "normal"),
                                                             (129, 1, 32, True, "normal"),
                                                             # above 9472 rows the kernel runs 128-row tiles (2 threads per row), below 32-row tiles (8)
                                                             (12001, 16, 128, False, "dreamerv3_normal"), (9601, 8, 32, True, "normal"),
                                                             (9472, 8, 128, False, "dreamerv3_normal")])
def test_fused_heads_loss_vs_unfused(rows, A, H, fixed_sigma, dist_type):
    """ppo_heads_loss_kernel + head_grads_kernel == heads forward (SIMT) + ppo_loss_kernel + heads backward."""
    from mineral_b200 import ops
    from mineral_b200.agents.ppo.models import ActorCritic
    from mineral_b200.nets.fused_towers import FusedTowers

    torch.manual_seed(rows + A)
    O = 12
    dist_kwargs = dict(dist_type=dist_type, minstd=0.1, maxstd=1.0) if dist_type != "normal" else dict(dist_type="normal")
    model = ActorCritic({"obs": (O,)}, A, mlp_kwargs=dict(units=[32, H], act_type="ELU"), fixed_sigma=fixed_sigma,
                        actor_dist_kwargs=dist_kwargs).to(DEV)
    with torch.no_grad():
        for p in model.parameters():
            p.add_(0.1 * torch.randn_like(p))
    for p in model.parameters():
        p.grad = torch.zeros_like(p)
    ft = FusedTowers(model, rows, DEV, use_tc=False)
    assert ft.fused_heads_ok
    g = torch.Generator(device=DEV).manual_seed(1)
    x = torch.randn(rows, O, device=DEV, generator=g)
    d = dict(actions=torch.randn(rows, A, device=DEV, generator=g), neglogp=torch.randn(rows, device=DEV, generator=g) + 8,
             advantages=torch.randn(rows, device=DEV, generator=g), values=torch.randn(rows, 1, device=DEV, generator=g),
             returns=torch.randn(rows, 1, device=DEV, generator=g), mu=torch.randn(rows, A, device=DEV, generator=g),
             sigma=torch.rand(rows, A, device=DEV, generator=g) + 0.2)
    cfg = ops.make_loss_cfg(dist_type, 0.1, 1.0, 0.2, False, True, 4.0, 0.002, 1e-3, "bound")
    g_mu, g_ls, g_v = (torch.zeros(rows, A, device=DEV), torch.zeros(rows, A, device=DEV), torch.zeros(rows, 1, device=DEV))
    ft.forward_towers(x)

    # unfused reference path
    d1 = {k: v.clone() for k, v in d.items()}
    mu_raw, ls, value = ft.heads_forward()
    m1, mm1, sm1 = torch.zeros(10, device=DEV), torch.zeros(A, device=DEV), torch.zeros(A, device=DEV)
    ops.ppo_loss_fwd_bwd(mu_raw, ls.contiguous(), value, d1["actions"], d1["neglogp"], d1["advantages"], d1["values"], d1["returns"],
                         d1["mu"], d1["sigma"], g_mu, g_ls, g_v, m1, cfg, ops.ppo_loss_workspace(rows, A, DEV), True, mm1, sm1)
    ft.heads_backward(g_mu, g_ls, g_v)
    head_names = [n for n, _ in model.named_parameters() if n.split(".")[0] in ("mu", "sigma", "value")]
    ref_grads = {n: dict(model.named_parameters())[n].grad.clone() for n in head_names}
    ref_dy = [t.clone() for t in ft.dY[-1]]

    # fused path
    for n in head_names:
        dict(model.named_parameters())[n].grad.fill_(float("nan"))
    for t in ft.dY[-1]:
        t.fill_(float("nan"))
    d2 = {k: v.clone() for k, v in d.items()}
    m2, mm2, sm2 = torch.zeros(10, device=DEV), torch.zeros(A, device=DEV), torch.zeros(A, device=DEV)
    ws = ops.ppo_heads_workspace(rows, A, H, DEV)
    for _ in range(2):  # twice: tickets must self-reset
        d2 = {k: v.clone() for k, v in d.items()}
        ft.heads_loss_fwd_bwd(d2, 0, rows, g_mu, g_ls, g_v.view(-1), m2, cfg, ws, mm2, sm2)
    scale_close(m2[:8], m1[:8], 1e-5, "metrics")
    scale_close(mm2, mm1, 1e-5, "mu mean"), scale_close(sm2, sm1, 1e-5, "sigma mean")
    scale_close(d2["mu"], d1["mu"], 1e-5, "mu write-back"), scale_close(d2["sigma"], d1["sigma"], 1e-5, "sigma write-back")
    for t, r in zip(ft.dY[-1], ref_dy):
        scale_close(t, r, 1e-5, "dY3")
    for n in head_names:
        scale_close(dict(model.named_parameters())[n].grad, ref_grads[n], 1e-5, f"head grad {n}")
    # cfg.dz_act (single-kernel form): the emitted gradient is dY3 * ELU'(Y3), taken from the staged Y3 tiles - bit for bit the
    # product of the dY3 above with act'(Y3); everything else is unchanged
    if ops.lib.mnr_ppo_heads_is_fused(A, H):
        dy_plain = [t.clone() for t in ft.dY[-1]]
        grads_plain = {n: dict(model.named_parameters())[n].grad.clone() for n in head_names}
        cfg.dz_act = 1
        d3 = {k: v.clone() for k, v in d.items()}
        m3 = torch.zeros(10, device=DEV)
        ft.heads_loss_fwd_bwd(d3, 0, rows, g_mu, g_ls, g_v.view(-1), m3, cfg, ws, mm2, sm2)
        for t, r, y3 in zip(ft.dY[-1], dy_plain, ft.Y[-1]):
            assert torch.equal(t, r * torch.where(y3 > 0, torch.ones_like(y3), y3 + 1)), "dZ3 != dY3 * act'(Y3)"
        assert torch.equal(m3, m2)
        for n in head_names:
            assert torch.equal(dict(model.named_parameters())[n].grad, grads_plain[n]), f"head grad {n} changed with dz_act"


def test_staged_host_upload_round_trip():
    """ExperienceBuffer.stage_from_host / commit_staged (double-buffered upload of a host-resident rollout): the committed
    storage equals the host rollout bit for bit, storage addresses do not move, and a second staged rollout does not
    disturb the committed one until it is committed itself."""
    from mineral_b200.agents.ppo.experience import ExperienceBuffer

    T, N, O, A = 8, 96, 37, 8
    buf = ExperienceBuffer(N, T, N * T, 256, {"obs": (O,)}, A, DEV)
    ptrs = {k: (v["obs"] if k == "obses" else v).data_ptr() for k, v in buf.storage_dict.items()}
    rollouts = []
    for seed in (0, 1):
        st, _ = oc.synth_rollout(T, N, O, A, seed=seed)
        host = {k: v.pin_memory() for k, v in st.items() if k not in ("returns", "obses")}
        host["obs"] = st["obses"]["obs"].pin_memory()
        rollouts.append(host)
    buf.stage_from_host(rollouts[0])
    buf.commit_staged()
    buf.stage_from_host(rollouts[1])          # in flight / landed in the staging buffers only
    torch.cuda.synchronize()
    for k, v in rollouts[0].items():
        eq(buf.storage_dict["obses"]["obs"] if k == "obs" else buf.storage_dict[k], v, f"first rollout {k}")
    buf.commit_staged()
    torch.cuda.synchronize()
    for k, v in rollouts[1].items():
        eq(buf.storage_dict["obses"]["obs"] if k == "obs" else buf.storage_dict[k], v, f"second rollout {k}")
    assert ptrs == {k: (v["obs"] if k == "obses" else v).data_ptr() for k, v in buf.storage_dict.items()}


def test_ppo_eval_runs_and_accounts_episodes():
    """PPO.eval (ppo.py:352-408): rolls the sampling policy until 2 x num_actors episodes finished; the device-side episode
    accounting must agree with a host recount of the same done / reward streams."""
    import json
    import os

    from mineral_b200.agents import PPO
    from oracle.reference_shim import SynthEnv, to_attr

    c = load_golden("ppo_e2e")[0]
    T, N, O, A = c["shape"]
    cfg = to_attr(c["cfg"])
    cfg.rl_device = DEV
    cpu_env = SynthEnv(N, O, A, 400, seed=7, done_p=0.1, timeout_p=0.0)
    log = []

    class OnDevice:
        observation_space, action_space = cpu_env.observation_space, cpu_env.action_space

        def reset(self):
            return cu(cpu_env.reset())

        def step(self, actions):
            assert actions.shape == (N, A) and float(actions.abs().max()) <= 1.0     # clamped actions reach the simulator
            o, r, d, info = cpu_env.step(actions)
            log.append((r.clone(), d.clone()))
            return cu(o), cu(r), cu(d), cu(info)

    agent = PPO(cfg, env=OnDevice())
    agent.model.load_state_dict(c["init_state_dict"])
    m = agent.eval()
    # host recount (Metrics._update_scores, mineral/common/metrics.py:59-67)
    cur_r, cur_l, ep_r, ep_l = torch.zeros(N), torch.zeros(N), [], []
    for r, d in log:
        cur_r += r
        cur_l += 1
        for i in torch.where(d)[0].tolist():
            ep_r.append(cur_r[i].item()), ep_l.append(cur_l[i].item())
            cur_r[i], cur_l[i] = 0, 0
    assert m["eval_scores/num_episodes"] == len(ep_r) >= 2 * N
    W = 2 * N
    scale_close(torch.tensor(m["eval_scores/episode_rewards"]), torch.tensor(ep_r[-W:]).mean() if len(ep_r) >= W else torch.tensor(sum(ep_r) / W), 1e-5, "mean reward")
    scale_close(torch.tensor(m["eval_scores/episode_lengths"]), torch.tensor(ep_l[-W:]).mean(), 1e-6, "mean length")
    sc = json.load(open(os.path.join(agent.logdir, "scores.json")))
    assert sc["eval_scores/num_episodes"] == len(ep_r) and len(sc["eval_scores/episode_rewards"]) == W
