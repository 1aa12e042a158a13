"""Pin the CPU oracle (oracle/mineral_oracle.py) against vectors produced by the
unmodified reference (oracle/make_golden.py -> tests/golden/*.pt).  CPU only."""

import numpy as np
import torch

from conftest import load_golden, scale_close
from oracle import mineral_oracle as oc


def test_gae_and_pack_bit_exact():
    for c in load_golden("experience"):
        st = {k: (dict(v) if isinstance(v, dict) else v.clone()) for k, v in c["inputs"].items()}
        ret, _ = oc.gae_scan(st["rewards"], st["values"], st["dones"], c["last_values"], c["gamma"], c["tau"])
        assert torch.equal(ret, c["returns_time_major"])
        st["returns"] = ret
        data = oc.pack_rollout(st)
        for k, v in c["data"].items():
            if isinstance(v, dict):
                for kk in v:
                    assert torch.equal(data[k][kk], v[kk]), (k, kk)
            else:
                assert data[k].shape == v.shape and data[k].dtype == v.dtype, k
                assert torch.equal(data[k], v), k
        s, e = oc.minibatch_rows(1, c["minibatch_size"])
        assert torch.equal(data["advantages"][s:e], c["minibatch1"]["advantages"])
        assert torch.equal(data["obses"]["obs"][s:e], c["minibatch1"]["obses"]["obs"])
        assert c["length"] == data["values"].shape[0] // c["minibatch_size"]


def test_rms_bit_exact():
    for c in load_golden("rms"):
        kw = c["kwargs"]
        dtype = torch.float64 if kw.get("dtype") == "torch.float64" else torch.float32
        st = oc.rms_init(c["shape"], kw["eps"], kw.get("initial_count", "eps"), dtype)
        clamp = (-5.0, 5.0) if kw.get("with_clamp") else None
        for s in c["steps"]:
            oc.rms_update(st, s["x"], kw.get("correction", 1))
            assert torch.equal(st[0], s["mean"]) and torch.equal(st[1], s["var"]) and torch.equal(st[2], s["count"])
            y = oc.rms_normalize(st, s["x"], kw["eps"], clamp)
            assert torch.equal(y, s["normalized"])
            assert torch.equal(oc.rms_unnormalize(st, y * 1.3, kw["eps"], clamp), s["unnormalized"])
        assert c["state_dict_keys"] == ["running_mean", "running_var", "running_count"]


def test_td_lambda_bit_exact():
    for c in load_golden("tdlambda"):
        if c["method"] == "td-lambda":
            out = oc.td_lambda_targets(c["rew"], c["done_mask"], c["next_values"], c["gamma"], c["lam"])
        else:
            out = oc.one_step_targets(c["rew"], c["next_values"], c["gamma"])
        assert torch.equal(out, c["target"])
        for v in c["entropy_variants"]:
            ent = oc.shac_entropy_term(c["distr_ent"], c["logprobs"], v["use_distr_ent"], c["target_entropy"], v["offset"], v["scale"])
            if c["method"] == "td-lambda":
                out = oc.td_lambda_targets(c["rew"], c["done_mask"], c["next_values"], c["gamma"], c["lam"], v["alpha"], ent)
            else:
                out = oc.one_step_targets(c["rew"], c["next_values"], c["gamma"], v["alpha"], ent)
            assert torch.equal(out, v["target"])


def test_critic_dataset_slices():
    for c in load_golden("critic_dataset"):
        obs = c["obs"]["obs"].reshape(-1, c["obs"]["obs"].shape[-1])
        tv = c["target_values"].reshape(-1)
        if c["shuffle"]:
            np.random.seed(c["np_seed"])
            perm = torch.from_numpy(np.random.permutation(tv.shape[0]))
            obs, tv = obs[perm], tv[perm]
        sl = oc.critic_dataset_slices(tv.shape[0], c["batch_size"], c["drop_last"])
        assert len(sl) == c["length"]
        for (s, e), b in zip(sl, c["batches"]):
            assert torch.equal(obs[s:e], b["obs"]["obs"]) and torch.equal(tv[s:e], b["target"])


def test_replay_ring_bit_exact():
    g = load_golden("replay")
    ring = oc.RingOracle(g["obs_dim"], g["act_dim"], g["capacity"])
    for ev in g["events"]:
        i = ev["insert"]
        ring.insert(i["obs"], i["act"], i["rew"], i["nobs"], i["done"])
        assert (ring.next_p, ring.full, ring.cur) == (ev["next_p"], ev["if_full"], ev["cur_capacity"])
        for name, buf in (("obs", ring.obs), ("act", ring.act), ("rew", ring.rew), ("nobs", ring.nobs), ("done", ring.done)):
            assert torch.equal(buf, ev["buf"][name]), name
        out = ring.gather(ev["idx"])
        for name, t in zip(("obs", "act", "rew", "nobs", "done"), out):
            assert torch.equal(t, ev["sample"][name]) and t.dtype == torch.float32, name


def test_nstep_bit_exact():
    for c in load_golden("nstep"):
        ns = oc.NStepOracle(c["N"], c["O"], c["A"], c["nstep"], c["gamma"])
        for call in c["calls"]:
            i = call["inp"]
            if call["out"] is None:  # reference raises while warming up; the oracle returns nothing
                try:
                    ns.push(i["obs"], i["act"], i["rew"], i["nobs"], i["done"])
                except (RuntimeError, ValueError):
                    pass
                continue
            out = ns.push(i["obs"], i["act"], i["rew"], i["nobs"], i["done"])
            for name, t in zip(("obs", "act", "rew", "nobs", "done"), out):
                assert torch.equal(t, call["out"][name]), (c["nstep"], name)


def test_losses_bit_exact_and_grads():
    for c in load_golden("losses"):
        nlp = c["nlp"].clone().requires_grad_(True)
        v = c["v"].clone().requires_grad_(True)
        mu = c["mu"].clone().requires_grad_(True)
        a, frac = oc.actor_loss(c["old_nlp"], nlp, c["adv"], c["e_clip"], c["smooth"])
        cl, ev = oc.critic_loss(c["vold"], v, c["e_clip"], c["ret"], c["clip_value"])
        b = oc.bounds_loss(mu, c["bounds_type"])
        kl = oc.policy_kl(mu.detach(), c["sigma"], c["old_mu"], c["old_sigma"])
        assert torch.equal(a.detach(), c["a_loss"]) and torch.equal(frac, c["clip_frac"])
        assert torch.equal(cl.detach(), c["c_loss"]) and torch.equal(ev, c["explained_var"])
        assert torch.equal(b.detach(), c["b_loss"]) and torch.equal(kl, c["kl"])
        (a.mean() + 0.5 * cl.mean() * 4 + b.mean() * 1e-4).backward()
        assert torch.equal(nlp.grad, c["g_nlp"]) and torch.equal(v.grad, c["g_v"]) and torch.equal(mu.grad, c["g_mu"])


def _hp(c):
    p = c["cfg"]["agent"]["ppo"]
    hp = {k: p[k] for k in ("gamma", "tau", "e_clip", "critic_coef", "entropy_coef", "bounds_loss_coef", "bounds_type", "clip_value_loss",
                            "use_smooth_clamp", "mini_epochs", "minibatch_size", "truncate_grads", "max_grad_norm", "lr_schedule", "kl_threshold")}
    ack = c["cfg"]["agent"]["network"]["actor_critic_kwargs"]
    mk, dk = ack["mlp_kwargs"], ack["actor_dist_kwargs"]
    net = dict(n_layers=len(mk["units"]), norm=mk.get("norm_type") is not None, act=mk["act_type"], fixed_sigma=ack["fixed_sigma"],
               dist_type=dk["dist_type"], minstd=dk.get("minstd", 1.0), maxstd=dk.get("maxstd", 1.0))
    return hp, net, p


def test_ppo_update_matches_reference_agent():
    """The whole play_steps tail + train_epoch of the real reference agent, replayed by the oracle
    from the recorded storage and initial parameters.  Same torch ops in the same order => exact."""
    for c in load_golden("ppo_e2e"):
        hp, net, p = _hp(c)
        params = {k: v.clone().requires_grad_(True) for k, v in c["init_state_dict"].items()}
        Opt = getattr(torch.optim, p["optim_type"])
        optim = Opt(list(params.values()), **p["optim_kwargs"])
        st = {k: (dict(v) if isinstance(v, dict) else v.clone()) for k, v in c["storage"].items()}
        vr = [c["value_rms_before"][k].clone() for k in ("running_mean", "running_var", "running_count")]
        orms = [c["obs_rms_before"][f"obs.{k}"].clone() for k in ("running_mean", "running_var", "running_count")]
        data, res, lr = oc.ppo_update(st, c["last_values"], params, optim, orms, vr, hp, net, p["optim_kwargs"]["lr"])
        ref = c["results"]
        for mine, theirs in (("loss", "loss/total"), ("a_loss", "loss/actor"), ("c_loss", "loss/critic"), ("b_loss", "loss/bounds"),
                             ("entropy", "loss/entropy"), ("clip_frac", "clip_frac"), ("explained_var", "explained_var"), ("grad_norm", "grad_norm/all"), ("avg_kl", "avg_kl")):
            scale_close(torch.stack(res[mine]), ref[theirs], 1e-6, f"{c['variant']} {mine}")
        assert lr == c["last_lr"]
        for k, v in c["final_state_dict"].items():
            scale_close(params[k].detach(), v, 1e-6, k)
        for i, k in enumerate(("running_mean", "running_var", "running_count")):
            scale_close(orms[i], c["obs_rms_after"][f"obs.{k}"], 1e-12, k)
            scale_close(vr[i], c["value_rms_after"][k], 1e-12, k)
        scale_close(data["mu"], c["data_after_train"]["mu"], 1e-6)
        scale_close(data["sigma"], c["data_after_train"]["sigma"], 1e-6)


def test_shac_update_critic_matches_reference():
    """a16: oracle restatement of SHAC.update_critic == the reference's own run (Critic and EnsembleCritic)."""
    for c in load_golden("shac_critic"):
        T, B, O = c["shape"]
        obs, tv = oc.shac_critic_inputs(c["seed"], T, B, O, c["wild"])
        params = {k: v.clone().requires_grad_(True) for k, v in c["init_state_dict"].items()}
        prefixes = ["critic_mlp"] if c["n_critics"] == 1 else [f"critics.{k}" for k in range(c["n_critics"])]
        optim = torch.optim.Adam(list(params.values()), lr=c["lr"], betas=c["betas"])
        res = oc.shac_update_critic(params, prefixes, optim, obs.view(-1, O), tv.view(-1), c["batch_size"], c["iterations"], 2, True, "ELU",
                                    True, c["max_grad_norm"])
        for k, v in c["results"].items():
            assert torch.equal(torch.stack(res[k]), v), k
        for k, v in c["final_state_dict"].items():
            assert torch.equal(params[k].detach(), v), k


def test_offpolicy_math_matches_reference():
    """8(f)-2: oracle restatements of soft_update / distl_projection / the SAC critic target == the reference's own."""
    c = load_golden("offpolicy")
    su = c["soft_update"]
    for k in su["target_before"]:
        assert torch.equal(oc.soft_update(su["target_before"][k], su["current"][k], su["tau"]), su["target_after"][k]), k
    d = c["distl"]
    assert torch.equal(oc.distl_projection(d["next_dist"], d["reward"], d["done"], d["gamma"], d["v_min"], d["v_max"], d["num_atoms"],
                                           d["support"]), d["proj"])
    s = c["sac_target"]
    for case in s["cases"]:
        got = oc.td_target(s["q"], s["reward"], s["done"], case["gamma"], case["nstep"], case["alpha"] if case["backup_entropy"] else None,
                           s["log_prob"])
        assert torch.equal(got, case["target"]), case


def test_sac_policy_head_matches_reference_autograd():
    """The squashed-normal sample / log-prob restatement and its CLOSED-FORM gradient against the reference's own
    Dist + SquashedNormal run under autograd (golden sac_policy.pt)."""
    for c in load_golden("sac_policy"):
        kw = c["dist_kwargs"]
        act, lp = oc.squashed_sample(c["mu"], c["logstd"], c["eps"], **kw)
        scale_close(act, c["actions"], 1e-6, "actions")
        scale_close(lp, c["log_prob"], 1e-6, "log_prob")
        sigma, _ = oc.squashed_sigma(c["logstd"], **kw)
        assert torch.equal(sigma, c["sigma"])
        B = c["mu"].shape[0]
        g_mu, g_ls = oc.squashed_bwd(c["mu"], c["logstd"], c["eps"], c["g_act"], c["alpha"] / B, **kw)
        # autograd forms d/dmu as (-d/sigma^2 + ...) + d/sigma^2: two terms of size c*eps/sigma (up to 1e3 x the result when
        # sigma = e^-5) that cancel, so the REFERENCE's f32 gradient carries their rounding (the closed form is ~100x closer
        # to the f64 value).  Tolerance: 1e-5 of scale + 2 ulp of the cancelling terms.
        floor = (c["alpha"] / B * c["eps"] / sigma).abs().max().item() * 2.0 ** -24
        for name, mine, ref, extra in (("g_mu", g_mu, c["g_mu"], 2 * floor), ("g_logstd", g_ls, c["g_logstd"], 2 * floor * sigma.max().item())):
            err = (mine - ref).abs().max().item()
            assert err <= 1e-5 * ref.abs().max().item() + extra, f"{name}: {err:.3e}"
        d = {k: v.double() for k, v in c.items() if torch.is_tensor(v)}   # and against the same formula in f64
        t_mu, t_ls = oc.squashed_bwd(d["mu"], d["logstd"], d["eps"], d["g_act"], c["alpha"] / B, **kw)
        scale_close(g_mu, t_mu, 2e-6, "g_mu vs f64")
        scale_close(g_ls, t_ls, 2e-6, "g_logstd vs f64")


def test_ddpg_noise_and_actor_objectives_match_reference():
    """oc.tanh_action against the reference's add_normal_noise / add_mixed_normal_noise (its own torch.normal draw recorded),
    oc.actor_objective against the reference's expressions under autograd (first-minimum tie rule of torch.min included)."""
    c = load_golden("offpolicy_small")
    for n in c["noise"]:
        kw = n["kwargs"]
        got = oc.tanh_action(n["mu"], n["noise"], kw.get("noise_bounds"), kw.get("out_bounds"))
        assert torch.equal(got, n["actions"]), n["kind"]
    for o in c["objective"]:
        a_loss, ent, al_loss, seeds = oc.actor_objective([q.view(-1) for q in o["q"]], o["log_prob"].view(-1) if o["use_lp"] else None,
                                                         o["alpha"] if o["use_lp"] else 0.0, o["target_entropy"])
        scale_close(a_loss, o["actor_loss"], 1e-6, "actor_loss")
        for s_, r_ in zip(seeds, o["seeds"]):
            assert torch.equal(s_, r_.view(-1)), "seeds (argmin / tie rule)"
        if o["use_lp"]:
            scale_close(ent, o["entropy"], 1e-6, "entropy"), scale_close(al_loss, o["alpha_loss"], 1e-6, "alpha_loss")
