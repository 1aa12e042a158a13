"""Whole-update parity AT THE BENCHMARKED SHAPES: one full PPO update - GAE + env-major pack +
value / obs normalisation + mini_epochs x minibatches of optimizer steps - of

  * cfg2  Ant 4096 envs x 32 steps, 4 minibatches of 32768 rows, towers [256,128,64] ELU, fixed sigma, Adam      and
  * cfg5  AllegroHand 16384 envs x 16 steps, 8 minibatches of 32768 rows, towers [512,256,128] LayerNorm+SiLU,
          learned sigma (dreamerv3_normal), AdamW

through the agent API (persistent multi-tile tcgen05 GEMMs, split weight-gradient reduce, the 256-CTA fused heads kernel
with its grid barrier, the fused optimizer), eagerly and as one CUDA graph, against the CPU oracle's `ppo_update`
(mineral/agents/ppo/ppo.py:172-183,248-350 restated; pinned bit-exact / 1e-6 against the real reference agent by
tests/test_oracle_golden.py) on IDENTICAL inputs and initial parameters.

Tolerance: north_star's 1e-5 of scale.  A quantity may exceed it only by what the f32 REFERENCE itself is away from an
f64 evaluation of the same update (`slack` = 2 x max|ref_f32 - ref_f64|): after 20-40 optimizer steps both f32
evaluations drift from the exact trajectory by rounding, and asking ours to be closer to the f32 reference than the
f32 reference is to the truth would test the summation order, not the arithmetic.  Every measured error (ours vs f32,
ours vs f64, f32 vs f64) is written to gpurun_out/parity_fullshape.json (committed as profiles/r02_parity_fullshape.json)."""

import json
import os
import sys

import pytest
import torch

from oracle import mineral_oracle as oc

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_CACHE = {}
_REPORT = {}


def _bench():
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    import bench
    return bench


def _oracle_run(workload, dtype):
    """(initial params, results, data, final params, obs_rms, value_rms, lr list) of the CPU oracle in `dtype`."""
    key = (workload, dtype)
    if key in _CACHE:
        return _CACHE[key]
    b = _bench()
    w = b.WORKLOADS[workload]
    torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    params, optim, net, hp, obs_rms, value_rms = b.oracle_setup(w, seed=0)
    init = {k: v.detach().clone() for k, v in params.items()}
    st, last = oc.synth_rollout(w["T"], w["N"], w["O"], w["A"], seed=0)
    if dtype == torch.float64:
        params = {k: v.detach().double().requires_grad_(True) for k, v in params.items()}
        optim = type(optim)(list(params.values()), **{k: v for k, v in optim.defaults.items() if k in ("lr", "eps", "betas", "weight_decay")})
        st = {k: ({kk: vv.double() for kk, vv in v.items()} if isinstance(v, dict) else (v.double() if v.is_floating_point() else v))
              for k, v in st.items()}
        last = last.double()
    st_run = {k: (dict(v) if isinstance(v, dict) else v.clone()) for k, v in st.items()}
    data, res, lr = oc.ppo_update(st_run, last, params, optim, obs_rms, value_rms, hp, net, 5e-4)
    out = dict(init=init, res={k: (torch.stack(v) if torch.is_tensor(v[0]) else list(v)) for k, v in res.items()},
               adv=data["advantages"], values=data["values"], returns=data["returns"], mu=data["mu"], sigma=data["sigma"],
               final={k: v.detach().clone() for k, v in params.items()}, obs_rms=[t.clone() for t in obs_rms],
               value_rms=[t.clone() for t in value_rms], lr=lr, storage=st, last=last)
    _CACHE[key] = out
    return out


def _check(tag, name, ours, r32, r64, tol=1e-5, abs_floor=0.0):
    ours, r32, r64 = ours.detach().double().cpu().reshape(-1), r32.detach().double().reshape(-1), r64.detach().double().reshape(-1)
    scale = r32.abs().max().item()
    e32, e64, eref = (ours - r32).abs().max().item(), (ours - r64).abs().max().item(), (r32 - r64).abs().max().item()
    _REPORT.setdefault(tag, {})[name] = dict(scale=scale, ours_vs_f32=e32, ours_vs_f64=e64, f32_vs_f64=eref,
                                             rel_ours_vs_f32=e32 / max(scale, 1e-300), rel_f32_vs_f64=eref / max(scale, 1e-300))
    assert e32 <= tol * scale + 2.0 * eref + abs_floor, (
        f"{tag} {name}: |ours - ref_f32| = {e32:.3e} (scale {scale:.3e}, {e32 / max(scale, 1e-300):.2e} relative) exceeds "
        f"{tol:g} of scale + 2 x the f32 reference's own distance from f64 ({eref:.3e})")


@pytest.mark.parametrize("cuda_graph", [False, True])
@pytest.mark.parametrize("workload", ["ppo_ant_4096x32", "ppo_allegro_16384x16"])
def test_full_update_at_benchmark_shape_vs_oracle(workload, cuda_graph):
    from mineral_b200.agents import PPO
    from oracle.reference_shim import SynthEnv

    b = _bench()
    w = b.WORKLOADS[workload]
    T, N, O, A = w["T"], w["N"], w["O"], w["A"]
    r32, r64 = _oracle_run(workload, torch.float32), _oracle_run(workload, torch.float64)

    agent = PPO(b.make_cfg(w, DEV, False, cuda_graph=cuda_graph), env=SynthEnv(4, O, A, 1))
    tag = f"{workload} graph={cuda_graph}"
    sd = agent.storage.storage_dict
    snap_optim = agent.optim.state_dict()
    reps = 3 if cuda_graph else 1   # eager warm-up, capture + replay, replay: each from the same initial state
    for rep in range(reps):
        agent.model.load_state_dict(r32["init"])
        agent.optim.load_state_dict(snap_optim)
        for rms, shape in ((agent.obs_rms["obs"], (O,)), (agent.value_rms, (1,))):
            m, v, c = oc.rms_init(shape, 1e-5, 1, torch.float64)
            rms.running_mean.copy_(m), rms.running_var.copy_(v), rms.running_count.copy_(c)
        for k, v in r32["storage"].items():
            if k == "obses":
                sd["obses"]["obs"].copy_(v["obs"])
            elif k != "returns":
                sd[k].copy_(v)
        agent.finish_rollout(r32["last"].to(DEV))
        data = agent.storage.data_dict
        agent.set_train()
        res = agent.train_epoch()
        torch.cuda.synchronize()
        t = f"{tag} rep={rep}"
        # post-rollout pass: advantages / normalised values / returns of the WHOLE batch
        _check(t, "advantages", data["advantages"], r32["adv"], r64["adv"])
        _check(t, "values", data["values"], r32["values"], r64["values"])
        _check(t, "returns", data["returns"], r32["returns"], r64["returns"])
        # every optimizer step's logged scalars
        names = {"loss/total": "loss", "loss/actor": "a_loss", "loss/critic": "c_loss", "loss/bounds": "b_loss", "loss/entropy": "entropy",
                 "explained_var": "explained_var", "grad_norm/all": "grad_norm"}
        for ours_k, ref_k in names.items():
            _check(t, ours_k, torch.stack(res[ours_k]), r32["res"][ref_k], r64["res"][ref_k])
        # KL: a sum of O(1) terms cancelling to ~1e-3; rounding floor ~1e-7 * A absolute in any f32 evaluation order
        _check(t, "kl", torch.stack(res["kl"]), r32["res"]["kl"], r64["res"]["kl"], abs_floor=2e-7)
        _check(t, "avg_kl", torch.stack(res["avg_kl"]), r32["res"]["avg_kl"], r64["res"]["avg_kl"], abs_floor=2e-7)
        mb = T * N // w["minibatches"]
        frac_err = (torch.stack(res["clip_frac"]).cpu() - r32["res"]["clip_frac"]).abs().max().item()
        assert frac_err <= 2.01 / mb, f"{t} clip_frac differs by {frac_err * mb:.1f} rows"   # a count: <= 2 rows on the clip boundary
        # learning-rate decisions of every mini-epoch (KL schedule), exact
        lrs = agent._engine.lr_out.cpu().tolist()
        assert lrs == [float(x) for x in r32["res"]["lr"]], f"{t} lr per mini-epoch {lrs} vs {r32['res']['lr']}"
        assert agent.last_lr == r32["lr"]
        # statistics
        for i, k in enumerate(("running_mean", "running_var", "running_count")):
            _check(t, f"obs_rms.{k}", getattr(agent.obs_rms["obs"], k), r32["obs_rms"][i], r64["obs_rms"][i], tol=2e-6)
            _check(t, f"value_rms.{k}", getattr(agent.value_rms, k), r32["value_rms"][i], r64["value_rms"][i], tol=2e-6)
        # parameters after all optimizer steps.  Scale: the tensor's own magnitude PLUS what the optimizer could have moved it
        # (Adam moves an element by ~lr per step whatever the gradient's size, so a 1e-5 relative error in one step's direction
        # is 1e-5 * lr absolute): a bias that started at zero and oscillates around it has max|b| ~ 1e-4 and no meaningful
        # "relative to its own magnitude" (SURVEY.md section 0 makes the same point about element-wise rtol next to zero)
        msd = agent.model.state_dict()
        travel = sum(float(x) for x in r32["res"]["lr"]) * (w["minibatches"])   # sum over optimizer steps of that step's lr
        for k, v in r32["final"].items():
            _check(t, f"param {k}", msd[k], v, r64["final"][k], abs_floor=1e-5 * travel)
        _check(t, "mu write-back", data["mu"], r32["mu"], r64["mu"])
        _check(t, "sigma write-back", data["sigma"], r32["sigma"], r64["sigma"])
    assert (agent._engine.graph is not None) == cuda_graph
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "parity_fullshape.json"), "w") as f:
        json.dump(_REPORT, f, indent=1)
