"""2-GPU equivalence: an env-sharded 2-rank PPO update equals the 1-rank update on the
concatenated rollout whose minibatch i is the union of the ranks' local minibatch i (SURVEY.md
section 8e).  Needs >= 2 visible GPUs (`gpurun --gpus 2`); skipped otherwise."""

import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import scale_close

pytestmark = pytest.mark.gpu

T, NL, O, A, NUM_MB, EPOCHS = 8, 64, 11, 3, 4, 2   # NL envs per rank


def _cfg(n_envs, mb, device, multi_gpu, graph=False):
    from oracle.reference_shim import to_attr

    return to_attr(dict(
        rl_device=device, multi_gpu=multi_gpu, task=dict(env_autoresets=True),
        agent=dict(network=dict(normalize_input=True, actor_critic="ActorCritic",
                                actor_critic_kwargs=dict(mlp_kwargs=dict(units=[32, 16], norm_type="LayerNorm", act_type="SiLU"), fixed_sigma=False,
                                                         actor_dist_kwargs=dict(dist_type="dreamerv3_normal", minstd=0.1, maxstd=1.0))),
                   ppo=dict(num_actors=n_envs, reward_shaper=dict(fn="scale", scale=1.0), max_agent_steps=1e9, horizon_len=T, minibatch_size=mb,
                            normalize_value=True, value_bootstrap=True, clip_value_loss=True, normalize_advantage=True, gamma=0.99, tau=0.95,
                            entropy_coef=0.001, e_clip=0.2, use_smooth_clamp=False, critic_coef=4, bounds_loss_coef=1e-4, bounds_type="bound",
                            mini_epochs=EPOCHS, optim_type="AdamW", optim_kwargs=dict(lr=5e-4, eps=1e-5), lr_schedule="kl", kl_threshold=0.008,
                            max_grad_norm=0.5, truncate_grads=True, cuda_graph=graph))))


def _load(agent, st):
    for k, v in st.items():
        if k == "obses":
            agent.storage.storage_dict["obses"]["obs"].copy_(v["obs"])
        elif k != "returns":
            agent.storage.storage_dict[k].copy_(v)


def _take(st, last, envs):
    out = {k: ({"obs": v["obs"][:, envs]} if k == "obses" else v[:, envs]) for k, v in st.items()}
    return out, last[envs]


def _run(agent, last, repeats=1):
    snap = _snapshot(agent)
    for rep in range(repeats):   # repeats > 1 (CUDA-graph mode): warm-up, capture, replay -- each from the same initial state
        if rep > 0:
            _restore(agent, snap)
        agent.finish_rollout(last.to(agent.device))
        agent.set_train()
        res = agent.train_epoch()
        torch.cuda.synchronize()
    return dict(loss=torch.stack(res["loss/total"]).cpu(), kl=torch.stack(res["avg_kl"]).cpu(), gn=torch.stack(res["grad_norm/all"]).cpu(),
                params=agent.optim.flat.cpu(), obs_mean=agent.obs_rms["obs"].running_mean.cpu(), obs_var=agent.obs_rms["obs"].running_var.cpu(),
                val_mean=agent.value_rms.running_mean.cpu(), val_var=agent.value_rms.running_var.cpu(), lr=agent.last_lr)


def _snapshot(agent):
    sd = agent.storage.storage_dict
    return dict(flat=agent.optim.flat.clone(), m=agent.optim.exp_avg.clone(), v=agent.optim.exp_avg_sq.clone(), st=agent.optim.state.clone(),
                obs={k: v.clone() for k, v in agent.obs_rms["obs"].state_dict().items()},
                val={k: v.clone() for k, v in agent.value_rms.state_dict().items()},
                mu=sd["mu"].clone(), sigma=sd["sigma"].clone())


def _restore(agent, snap):
    sd = agent.storage.storage_dict
    agent.optim.flat.copy_(snap["flat"]), agent.optim.exp_avg.copy_(snap["m"]), agent.optim.exp_avg_sq.copy_(snap["v"])
    agent.optim.state.copy_(snap["st"])
    for name, rms in (("obs", agent.obs_rms["obs"]), ("val", agent.value_rms)):
        for k, v in snap[name].items():
            getattr(rms, k).copy_(v)
    sd["mu"].copy_(snap["mu"]), sd["sigma"].copy_(snap["sigma"])


def _worker(rank, world, port, init_sd, ret, graph):
    from mineral_b200.agents import PPO
    from oracle import mineral_oracle as oc
    from oracle.reference_shim import SynthEnv

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    try:
        st, last = oc.synth_rollout(T, NL * world, O, A, seed=3, done_p=0.05)
        envs = torch.arange(rank * NL, (rank + 1) * NL)
        st_r, last_r = _take(st, last, envs)
        agent = PPO(_cfg(NL, NL * T // NUM_MB, f"cuda:{rank}", True, graph), env=SynthEnv(4, O, A, 1))
        assert agent.comm is not None
        agent.model.load_state_dict(init_sd)
        _load(agent, st_r)
        ret[rank] = _run(agent, last_r, repeats=3 if graph else 1)
        assert agent.comm.error() == 0
        assert agent._engine.graph is not None or not graph
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("graph", [False, True])
def test_two_rank_update_equals_single_rank_on_union(graph):
    """Gradient / metrics / statistics exchanged by our own kernels over NVLink peer memory (csrc/peer.cu), optionally with
    the whole update replayed as one CUDA graph."""
    from mineral_b200.agents import PPO
    from oracle import mineral_oracle as oc
    from oracle.reference_shim import SynthEnv

    world = 2
    torch.manual_seed(0)
    single = PPO(_cfg(NL * world, world * NL * T // NUM_MB, "cuda:0", False), env=SynthEnv(4, O, A, 1))
    init_sd = {k: v.detach().cpu().clone() for k, v in single.model.state_dict().items()}

    st, last = oc.synth_rollout(T, NL * world, O, A, seed=3, done_p=0.05)
    # env order of the single-rank run: minibatch i = [rank0 slice i, rank1 slice i]
    E = NL // NUM_MB
    order = torch.cat([torch.arange(r * NL + i * E, r * NL + (i + 1) * E) for i in range(NUM_MB) for r in range(world)])
    st_u, last_u = _take(st, last, order)
    _load(single, st_u)
    ref = _run(single, last_u)

    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, 29571 + int(graph), init_sd, ret, graph), nprocs=world, join=True)
    for r in range(world):
        got = ret[r]
        scale_close(got["loss"], ref["loss"], 1e-5, f"rank{r} losses")
        scale_close(got["gn"], ref["gn"], 1e-5, f"rank{r} grad norms")
        assert (got["kl"] - ref["kl"]).abs().max().item() <= 1e-5 * ref["kl"].abs().max().item() + 2e-7
        assert got["lr"] == ref["lr"]
        scale_close(got["params"], ref["params"], 1e-5, f"rank{r} params")
        for k in ("obs_mean", "obs_var", "val_mean", "val_var"):
            scale_close(got[k], ref[k], 1e-6, f"rank{r} {k}")
    assert torch.equal(ret[0]["params"], ret[1]["params"]), "ranks diverged"
