"""Host-side construction of the agent plug-ins on CPU (no kernel runs): the reference's constructor signature and config
keys are accepted, the networks are routed to the fused engines, parameters live in the flat
optimizer buffers, and the first kernel call refuses CPU tensors (no fallback)."""
import pytest
import torch

from conftest import load_golden
from oracle.reference_shim import SynthEnv, to_attr


def test_shac_accepts_the_reference_constructor():
    from mineral_b200.agents import SHAC

    from mineral_b200.agents.agent import Agent
    from mineral_b200.agents.diffrl.shac import _ShapeEnv

    cfg = to_attr(dict(rl_device="cpu", multi_gpu=False, task=dict(env_autoresets=True), agent=dict(
        network=dict(normalize_input=True, tanh_clamp=True, actor="Actor",
                     actor_kwargs=dict(mlp_kwargs=dict(units=[128, 64, 32], norm_type="LayerNorm", act_type="ELU")),
                     critic="EnsembleCritic", critic_kwargs=dict(n_critics=2, mlp_kwargs=dict(units=[64, 64], norm_type="LayerNorm", act_type="ELU"))),
        shac=dict(num_actors=64, max_agent_steps=1e6, horizon_len=32, critic_method="td-lambda", **{"lambda": 0.9}, optim_type="AdamW",
                  actor_optim_kwargs=dict(lr=2e-3, betas=[0.7, 0.95]), critic_optim_kwargs=dict(lr=2e-3, betas=[0.7, 0.95]),
                  max_grad_norm=1.0, truncate_grads=True, with_logprobs=True, reward_shaper=dict(fn="scale", scale=1.0)))))
    a = SHAC(cfg, logdir=None, env=_ShapeEnv(64, 76, 21))
    assert isinstance(a, Agent) and a.ckpt_dir and (a.epoch, a.mini_epoch, a.agent_steps) == (-1, -1, 0)   # agent.py:60,79-81
    assert (a.critic_batch_size, a.lam, a.critic_method, a.num_actors) == (512, 0.9, "td-lambda", 64)   # shac.py:64-73
    assert type(a.critic).__name__ == "EnsembleCritic" and a.distr_ent.shape == (32, 64)
    assert a.critic_optim.betas == (0.7, 0.95) and a.critic_optim.decoupled
    # legacy (tanh_clamp) normaliser configuration, shac.py:77-80
    rms = a.obs_rms["obs"]
    assert (rms.correction, rms.running_mean.dtype, float(rms.running_count)) == (0, torch.float32, pytest.approx(1e-4))
    assert a.critic_target is not a.critic and a._critic_target_flat.numel() == a.critic_optim.flat.numel()
    b = SHAC.from_shapes(64, 76, 21, 32, "cpu", tanh_clamp=False, normalize_input=True)   # explicit form
    assert type(b.critic).__name__ == "Critic" and b.critic_batch_size == 512
    assert b.obs_rms["obs"].running_mean.dtype == torch.float64 and float(b.obs_rms["obs"].running_count) == 1.0   # shac.py:81-82
    # checkpoints carry the reference's keys (shac.py:969-982) and round-trip on the host
    import os
    import tempfile
    f = os.path.join(tempfile.mkdtemp(), "shac.pth")
    a.epoch, a.agent_steps = 7, 12345
    a.save(f)
    ck = torch.load(f, weights_only=False)
    assert set(ck) == {"epoch", "mini_epoch", "agent_steps", "obs_rms", "actor_encoder", "encoder", "actor", "critic", "encoder_target", "critic_target"}
    assert set(ck["obs_rms"]) == {"obs.running_mean", "obs.running_var", "obs.running_count"} and ck["actor_encoder"] is None
    a2 = SHAC(cfg, logdir=None, env=_ShapeEnv(64, 76, 21))
    a2.load(f)
    assert (a2.epoch, a2.agent_steps) == (7, 12345)
    for k, v in a.critic.state_dict().items():
        assert torch.equal(a2.critic.state_dict()[k], v) and torch.equal(a2.critic_target.state_dict()[k], a.critic_target.state_dict()[k])
    # the loaded parameters landed IN the flat optimizer buffers (views), not in fresh tensors
    p0 = next(a2.critic.parameters())
    assert p0.data.data_ptr() == a2.critic_optim.flat.data_ptr()
    with pytest.raises(NotImplementedError):
        a.update_actor()


@pytest.mark.parametrize("case_index", [0, 3])
def test_offpolicy_agents_construct_from_reference_config(case_index):
    from mineral_b200 import _native as nv
    from mineral_b200.agents import DDPG, SAC

    c = load_golden("offpolicy_e2e")[case_index]
    N, O, A = c["shape"]
    cfg = to_attr(c["cfg"])
    agent = (SAC if c["algo"] == "sac" else DDPG)(cfg, env=SynthEnv(N, O, A, 8, seed=1))
    agent.actor.load_state_dict(c["init"]["actor"])
    agent.critic.load_state_dict(c["init"]["critic"])
    # parameters are views into the optimizer's flat buffer (load_state_dict copied INTO it)
    p0 = next(agent.actor.parameters())
    assert p0.data.data_ptr() == agent.actor_optim.flat.data_ptr()
    assert torch.equal(agent.actor_optim.flat[:p0.numel()].view_as(p0), c["init"]["actor"]["actor_mlp.mlp.0.weight"])
    assert agent.critic_target_flat.numel() == agent.critic_optim.flat.numel()
    assert agent.x_ld % 4 == 0 and agent.x_ld >= O + A and agent._fq.in_dim == O + A and agent._fa.in_dim == O
    assert set(agent.actor.state_dict()) == set(c["init"]["actor"]) and set(agent.critic.state_dict()) == set(c["init"]["critic"])
    agent.obs = agent._convert_obs(agent.env.reset())
    with pytest.raises(nv.MineralB200Error, match="no CPU path"):
        agent.explore_env(agent.env, 1, random=True)
