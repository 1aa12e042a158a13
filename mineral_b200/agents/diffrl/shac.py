"""SHAC / SAPO: the critic-target and critic-regression half of mineral/agents/diffrl/shac.py.

On the B200 path: the (T,B) rollout buffers (`create_buffers`,
shac.py:209-224), `compute_target_values` / `compute_target_values_with_entropy` (shac.py:877-916) as
one scan kernel, `CriticDataset` and the critic minibatch regression loop (`update_critic`,
shac.py:824-875) on our own kernels: critic MLP forward / backward (nets/fused_towers.py: tcgen05 or SIMT
linear layers + LayerNorm/activation passes), the MSE of `compute_critic_loss` and its gradient
(`mnr_critic_mse_fwd_bwd`), grad.nan_to_num_ + clip_grad_norm_ + Adam in one launch (`mnr_fused_clip_adam`),
the whole critic_iterations x num_critic_batches loop replayed as ONE CUDA graph.  The actor update back-propagates through a differentiable simulator
(shac.py:632-822), which is third-party and not in the reference repo: `update_actor` raises.
"""

import collections

import torch
import torch.nn as nn
import torch.nn.functional as F

from ... import ops
from ...nets.fused_towers import FusedCritics
from ..agent import cfg_get
from ..ppo.ppo import FusedAdam
from .models import Critic, EnsembleCritic
from .utils import CriticDataset, grad_norm


class SHAC:
    def __new__(cls, *args, **kwargs):
        """Drop-in construction (mineral/scripts/run.py:140-144): `SHAC(full_cfg, logdir=, accelerator=, datasets=, env=)` with
        the reference's config keys (`full_cfg.agent.{network,shac}`, mineral/cfgs/agent/DFlexAnt{SHAC,SHAC2,SAPO}.yaml) is
        translated into the explicit-argument form below."""
        self = object.__new__(cls)
        if args and hasattr(args[0], 'agent'):
            self._from_cfg = cls._cfg_to_kwargs(args[0], kwargs.get('env'))
            self._cfg_ctx = (args[0], kwargs)
        return self

    @staticmethod
    def _cfg_to_kwargs(full_cfg, env):
        """shac.py:31-75,124-150,192-194: the keys the critic half reads, with the reference's defaults."""
        net, cfg = full_cfg.agent.network, full_cfg.agent.shac
        if env is None:
            raise ValueError('SHAC needs env= (num_envs / num_obs / num_actions are read from it, shac.py:38-41)')
        if cfg_get(net, 'encoder', None) is not None:
            raise NotImplementedError('encoders are outside the rollout->update hot path (SURVEY.md section 8)')
        critic_name = cfg_get(net, 'critic', 'Critic')
        if critic_name not in ('Critic', 'EnsembleCritic'):
            raise NotImplementedError(f'critic {critic_name}')
        ck = dict(cfg_get(net, 'critic_kwargs', {}) or {})
        ok = dict(cfg_get(cfg, 'critic_optim_kwargs', {}) or {})
        return dict(
            num_envs=env.num_envs, num_obs=env.num_obs, num_actions=env.num_actions, horizon_len=cfg.horizon_len, device=full_cfg.rl_device,
            gamma=cfg_get(cfg, 'gamma', 0.99), lam=cfg_get(cfg, 'lambda', 0.95), critic_method=cfg_get(cfg, 'critic_method', 'one-step'),
            critic_iterations=cfg_get(cfg, 'critic_iterations', 16), num_critic_batches=cfg_get(cfg, 'num_critic_batches', 4),
            num_critics=ck.get('n_critics', 1) if critic_name == 'EnsembleCritic' else 1,
            critic_mlp_kwargs=ck.get('mlp_kwargs', None), optim_type=cfg.optim_type, critic_lr=ok.get('lr', 1e-3),
            critic_betas=tuple(ok.get('betas', (0.9, 0.999))), max_grad_norm=cfg_get(cfg, 'max_grad_norm', None),
            truncate_grads=bool(cfg_get(cfg, 'truncate_grads', False)), with_logprobs=cfg_get(cfg, 'with_logprobs', False),
            use_distr_ent=cfg_get(cfg, 'use_distr_ent', False), entropy_coef=cfg_get(cfg, 'entropy_coef', None),
            offset_by_target_entropy=cfg_get(cfg, 'offset_by_target_entropy', False),
            scale_by_target_entropy=cfg_get(cfg, 'scale_by_target_entropy', False),
            target_entropy_scalar=cfg_get(cfg, 'target_entropy_scalar', 1.0), cuda_graph=cfg_get(cfg, 'cuda_graph', True))

    def __init__(self, *args, **kwargs):
        if getattr(self, '_from_cfg', None) is not None:
            full_cfg, kw = self._cfg_ctx
            self._init(**self._from_cfg)
            # the Agent-level attributes run.py and the writers read (agent.py:14-81)
            self.full_cfg, self.env, self.logdir = full_cfg, kw.get('env'), kw.get('logdir')
            self.network_config, self.shac_config = full_cfg.agent.network, full_cfg.agent.shac
            self.num_actors = self.shac_config.num_actors
            self.max_agent_steps = int(self.shac_config.max_agent_steps)
            self.epoch, self.mini_epoch, self.agent_steps = -1, -1, 0
        else:
            self._init(*args, **kwargs)

    def _init(self, num_envs, num_obs, num_actions, horizon_len, device, gamma=0.99, lam=0.95, critic_method='td-lambda',
                 critic_iterations=16, num_critic_batches=4, critic_units=(64, 64), num_critics=1, critic_lr=2e-3,
                 critic_betas=(0.7, 0.95), critic_mlp_kwargs=None, optim_type='Adam', max_grad_norm=1.0, truncate_grads=True,
                 with_logprobs=False, use_distr_ent=True, entropy_coef=None, offset_by_target_entropy=False,
                 scale_by_target_entropy=False, target_entropy_scalar=1.0, cuda_graph=True, fused=True):
        """Defaults follow mineral/cfgs/agent/DFlexAntSHAC.yaml:18-50 (critic [64,64] LayerNorm ELU, Adam lr 2e-3
        betas (0.7, 0.95), 16 iterations x 4 batches, max_grad_norm 1.0).  num_critics > 1 builds SAPO's EnsembleCritic."""
        self.device = device
        self.num_envs, self.num_obs, self.num_actions = num_envs, num_obs, num_actions
        self.action_dim = num_actions
        self.obs_space = {'obs': (num_obs,)}
        self.horizon_len = horizon_len
        self.gamma, self.lam, self.critic_method = gamma, lam, critic_method
        self.critic_iterations, self.num_critic_batches = critic_iterations, num_critic_batches
        self.critic_batch_size = num_envs * horizon_len // num_critic_batches  # shac.py:73
        self.max_grad_norm, self.truncate_grads = max_grad_norm, truncate_grads
        self.with_logprobs = with_logprobs
        self.use_distr_ent = use_distr_ent
        self.entropy_coef = entropy_coef
        self.with_autoent = False
        self.offset_by_target_entropy = offset_by_target_entropy
        self.scale_by_target_entropy = scale_by_target_entropy
        self.target_entropy = -self.action_dim * target_entropy_scalar  # shac.py:184-185
        mlp_kwargs = dict(units=list(critic_units), norm_type='LayerNorm', act_type='ELU') if critic_mlp_kwargs is None else dict(critic_mlp_kwargs)
        if num_critics == 1:
            self.critic = Critic(num_obs, num_actions, mlp_kwargs=mlp_kwargs).to(device)
        else:
            self.critic = EnsembleCritic(num_obs, num_actions, n_critics=num_critics, mlp_kwargs=mlp_kwargs).to(device)
        self.encoder = lambda obs: obs['obs']
        if optim_type not in ('Adam', 'AdamW'):
            raise NotImplementedError(f'optim_type {optim_type}: only Adam / AdamW have a fused kernel')
        # flat-buffer Adam(W) with device-resident scalars (mineral_b200/agents/ppo/ppo.py); shac.py:139-141
        self.critic_optim = FusedAdam(self.critic.parameters(), decoupled=(optim_type == 'AdamW'), lr=critic_lr, betas=tuple(critic_betas))
        self.use_cuda_graph = bool(cuda_graph)
        self.fused = bool(fused) and FusedCritics.parse(self.critic.mlps()) is not None
        self._critic_engine = None
        self.create_buffers(horizon_len, num_envs)

    def create_buffers(self, T, B):
        """shac.py:209-224 (done_mask is FLOAT here)."""
        f = dict(dtype=torch.float32, device=self.device)
        self.obs_buf = {k: torch.zeros((T, B) + v, **f) for k, v in self.obs_space.items()}
        self.action_buf = torch.zeros((T, B, self.action_dim), **f)
        self.rew_buf = torch.zeros((T, B), **f)
        self.done_mask = torch.zeros((T, B), **f)
        self.next_values = torch.zeros((T, B), **f)
        self.avg_next_values = torch.zeros((T, B), **f)
        self.target_values = torch.zeros((T, B), **f)
        self.ret = torch.zeros((B), **f)
        self.mus = torch.zeros((T, B, self.num_actions), **f)
        self.sigmas = torch.zeros((T, B, self.num_actions), **f)
        if self.with_logprobs:
            self.logprobs = torch.zeros((T, B), **f)
            self.distr_ent = torch.zeros((T, B), **f)

    @torch.no_grad()
    def compute_target_values(self):
        """shac.py:877-891: one-step or TD-lambda targets, one launch, bit-exact."""
        ops.critic_targets(self.rew_buf, self.done_mask, self.next_values, self.gamma, getattr(self, 'lam', 0.95),
                           self.critic_method, out=self.target_values)

    def get_alpha(self, scalar=True):
        return self.entropy_coef

    @torch.no_grad()
    def compute_target_values_with_entropy(self):
        """shac.py:893-916 (entropy-augmented reward), same kernel with the entropy term enabled."""
        ent = self.distr_ent if self.use_distr_ent else self.logprobs
        alpha = self.get_alpha(scalar=True) if self.with_autoent else self.entropy_coef
        ops.critic_targets(self.rew_buf, self.done_mask, self.next_values, self.gamma, getattr(self, 'lam', 0.95),
                           self.critic_method, alpha=alpha, entropy=ent, negate_entropy=not self.use_distr_ent,
                           offset_by_target_entropy=self.offset_by_target_entropy,
                           scale_by_target_entropy=self.scale_by_target_entropy, target_entropy=self.target_entropy,
                           out=self.target_values)

    def compute_critic_loss(self, obs, target_v):
        """shac.py:869-875: mean over critics of the MSE."""
        z = self.encoder(obs)
        pred_vs = self.critic(z, return_type='all')
        return torch.stack([F.mse_loss(p.squeeze(-1), target_v) for p in pred_vs]).mean()

    def update_critic(self, dataset=None):
        """shac.py:824-867: critic_iterations x num_critic_batches regression steps, all on our kernels (one CUDA graph
        once warm).  Returns the reference's results dict (per-iteration value_loss and mean grad norms)."""
        if dataset is None:
            dataset = CriticDataset(self.critic_batch_size, self.obs_buf, self.target_values)
        if not self.fused or list(dataset.obs) != ['obs']:
            return self._update_critic_autograd(dataset)
        if self._critic_engine is None:
            self._critic_engine = _CriticEngine(self)
        return self._critic_engine.run(dataset)

    def _update_critic_autograd(self, dataset):
        """Architectures FusedCritics does not cover: torch modules + autograd for the MLP, our optimizer kernel."""
        results = collections.defaultdict(list)
        opt = self.critic_optim
        dev = self.device
        for _ in range(self.critic_iterations):
            total, before, after = 0.0, [], []
            for i in range(len(dataset)):
                b_obs, b_t = dataset[i]
                loss = self.compute_critic_loss(b_obs, b_t)
                loss.backward()
                nb, na = torch.zeros(1, device=dev), torch.zeros(1, device=dev)
                opt.step(float(self.max_grad_norm) if (self.truncate_grads and self.max_grad_norm is not None) else 0.0,
                         grad_norm_out=nb, nan_to_num=True, clipped_norm_out=na)
                before.append(nb[0]), after.append(na[0])
                total = total + loss.detach()
            results['value_loss'].append(total / len(dataset))
            if self.truncate_grads:
                results['grad_norm_before_clip/critic'].append(torch.mean(torch.stack(before)))
                results['grad_norm_after_clip/critic'].append(torch.mean(torch.stack(after)))
        return results

    def update_actor(self):
        raise NotImplementedError('SHAC actor BPTT needs a differentiable simulator (dflex/warp), which is not part of the '
                                  'reference repository (SURVEY.md section 2 row 6)')


class _CriticEngine:
    """Pre-allocated buffers + the launch sequence of one `update_critic` call (optionally one CUDA graph)."""

    def __init__(self, agent):
        a = self.a = agent
        dev = a.device
        self.bs = a.critic_batch_size
        self.fc = FusedCritics(a.critic.mlps(), self.bs, dev)
        self.n_rows = a.num_envs * a.horizon_len
        self.n_batches = (self.n_rows - 1) // self.bs + 1
        f32 = dict(dtype=torch.float32, device=dev)
        self.losses = torch.zeros(a.critic_iterations, self.n_batches, **f32)
        self.norm_before = torch.zeros(a.critic_iterations, self.n_batches, **f32)
        self.norm_after = torch.zeros(a.critic_iterations, self.n_batches, **f32)
        # first layer input through a 16 B-multiple row stride (TMA); the copy is one strided torch copy per update
        self.in_dim = a.num_obs
        self.x_ld = (self.in_dim + 3) // 4 * 4
        self.x = torch.zeros(self.n_rows, self.x_ld, **f32)
        self.graph, self._warm, self._ptrs = None, 0, None

    def _enqueue(self, obs, targets):
        a, fc = self.a, self.fc
        max_norm = float(a.max_grad_norm) if (a.truncate_grads and a.max_grad_norm is not None) else 0.0
        if self.x_ld != self.in_dim:
            self.x[:, :self.in_dim].copy_(obs)
            x_all, ld = self.x, self.x_ld
        else:
            x_all, ld = obs, None
        for j in range(a.critic_iterations):
            for i in range(self.n_batches):
                s, e = i * self.bs, min((i + 1) * self.bs, self.n_rows)
                fc.loss_fwd_bwd(x_all[s:e], targets[s:e], self.losses[j, i:i + 1], ld)
                a.critic_optim.step(max_norm, grad_norm_out=self.norm_before[j, i:i + 1], nan_to_num=True,
                                    clipped_norm_out=self.norm_after[j, i:i + 1])

    def run(self, dataset):
        a = self.a
        obs, targets = dataset.obs['obs'], dataset.target_values
        assert obs.shape[0] == self.n_rows and obs.is_contiguous() and targets.is_contiguous()
        ptrs = (obs.data_ptr(), targets.data_ptr())
        with torch.no_grad():
            if a.use_cuda_graph:
                if self.graph is not None and ptrs != self._ptrs:   # the dataset moved (e.g. shuffled copy): re-capture
                    self.graph = None
                if self.graph is None and self._warm >= 1:
                    torch.cuda.synchronize()
                    self.graph = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(self.graph):
                        self._enqueue(obs, targets)
                    self._ptrs = ptrs
                if self.graph is not None:
                    self.graph.replay()
                else:
                    self._enqueue(obs, targets)
                    self._warm += 1
            else:
                self._enqueue(obs, targets)
        results = collections.defaultdict(list)
        for j in range(a.critic_iterations):
            results['value_loss'].append(self.losses[j].mean())
            if a.truncate_grads:
                results['grad_norm_before_clip/critic'].append(self.norm_before[j].mean())
                results['grad_norm_after_clip/critic'].append(self.norm_after[j].mean())
        return results
