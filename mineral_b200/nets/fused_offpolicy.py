"""SAC / DDPG networks on the hand-written kernels: explicit forward / backward of the reference's `Actor` and `EnsembleQ`
(mineral/agents/ddpg/models.py:8-111) through `TowerStack` (tcgen05 3xTF32 or SIMT linear layers + LayerNorm/activation
passes, nets/fused_towers.py), the 1-wide Q heads and the mu / sigma heads on the SIMT linear kernels.  No autograd graph:
the agents seed d loss / d output themselves (csrc/offpolicy.cu) and these classes push it back to every parameter's
`.grad` - or, for a frozen critic inside the actor update (sac.py:313-326, ddpg.py:299-310), only to its action input."""

import torch
import torch.nn as nn

from .. import ops
from .fused_towers import TowerStack, _same_shape, _tower_layers


def pad4(n):
    return (n + 3) // 4 * 4


class FusedQ:
    """EnsembleQ / DistributionalEnsembleQ with 1 or 2 critics: Q_k(cat(z, a)).  The input is read from a (rows, x_ld) buffer
    whose columns [0, obs_dim) hold z and [obs_dim, obs_dim + act_dim) the action (pad columns zero).  The last linear layer
    is n_out wide: 1 (scalar Q, mineral/agents/ddpg/models.py:70-111) or num_atoms (categorical critic, models.py:117-170)."""

    @staticmethod
    def parse(module):
        mlps = list(getattr(module, 'critics', []))
        if not 1 <= len(mlps) <= 2:
            return None
        parsed = [_tower_layers(m, plain_last=True) for m in mlps]
        if any(p is None for p in parsed) or any(not _same_shape(parsed[0], p) for p in parsed[1:]):
            return None
        return parsed

    def __init__(self, module, rows, device, use_tc=True):
        parsed = self.parse(module)
        if parsed is None:
            raise NotImplementedError('critic architecture outside the fused path (1-2 MLP critics of equal shape, Linear[-LayerNorm]-act blocks)')
        self.module, self.G, self.M = module, len(parsed), rows
        self.stack = TowerStack([p[0][:-1] for p in parsed], parsed[0][1], rows, device, use_tc)
        self.last = [p[0][-1][0] for p in parsed]
        self.n_out = n_out = self.last[0].weight.shape[0]
        self.in_dim, self.x_ld = self.stack.in_dim, pad4(self.stack.in_dim)
        f32 = dict(dtype=torch.float32, device=device)
        self.pred = [torch.empty(rows, n_out, **f32) for _ in range(self.G)]
        self.g_pred = [torch.empty(rows, n_out, **f32) for _ in range(self.G)]
        # d objective / d input per critic tower (tensor-core path: one launch, the consumer adds the towers) or summed (SIMT)
        self.dx = [torch.empty(rows, self.in_dim, **f32) for _ in range(self.G if self.stack.use_tc else 1)]
        h = self.stack.widths[-1]
        self.ws = torch.zeros(max(ops.lib.mnr_linear_bwd_weight_workspace_bytes(self.G, rows, n_out, h), 256), dtype=torch.uint8, device=device)
        self.mse_ws = torch.zeros(ops.lib.mnr_critic_mse_workspace_bytes(rows), dtype=torch.uint8, device=device)

    def forward(self, x, rows=None, split=True):
        """-> self.pred[k] (rows, n_out).  split=False: the weights did not change since the last forward of THIS object."""
        rows = self.M if rows is None else rows
        st, h = self.stack, self.stack.widths[-1]
        st.forward(x, self.x_ld, rows, split)
        ops.linear_fwd([dict(a=st.Y[-1][g], b=self.last[g].weight.data, c=self.pred[g], bias=self.last[g].bias.data, N=self.n_out, K=h)
                        for g in range(self.G)], rows, 0)
        return self.pred

    def backward_from_seeds(self, x, rows=None):
        """With d loss / d output of every critic in self.g_pred (rows, n_out): every critic parameter's .grad."""
        rows = self.M if rows is None else rows
        st, G, n = self.stack, self.G, self.n_out
        h = st.widths[-1]
        ops.linear_bwd_weight([dict(a=self.g_pred[g], b=st.Y[-1][g], c=self.last[g].weight.grad, bias_grad=self.last[g].bias.grad, N=n, K=h)
                               for g in range(G)], rows, 0, self.ws)
        ops.linear_bwd_input([dict(a=self.g_pred[g], b=self.last[g].weight.data, c=st.dY[-1][g], N=n, K=h) for g in range(G)], rows, 0)
        st.backward(x, rows)

    def loss_fwd_bwd(self, x, target, loss_out, split=True):
        """sum_k MSE(Q_k(x), target) (sac.py:301-303, ddpg.py:295-296) -> loss_out (1 float) and every critic parameter's .grad."""
        assert self.n_out == 1
        rows = target.numel()
        self.forward(x, rows, split)
        ops.critic_mse_sum_fwd_bwd([p.view(-1)[:rows] for p in self.pred], target.view(-1), [g_.view(-1)[:rows] for g_ in self.g_pred],
                                   loss_out, self.mse_ws)
        self.backward_from_seeds(x, rows)

    def input_grad(self, x, rows=None):
        """After forward(x) and with d objective / d Q_k in self.g_pred: d objective / d x -> self.dx, a list of (rows, in_dim)
        tensors whose SUM is the gradient (one per critic on the tensor-core path, a single summed one otherwise).  No parameter
        gradient is formed (the critic is frozen here: `critic.requires_grad_(False)`)."""
        rows = self.M if rows is None else rows
        st, h = self.stack, self.stack.widths[-1]
        ops.linear_bwd_input([dict(a=self.g_pred[g], b=self.last[g].weight.data, c=st.dY[-1][g], N=self.n_out, K=h) for g in range(self.G)], rows, 0)
        st.backward(x, rows, weight_grads=False, input_grad=self.dx if st.use_tc else self.dx[0])
        return self.dx


class FusedActor:
    """Actor: tower -> mu head [-> sigma head | sigma parameter] (mineral/agents/ddpg/models.py:8-71).  Outputs the RAW
    heads (mu before tanh, logstd before Dist); the distribution maths belongs to the squashed-normal / tanh kernels."""

    @staticmethod
    def parse(module):
        p = _tower_layers(module.actor_mlp)
        if p is None:
            return None
        if not module.tanh_policy and module.fixed_sigma is None:
            return None  # sigma passed in by the caller (`std=`): not used by the shipped SAC / DDPG recipes
        return p

    def __init__(self, module, rows, device, use_tc=True):
        parsed = self.parse(module)
        if parsed is None:
            raise NotImplementedError('actor architecture outside the fused path (Linear[-LayerNorm]-act blocks; sigma head or parameter)')
        self.module, self.M = module, rows
        self.stack = TowerStack([parsed[0]], parsed[1], rows, device, use_tc)
        self.in_dim = self.stack.in_dim
        self.A = module.mu.weight.shape[0]
        self.has_sigma = not module.tanh_policy
        self.sigma_param = self.has_sigma and isinstance(module.sigma, nn.Parameter)
        f32 = dict(dtype=torch.float32, device=device)
        self.mu_raw = torch.empty(rows, self.A, **f32)
        self.g_mu = torch.empty(rows, self.A, **f32)
        if self.has_sigma:
            self.ls = None if self.sigma_param else torch.empty(rows, self.A, **f32)
            self.g_ls = torch.empty(rows, self.A, **f32)
        h = self.stack.widths[-1]
        self.ws = torch.zeros(max(ops.lib.mnr_linear_bwd_weight_workspace_bytes(2, rows, self.A, h), 256), dtype=torch.uint8, device=device)
        if self.sigma_param:
            self.colsum_ws = torch.zeros(ops.lib.mnr_colsum_workspace_bytes(self.A), dtype=torch.uint8, device=device)

    @property
    def logstd(self):
        """(rows, A) head output, or the (A,) parameter (broadcast by the kernels) when sigma is fixed."""
        return self.module.sigma.data if self.sigma_param else self.ls

    def forward(self, x, x_ld, rows=None, split=True):
        rows = self.M if rows is None else rows
        m, st = self.module, self.stack
        h = st.widths[-1]
        st.forward(x, x_ld, rows, split)
        y = st.Y[-1][0]
        heads = [dict(a=y, b=m.mu.weight.data, c=self.mu_raw, bias=m.mu.bias.data, N=self.A, K=h)]
        if self.has_sigma and not self.sigma_param:
            heads.append(dict(a=y, b=m.sigma.weight.data, c=self.ls, bias=m.sigma.bias.data, N=self.A, K=h))
        ops.linear_fwd(heads, rows, 0)

    def backward(self, x, rows=None):
        """From self.g_mu (and self.g_ls): every actor parameter's .grad."""
        rows = self.M if rows is None else rows
        m, st = self.module, self.stack
        h = st.widths[-1]
        y = st.Y[-1][0]
        wg = [dict(a=self.g_mu, b=y, c=m.mu.weight.grad, bias_grad=m.mu.bias.grad, N=self.A, K=h)]
        head_sigma = self.has_sigma and not self.sigma_param
        if head_sigma:
            wg.append(dict(a=self.g_ls, b=y, c=m.sigma.weight.grad, bias_grad=m.sigma.bias.grad, N=self.A, K=h))
        ops.linear_bwd_weight(wg, rows, 0, self.ws)
        if self.sigma_param:
            ops.colsum(self.g_ls[:rows], m.sigma.grad, self.colsum_ws)
        ops.linear_bwd_input([dict(a=self.g_mu, b=m.mu.weight.data, c=st.dY[-1][0], N=self.A, K=h)], rows, 0)
        if head_sigma:
            ops.linear_bwd_input([dict(a=self.g_ls, b=m.sigma.weight.data, c=st.dY[-1][0], N=self.A, K=h)], rows, 0, accumulate=True)
        st.backward(x, rows)


def flatten_like(module, optim):
    """Re-point the parameters of `module` (a deepcopy of the module `optim` owns) into ONE flat f32 buffer with the optimizer's
    layout, so the Polyak update of the whole network is one launch (mnr_soft_update on the two flat buffers)."""
    params = list(module.parameters())
    assert len(params) == len(optim.params)
    flat = torch.zeros_like(optim.flat)
    for p, o in zip(params, optim.offsets):
        view = flat[o:o + p.numel()].view_as(p)
        view.copy_(p.data)
        p.data = view
    return flat
