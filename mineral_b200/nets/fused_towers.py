"""Explicit forward / backward of `ActorCritic._actor_critic` on the hand-written kernels: tcgen05 linear layers
(csrc/tc_linear.cu; SIMT twins in csrc/mlp.cu), LayerNorm + activation passes (csrc/norm_act.cu) and the fused
heads + loss kernel (csrc/heads_fused.cu).  No nn.Linear / nn.LayerNorm / activation module, no autograd graph on the
PPO update path.

Replaces, for the update, `MLP.forward` (mineral/nets/mlp.py:76-77) of both towers, the `mu` / `sigma` / `value`
heads (mineral/agents/ppo/models.py:136-151) and torch autograd's backward through them.  The module's parameters
stay ordinary `nn.Parameter`s (views into FusedAdam's flat buffers); this class only reads `.data` and writes `.grad`.

Supported towers (separate value tower, both towers of the same shape, no dropout):
  * Linear -> {ELU | ReLU | none}                      bias + activation fused into the GEMM epilogue, act' into the
                                                       backward GEMMs' operand load;
  * Linear -> LayerNorm -> {SiLU | ELU | ReLU | none}  the shipped PPO recipes (mineral/cfgs/agent/DFlexAntPPO.yaml:17-21,
                                                       AllegroHandPPO.yaml:18-28): GEMM (+bias) -> one LN+act pass;
                                                       backward: one LN+act backward pass (dgamma, dbeta, dZ in place) ->
                                                       GEMMs on dZ.
"""

import os

import torch
import torch.nn as nn

from .. import ops

_ACT = {nn.ELU: 1, nn.ReLU: 2, nn.SiLU: 3}


def _tower_layers(mlp, plain_last=False):
    """-> (list of (nn.Linear, nn.LayerNorm | None), act code) or None if the tower has an unsupported layer.
    plain_last: the final Linear carries neither norm nor activation (MLP(..., plain_last=True), mlp.py:55-56); it is
    returned as the last entry and excluded from the uniformity checks."""
    blocks, act = [], None
    for m in mlp.mlp:
        if isinstance(m, nn.Linear):
            blocks.append([m, None, False])
        elif isinstance(m, nn.LayerNorm):
            if not blocks or blocks[-1][1] is not None or blocks[-1][2] or not m.elementwise_affine or m.bias is None:
                return None
            if tuple(m.normalized_shape) != (blocks[-1][0].weight.shape[0],):
                return None
            blocks[-1][1] = m
        elif type(m) in _ACT:
            code = _ACT[type(m)]
            if isinstance(m, nn.ELU) and m.alpha != 1.0:
                return None
            if not blocks or blocks[-1][2] or act not in (None, code):
                return None
            act, blocks[-1][2] = code, True
        elif isinstance(m, nn.Identity):
            continue
        else:
            return None
    if not blocks:
        return None
    last = None
    if plain_last:
        last = blocks.pop()
        if last[1] is not None or last[2] or last[0].bias is None or not blocks:
            return None
    normed = [b[1] is not None for b in blocks]
    acted = [b[2] for b in blocks]
    if any(normed) != all(normed) or (act is not None and not all(acted)):
        return None
    if act == 3 and not all(normed):  # SiLU' needs the pre-activation: only behind a LayerNorm (recomputed from z)
        return None
    if any(b[0].bias is None for b in blocks):
        return None
    out = [(b[0], b[1]) for b in blocks]
    if last is not None:
        out.append((last[0], None))
    return out, (act or 0)


class TowerStack:
    """G (1 or 2) towers of identical shape run side by side: every launch serves all of them.
    towers[g] = list of (nn.Linear, nn.LayerNorm | None); all towers read the same input."""

    def __init__(self, towers, act, rows, device, use_tc=True):
        self.G, self.M, self.act = len(towers), rows, act
        assert 1 <= self.G <= 2
        self.lin = [[b[0] for b in t] for t in towers]    # [g][i]
        self.norm = [[b[1] for b in t] for t in towers]
        self.normed = self.norm[0][0] is not None
        # activation applied by the GEMM kernels themselves (epilogue / operand load); behind a LayerNorm it is 0
        self.gemm_act = 0 if self.normed else act
        self.L = len(self.lin[0])
        f32 = dict(dtype=torch.float32, device=device)
        widths = [l.weight.shape[0] for l in self.lin[0]]
        self.widths, self.in_dim = widths, self.lin[0][0].weight.shape[1]
        G = self.G
        # saved forward outputs and their gradients, per layer, per tower
        self.Y = [[torch.empty(rows, w, **f32) for _ in range(G)] for w in widths]
        self.dY = [[torch.empty(rows, w, **f32) for _ in range(G)] for w in widths]
        if self.normed:  # pre-normalisation GEMM outputs and the row statistics
            self.Z = [[torch.empty(rows, w, **f32) for _ in range(G)] for w in widths]
            self.mean = [[torch.empty(rows, **f32) for _ in range(G)] for _ in widths]
            self.rstd = [[torch.empty(rows, **f32) for _ in range(G)] for _ in widths]
            self.ln_ws = [ops.ln_act_bwd_workspace(G, rows, w, device) for w in widths]
        big = max(ops.lib.mnr_linear_bwd_weight_workspace_bytes(G, rows, *l.weight.shape) for l in self.lin[0])
        self.ws = torch.zeros(max(big, 256), dtype=torch.uint8, device=device)
        # tensor-core path (tcgen05 3xTF32, csrc/tc_linear.cu): needs widths % 16 == 0 (and whole 128-column tiles above
        # 128) and 16 B-multiple row strides, so W is kept pre-split (hi / lo) in K-padded buffers and the first layer's
        # input is read through a padded row stride.
        self.in_ld = (self.in_dim + 3) // 4 * 4
        self.use_tc = bool(use_tc) and ops.tc_supported() and all(w % 16 == 0 and (w <= 128 or w % 128 == 0) for w in widths)
        if self.use_tc:
            self.w_hi, self.w_lo, self.w_ld = [], [], []
            k = self.in_dim
            for w in widths:
                ld = (k + 3) // 4 * 4
                self.w_ld.append(ld)
                self.w_hi.append([torch.zeros(w, ld, **f32) for _ in range(G)])
                self.w_lo.append([torch.zeros(w, ld, **f32) for _ in range(G)])
                k = w
            # MNR_BWD_FORK=1: a layer's input-gradient and weight-gradient kernels run SIDE BY SIDE on disjoint SMs (two streams,
            # mnr_tc_grid_cap); MNR_BWD_SPLIT = share of the SMs the input-gradient kernel gets, per layer from the top
            self.bwd_fork = os.environ.get('MNR_BWD_FORK', '0') == '1'
            self.bwd_split = [float(v) for v in os.environ.get('MNR_BWD_SPLIT', '0.42,0.53,0.5,0.5').split(',')]
            self._side = None
            sms = ops.lib.mnr_sm_count()

            def ws_bytes(l):
                sizes = [ops.lib.mnr_tc_linear_bwd_weight_workspace_bytes(G, rows, *l.weight.shape)]
                if self.bwd_fork:   # the split count follows the CTA count: size for every cap this stack may launch under
                    for f in self.bwd_split:
                        ops.tc_grid_cap(sms - max(1, min(sms - 1, round(sms * f))))
                        sizes.append(ops.lib.mnr_tc_linear_bwd_weight_workspace_bytes(G, rows, *l.weight.shape))
                    ops.tc_grid_cap(0)
                return max(max(sizes), 256)
            # one split-partials workspace per layer: the split reduce of all layers is ONE launch after the last layer
            self.tc_ws = [torch.zeros(ws_bytes(l), dtype=torch.uint8, device=device) for l in self.lin[0]]
        # whole-tower forward kernel: three plain (Linear -> act) layers of the shapes csrc/tc_tower.cu takes
        self.fused_fwd = (self.use_tc and not self.normed and self.L == 3
                          and ops.tc_tower_fwd_supported(self.in_dim, widths, self.gemm_act))
        # set by the owner when the producer of dY[-1] already applied act'(Y[-1]) (the fused heads + loss kernel with
        # cfg.dz_act): the top layer's backward GEMMs then take dZ directly and load no saved-output tile
        self.top_is_dz = False

    def _split_weights(self):
        """W -> (hi, lo) tf32 parts for every layer of every tower, one launch per 8 matrices."""
        jobs = []
        for i in range(self.L):
            n, k = self.lin[0][i].weight.shape
            for g in range(self.G):
                jobs.append((self.lin[g][i].weight.data, self.w_hi[i][g], self.w_lo[i][g], n, k, k, self.w_ld[i]))
        for s in range(0, len(jobs), 8):
            ops.split_tf32(jobs[s:s + 8])

    def split_jobs(self):
        """[(weight Parameter, hi, lo, ld)] of every tensor-core layer (FusedAdam.attach_splits)."""
        if not self.use_tc:
            return []
        return [(self.lin[g][i].weight, self.w_hi[i][g], self.w_lo[i][g], self.w_ld[i]) for i in range(self.L) for g in range(self.G)]

    def forward(self, x, x_ld=None, rows=None, split=True):
        """x (rows, in_dim) [row stride x_ld floats] -> self.Y[-1][g] for every tower g.  split=False: the hi / lo weight
        copies are current (the weights did not change since the last forward, e.g. inside a rollout)."""
        M = self.M if rows is None else rows
        G = self.G
        x_ld = self.in_dim if x_ld is None else x_ld
        self._x_ld = x_ld
        tc0 = self.use_tc and x_ld % 4 == 0  # TMA needs 16 B-multiple row strides for the first layer's input
        if self.use_tc and split:
            self._split_weights()
        if tc0 and self.fused_fwd:
            # the whole tower in one launch: layer outputs travel through tensor memory (csrc/tc_tower.cu)
            ops.tc_tower_fwd([dict(x=x, ldx=x_ld, w_hi=[self.w_hi[i][g] for i in range(3)], w_lo=[self.w_lo[i][g] for i in range(3)],
                                   ldw=self.w_ld, bias=[self.lin[g][i].bias.data for i in range(3)], y=[self.Y[i][g] for i in range(3)])
                              for g in range(G)], M, self.in_dim, self.widths, self.gemm_act)
            return
        inp = [x] * G
        for i in range(self.L):
            n, k = self.lin[0][i].weight.shape
            lda = x_ld if i == 0 else k
            out = self.Z[i] if self.normed else self.Y[i]
            if self.use_tc and (i > 0 or tc0):
                ops.tc_linear_fwd([dict(a=inp[g], b=self.w_hi[i][g], b_lo=self.w_lo[i][g], c=out[g], bias=self.lin[g][i].bias.data, N=n, K=k,
                                        lda=lda, ldb=self.w_ld[i]) for g in range(G)], M, self.gemm_act)
            else:
                ops.linear_fwd([dict(a=inp[g], b=self.lin[g][i].weight.data, c=out[g], bias=self.lin[g][i].bias.data, N=n, K=k, lda=lda)
                                for g in range(G)], M, self.gemm_act)
            if self.normed:
                ops.ln_act_fwd([dict(z=self.Z[i][g], y=self.Y[i][g], gamma=self.norm[g][i].weight.data, beta=self.norm[g][i].bias.data,
                                     mean=self.mean[i][g], rstd=self.rstd[i][g]) for g in range(G)], M, n, self.norm[0][i].eps, self.act)
            inp = self.Y[i]

    def fold_begin(self):
        """Start deferring the cross-CTA sums of the launches that follow (the heads + loss kernel, LayerNorm's backward) into
        this step's split-reduce launch (ops.fold_defer; include/mineral_b200.h).  Returns the list to hand to backward(folds=),
        or None when this stack's backward does not end in exactly one split-reduce launch (SIMT layers, forked streams) or
        MNR_FOLD_DEFER=0.  The list object is reused: graph capture bakes its contents into the launch, eager calls refill it."""
        x_ld = getattr(self, '_x_ld', self.in_dim)
        ok = (self.use_tc and x_ld % 4 == 0 and self.L <= 3 and not getattr(self, 'bwd_fork', False)
              and (not self.normed or self.G * self.L + 1 <= 8) and os.environ.get('MNR_FOLD_DEFER', '1') != '0')
        if not ok:
            return None
        if getattr(self, '_folds', None) is None:
            from .. import _native as nv
            self._folds = nv.FoldList()
        ops.fold_defer(self._folds)
        return self._folds

    def backward(self, x, rows=None, weight_grads=True, input_grad=None, folds=None):
        """Layers top first, from self.dY[-1]; writes every .grad.  folds: the list fold_begin() returned - the deferred sums
        run inside the final split-reduce launch.  Without a norm, act'(.) is applied on the fly from the
        saved outputs inside the GEMMs; with LayerNorm one pass turns dY into dZ in place (and emits dgamma / dbeta).
        weight_grads=False skips the weight-gradient GEMMs (a frozen network that is only differentiated through, e.g. the
        critics inside SAC.update_actor, sac.py:313-326; LayerNorm's dgamma / dbeta are still written and are garbage to
        the caller).  input_grad: contiguous (rows, in_dim) tensor receiving d loss / d x summed over the towers, or a list of G
        such tensors, one per tower (tensor-core path, one launch)."""
        M = self.M if rows is None else rows
        G, top = self.G, self.L - 1
        x_ld = getattr(self, '_x_ld', self.in_dim)
        pending = []  # tensor-core dW layers whose split partials still wait for the (single) reduce launch
        # L2 hints (mnr_tc_cache_hints; MNR_TC_HINTS=0 disables): dX_i is the last reader of dY_i / Y_i and its result is read by the
        # next two launches; dW_0 is the last reader of everything it loads
        hints = self.use_tc and os.environ.get('MNR_TC_HINTS', '1') != '0'
        for i in range(top, -1, -1):
            n, k = self.lin[0][i].weight.shape
            src = self.Y[i - 1] if i > 0 else [x] * G
            ldb = x_ld if i == 0 else k
            tc = self.use_tc and ldb % 4 == 0
            if self.normed:
                ops.ln_act_bwd([dict(z=self.Z[i][g], gamma=self.norm[g][i].weight.data, beta=self.norm[g][i].bias.data, mean=self.mean[i][g],
                                     rstd=self.rstd[i][g], dy=self.dY[i][g], dz=self.dY[i][g], dgamma=self.norm[g][i].weight.grad,
                                     dbeta=self.norm[g][i].bias.grad) for g in range(G)], M, n, self.act, self.ln_ws[i])
            dz_in = self.normed or (i == top and self.top_is_dz)
            saved = [None] * G if dz_in else self.Y[i]
            act_i = 0 if dz_in else self.gemm_act
            wg = [dict(a=self.dY[i][g], b=src[g], c=self.lin[g][i].weight.grad, saved=saved[g], bias_grad=self.lin[g][i].bias.grad, N=n, K=k,
                       ldb=ldb) for g in range(G)]
            fork = self.use_tc and getattr(self, 'bwd_fork', False) and tc and weight_grads and i > 0
            if fork:   # dX_i on a side stream with its share of the SMs, dW_i on this stream with the rest; joined below
                sms = ops.lib.mnr_sm_count()
                f = self.bwd_split[min(top - i, len(self.bwd_split) - 1)]
                cap_dx = max(1, min(sms - 1, round(sms * f)))
                main = torch.cuda.current_stream()
                if self._side is None:
                    self._side = torch.cuda.Stream(device=self.dY[0][0].device)
                self._side.wait_stream(main)
                with torch.cuda.stream(self._side):
                    ops.tc_grid_cap(cap_dx)
                    ops.tc_linear_bwd_input([dict(a=self.dY[i][g], b=self.w_hi[i][g], b_lo=self.w_lo[i][g], c=self.dY[i - 1][g],
                                                  saved=saved[g], N=n, K=k, ldb=self.w_ld[i]) for g in range(G)], M, act_i)
                ops.tc_grid_cap(sms - cap_dx)
                pending.append(ops.tc_linear_bwd_weight_partials(wg, M, act_i, self.tc_ws[i]))
                ops.tc_grid_cap(0)
                main.wait_stream(self._side)
                if len(pending) == 4:
                    ops.tc_reduce_partials(pending)
                    pending = []
                continue
            if not weight_grads:
                pass
            elif tc:
                if hints and i == 0:
                    ops.lib.mnr_tc_cache_hints(2)
                pending.append(ops.tc_linear_bwd_weight_partials(wg, M, act_i, self.tc_ws[i]))
                if hints and i == 0:
                    ops.lib.mnr_tc_cache_hints(0)
                if len(pending) == 4:
                    ops.tc_reduce_partials(pending)
                    pending = []
            else:
                ops.linear_bwd_weight(wg, M, act_i, self.ws)
            if i == 0 and input_grad is not None:
                if isinstance(input_grad, (list, tuple)):
                    # one tensor-core launch for all towers, each into its own contiguous (rows, K) tensor (the caller adds them)
                    assert self.use_tc and len(input_grad) == G
                    ops.tc_linear_bwd_input([dict(a=self.dY[0][g], b=self.w_hi[0][g], b_lo=self.w_lo[0][g], c=input_grad[g], saved=saved[g],
                                                  N=n, K=k, ldb=self.w_ld[0]) for g in range(G)], M, act_i)
                else:
                    # SIMT kernel, one launch per tower: the second accumulates onto the first's result
                    for g in range(G):
                        ops.linear_bwd_input([dict(a=self.dY[0][g], b=self.lin[g][0].weight.data, c=input_grad, saved=saved[g], N=n, K=k)],
                                             M, act_i, accumulate=g > 0)
            if i > 0:
                if self.use_tc:
                    if hints:
                        ops.lib.mnr_tc_cache_hints(int(os.environ.get('MNR_TC_HINTS_DX', '3,0').split(',')[min(top - i, 1)]))
                    ops.tc_linear_bwd_input([dict(a=self.dY[i][g], b=self.w_hi[i][g], b_lo=self.w_lo[i][g], c=self.dY[i - 1][g],
                                                  saved=saved[g], N=n, K=k, ldb=self.w_ld[i]) for g in range(G)], M, act_i)
                    if hints:
                        ops.lib.mnr_tc_cache_hints(0)
                else:
                    ops.linear_bwd_input([dict(a=self.dY[i][g], b=self.lin[g][i].weight.data, c=self.dY[i - 1][g], saved=saved[g], N=n, K=k)
                                          for g in range(G)], M, act_i)
        if folds is not None:
            ops.fold_defer(None)
            assert pending and weight_grads, "deferred folds need the split-reduce launch"
        if pending:
            ops.tc_reduce_partials(pending, folds)


def _same_shape(a, v):
    if a is None or v is None or a[1] != v[1] or len(a[0]) != len(v[0]):
        return False
    for (la, na), (lv, nv_) in zip(a[0], v[0]):
        if la.weight.shape != lv.weight.shape or (na is None) != (nv_ is None):
            return False
        if na is not None and (na.eps != nv_.eps or la.weight.shape[0] > 512):
            return False
    return True


class FusedTowers:
    """PPO actor + value towers (TowerStack with G = 2) plus the policy / value heads."""

    @staticmethod
    def supported(model):
        if not getattr(model, 'separate_value_mlp', False):
            return False
        return _same_shape(_tower_layers(model.actor_mlp), _tower_layers(model.value_mlp))

    def __init__(self, model, rows, device, use_tc=True):
        self.model, self.M = model, rows
        (ba, self.act), (bv, _) = _tower_layers(model.actor_mlp), _tower_layers(model.value_mlp)
        self.stack = st = TowerStack([ba, bv], self.act, rows, device, use_tc)
        # aliases (bench.py and the PPO engine read these)
        self.la, self.lv = st.lin
        self.normed, self.gemm_act, self.L, self.widths, self.in_dim = st.normed, st.gemm_act, st.L, st.widths, st.in_dim
        self.Y, self.dY, self.use_tc, self.in_ld = st.Y, st.dY, st.use_tc, st.in_ld
        if st.normed:
            self.Z = st.Z
        if st.use_tc:
            self.w_hi, self.w_lo, self.w_ld = st.w_hi, st.w_lo, st.w_ld
        self.A = model.mu.weight.shape[0]
        self.fixed_sigma = model.fixed_sigma
        f32 = dict(dtype=torch.float32, device=device)
        self.mu_raw = torch.empty(rows, self.A, **f32)
        self.value = torch.empty(rows, 1, **f32)
        self.ls = None if self.fixed_sigma else torch.empty(rows, self.A, **f32)
        n, k = model.mu.weight.shape
        self.ws = torch.zeros(max(ops.lib.mnr_linear_bwd_weight_workspace_bytes(2, rows, n, k), 256), dtype=torch.uint8, device=device)
        self.colsum_ws = torch.zeros(ops.lib.mnr_colsum_workspace_bytes(self.A), dtype=torch.uint8, device=device)

    # ------------------------------------------------------------------------------------------------
    def forward(self, x, x_ld=None):
        """x (rows, in_dim) [row stride x_ld floats] -> (mu_raw (rows,A), logstd ((A,) parameter if fixed sigma
        else (rows,A)), value (rows,1))."""
        self.forward_towers(x, x_ld)
        return self.heads_forward()

    def forward_towers(self, x, x_ld=None, split=True):
        """Both towers up to their last activation (self.Y[-1]).  split=False: the (hi, lo) weight copies are current."""
        self.stack.forward(x, x_ld, split=split)

    def backward_towers(self, x, folds=None):
        self.stack.backward(x, folds=folds)

    def heads_forward(self):
        M, m = self.M, self.model
        inp = self.Y[-1]
        h = self.widths[-1]
        heads = [dict(a=inp[0], b=m.mu.weight.data, c=self.mu_raw, bias=m.mu.bias.data, N=self.A, K=h),
                 dict(a=inp[1], b=m.value.weight.data, c=self.value, bias=m.value.bias.data, N=1, K=h)]
        if not self.fixed_sigma:
            heads.append(dict(a=inp[0], b=m.sigma.weight.data, c=self.ls, bias=m.sigma.bias.data, N=self.A, K=h))
        ops.linear_fwd(heads, M, 0)
        return self.mu_raw, (m.sigma.data if self.fixed_sigma else self.ls), self.value

    def backward(self, x, g_mu, g_ls, g_v):
        """Seeds: d loss / d (mu_raw, logstd, value), each (rows, .).  Writes every parameter's .grad."""
        self.heads_backward(g_mu, g_ls, g_v)
        self.backward_towers(x)

    @property
    def fused_heads_ok(self):
        h = self.widths[-1]
        return h % 8 == 0 and h <= 128 and self.A <= 32

    def heads_dz_act(self):
        """Activation code the fused heads kernel should fold into its dY3 output (cfg.dz_act), or 0.  Enables
        `stack.top_is_dz` as a side effect so that the towers' backward matches."""
        ok = (self.fused_heads_ok and not self.normed and self.gemm_act in (1, 2)
              and bool(ops.lib.mnr_ppo_heads_is_fused(self.A, self.widths[-1])))
        self.stack.top_is_dz = ok
        return self.gemm_act if ok else 0

    def heads_loss_fwd_bwd(self, d, s, e, g_mu, g_ls, g_v, metrics, cfg, ws, mu_mean, sigma_mean):
        """Heads forward + PPO loss + its backward + head parameter gradients + d loss / d Y[-1] (csrc/heads_fused.cu,
        one persistent kernel; csrc/ppo_loss.cu two-kernel form for odd shapes).  `d` is the env-major data dict,
        [s:e) the minibatch rows."""
        m = self.model
        ya, yv = self.Y[-1]
        fixed = self.fixed_sigma
        ops.ppo_heads_loss_fwd_bwd(
            ya, yv, m.mu.weight.data, m.mu.bias.data, m.value.weight.data, m.value.bias.data,
            None if fixed else m.sigma.weight.data, m.sigma.data if fixed else m.sigma.bias.data,
            d['actions'][s:e], d['neglogp'][s:e], d['advantages'][s:e], d['values'][s:e], d['returns'][s:e], d['mu'][s:e], d['sigma'][s:e],
            self.dY[-1][0], self.dY[-1][1], g_mu, g_ls, g_v, m.mu.weight.grad, m.mu.bias.grad, m.value.weight.grad, m.value.bias.grad,
            None if fixed else m.sigma.weight.grad, m.sigma.grad if fixed else m.sigma.bias.grad, metrics, cfg, ws, True, mu_mean, sigma_mean)

    def heads_backward(self, g_mu, g_ls, g_v):
        M, m, top = self.M, self.model, self.L - 1
        h = self.widths[-1]
        ya, yv = self.Y[top]
        # heads: weight/bias grads, then grads w.r.t. the towers' last outputs
        wg = [dict(a=g_mu, b=ya, c=m.mu.weight.grad, bias_grad=m.mu.bias.grad, N=self.A, K=h)]
        if not self.fixed_sigma:
            wg.append(dict(a=g_ls, b=ya, c=m.sigma.weight.grad, bias_grad=m.sigma.bias.grad, N=self.A, K=h))
        ops.linear_bwd_weight(wg, M, 0, self.ws)
        ops.linear_bwd_weight([dict(a=g_v, b=yv, c=m.value.weight.grad, bias_grad=m.value.bias.grad, N=1, K=h)], M, 0, self.ws)
        if self.fixed_sigma:
            ops.colsum(g_ls, m.sigma.grad, self.colsum_ws)
        ops.linear_bwd_input([dict(a=g_mu, b=m.mu.weight.data, c=self.dY[top][0], N=self.A, K=h),
                              dict(a=g_v, b=m.value.weight.data, c=self.dY[top][1], N=1, K=h)], M, 0)
        if not self.fixed_sigma:
            ops.linear_bwd_input([dict(a=g_ls, b=m.sigma.weight.data, c=self.dY[top][0], N=self.A, K=h)], M, 0, accumulate=True)


class FusedCritics:
    """1 or 2 state-value MLPs with a plain last Linear (SHAC `Critic`, SAPO `EnsembleCritic`,
    mineral/agents/diffrl/models.py:149-256) trained on the MSE of SHAC.compute_critic_loss (shac.py:869-875):
    hidden layers through a TowerStack, the 1-wide output layer and its backward on the SIMT linear kernels, loss and
    its gradient in one small kernel."""

    @staticmethod
    def parse(mlps):
        parsed = [_tower_layers(m, plain_last=True) for m in mlps]
        if not 1 <= len(parsed) <= 2 or any(p is None for p in parsed):
            return None
        if any(not _same_shape(parsed[0], p) for p in parsed[1:]):
            return None
        if parsed[0][0][-1][0].weight.shape[0] != 1:
            return None
        return parsed

    def __init__(self, mlps, rows, device, use_tc=True):
        parsed = self.parse(mlps)
        assert parsed is not None, 'unsupported critic architecture'
        self.G, self.M = len(mlps), rows
        self.stack = TowerStack([p[0][:-1] for p in parsed], parsed[0][1], rows, device, use_tc)
        self.last = [p[0][-1][0] for p in parsed]
        f32 = dict(dtype=torch.float32, device=device)
        self.pred = [torch.empty(rows, 1, **f32) for _ in range(self.G)]
        self.g_pred = [torch.empty(rows, 1, **f32) for _ in range(self.G)]
        h = self.stack.widths[-1]
        self.ws = torch.zeros(max(ops.lib.mnr_linear_bwd_weight_workspace_bytes(self.G, rows, 1, h), 256), dtype=torch.uint8, device=device)
        self.mse_ws = torch.zeros(ops.lib.mnr_critic_mse_workspace_bytes(rows), dtype=torch.uint8, device=device)

    def forward(self, x, x_ld=None, rows=None):
        """x (rows, in_dim) [row stride x_ld floats] -> self.pred[g] (rows, 1) for every critic g."""
        rows = self.M if rows is None else rows
        st = self.stack
        h = st.widths[-1]
        st.forward(x, x_ld, rows)
        ops.linear_fwd([dict(a=st.Y[-1][g], b=self.last[g].weight.data, c=self.pred[g], bias=self.last[g].bias.data, N=1, K=h)
                        for g in range(self.G)], rows, 0)
        return self.pred

    def loss_fwd_bwd(self, x, target, loss_out, x_ld=None):
        """x (rows, in_dim), target (rows) -> loss_out (1 float), every critic parameter's .grad."""
        rows, G, st = target.numel(), self.G, self.stack
        h = st.widths[-1]
        self.forward(x, x_ld, rows)
        ops.critic_mse_fwd_bwd([p.view(-1)[:rows] for p in self.pred], target, [g_.view(-1)[:rows] for g_ in self.g_pred], loss_out, self.mse_ws)
        ops.linear_bwd_weight([dict(a=self.g_pred[g], b=st.Y[-1][g], c=self.last[g].weight.grad, bias_grad=self.last[g].bias.grad, N=1, K=h)
                               for g in range(G)], rows, 0, self.ws)
        ops.linear_bwd_input([dict(a=self.g_pred[g], b=self.last[g].weight.data, c=st.dY[-1][g], N=1, K=h) for g in range(G)], rows, 0)
        st.backward(x, rows)
