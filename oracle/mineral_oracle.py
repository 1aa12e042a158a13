"""CPU oracle for the rollout->update data path of jacr13/mineral.

TEST INFRASTRUCTURE ONLY.  This module is the *checker*: only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import it.  Nothing under `mineral_b200/` imports it and the product
path has no CPU fallback.

It restates, as plain functions over CPU torch tensors, the algorithms the
reference implements with eager torch ops.  The arithmetic itself lives in a
third-party dependency of the reference: `torch` (pinned 2.6.0+cu124 in
/root/reference/poetry.lock:2420-2421; this image has 2.11.0+cu128) and numpy
for `np.random.permutation`.  Every function cites the reference file:line it
follows (paths relative to /root/reference).

Pinning: the reference ships no tests, fixtures or golden vectors
(SURVEY.md section 4), so the oracle is pinned against outputs of the reference
itself, run in the build container through `oracle/reference_shim.py` by
`oracle/make_golden.py`; the frozen vectors live in `tests/golden/*.pt` and
`tests/test_oracle_golden.py` checks every function below against them.

Scalar conventions copied from the reference: Python scalars such as
`gamma * tau`, `1.0 - lam`, `lam * gamma` are formed in double precision first
and rounded to f32 only when they meet an f32 tensor.
"""

import math

import numpy as np
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------
# K1a  GAE reverse scan            mineral/agents/ppo/experience.py:96-108
# --------------------------------------------------------------------------


def gae_scan(rewards, values, dones, last_values, gamma, tau):
    """rewards/values (T,N,1) f32, dones (T,N) u8, last_values (N,1).
    Returns (returns, advantages), both (T,N,1) f32, time-major.
    Association order is the reference's: ((g*nv)*alive), ((g*tau)*alive)*carry."""
    T = rewards.shape[0]
    adv = torch.zeros_like(rewards)
    ret = torch.zeros_like(rewards)
    carry = 0
    for t in range(T - 1, -1, -1):
        nxt = last_values if t == T - 1 else values[t + 1]
        alive = (1.0 - dones[t].float()).unsqueeze(1)
        delta = rewards[t] + gamma * nxt * alive - values[t]
        carry = delta + gamma * tau * alive * carry
        adv[t] = carry
        ret[t] = carry + values[t]
    return ret, adv


# --------------------------------------------------------------------------
# K3a  env-major pack + advantage normalisation
#      mineral/agents/ppo/experience.py:7-14 (transform_op), :110-121
# --------------------------------------------------------------------------


def env_major(x):
    """(T,N,...) -> (N*T,...), row = n*T + t   (experience.py:7-14)."""
    T, N = x.shape[0], x.shape[1]
    return x.transpose(0, 1).reshape(N * T, *x.shape[2:])


def pack_rollout(storage):
    """storage: dict of (T,N,...) tensors, 'obses' a dict of them, 'returns'
    already filled by gae_scan.  Returns the env-major training dict with
    normalised advantages (experience.py:110-121; the std is the unbiased one
    and the 1e-8 is added to the std, not the variance)."""
    out = {}
    for k, v in storage.items():
        out[k] = {kk: env_major(vv) for kk, vv in v.items()} if isinstance(v, dict) else env_major(v)
    a = out["returns"] - out["values"]
    out["advantages"] = ((a - a.mean()) / (a.std() + 1e-8)).squeeze(1)
    return out


def minibatch_rows(i, minibatch_size):
    """contiguous slice, no shuffle (experience.py:61-71)."""
    return i * minibatch_size, (i + 1) * minibatch_size


# --------------------------------------------------------------------------
# K2  RunningMeanStd               mineral/common/normalizers.py:19-81
# --------------------------------------------------------------------------


def rms_init(shape, eps=1e-4, initial_count="eps", dtype=torch.float32):
    """state = [mean, var, count]   (normalizers.py:36-42)."""
    c = eps if initial_count == "eps" else initial_count
    return [torch.zeros(shape, dtype=dtype), torch.ones(shape, dtype=dtype), torch.tensor(c, dtype=dtype)]


def rms_merge(mean, var, count, b_mean, b_var, b_count):
    """Chan parallel merge (normalizers.py:44-55), evaluated left to right as written there."""
    delta = b_mean - mean
    tot = count + b_count
    new_mean = mean + delta * b_count / tot
    m2 = var * count + b_var * b_count + delta**2 * count * b_count / tot
    return new_mean, m2 / tot, tot


def rms_update(state, x, correction=1):
    """batch moments in x.dtype (f32), merge in the running dtype (normalizers.py:57-63)."""
    bm = x.mean(dim=0)
    bv = x.var(dim=0, correction=correction)
    state[0], state[1], state[2] = rms_merge(state[0], state[1], state[2], bm, bv, x.shape[0])
    return state


def _clip(x, clamp_range):
    return x if clamp_range is None else torch.clamp(x, min=clamp_range[0], max=clamp_range[1])


def rms_normalize(state, x, eps, clamp_range=None):
    """(x - mean.f32) / sqrt(var.f32 + eps) then clamp (normalizers.py:65-68)."""
    return _clip((x - state[0].float()) / torch.sqrt(state[1].float() + eps), clamp_range)


def rms_unnormalize(state, x, eps, clamp_range=None):
    """clamp the INPUT first, then scale back (normalizers.py:70-73)."""
    return torch.sqrt(state[1].float() + eps) * _clip(x, clamp_range) + state[0].float()


def ppo_value_block(value_rms, data, eps=1e-5, clamp_range=(-5.0, 5.0)):
    """update(V); V=norm(V); update(ret); ret=norm(ret)   (mineral/agents/ppo/ppo.py:175-183)."""
    v, r = data["values"], data["returns"]
    rms_update(value_rms, v)
    v = rms_normalize(value_rms, v, eps, clamp_range)
    rms_update(value_rms, r)
    r = rms_normalize(value_rms, r, eps, clamp_range)
    data["values"], data["returns"] = v, r
    return data


# --------------------------------------------------------------------------
# K1b  SHAC / SAPO critic targets   mineral/agents/diffrl/shac.py:877-916
# --------------------------------------------------------------------------


def shac_entropy_term(distr_ent, logprobs, use_distr_ent, target_entropy, offset, scale):
    """shac.py:894-898."""
    ent = distr_ent if use_distr_ent else -1.0 * logprobs
    if offset:
        ent = (ent + abs(target_entropy)) * 0.5
    if scale:
        ent = ent * (1.0 / abs(target_entropy))
    return ent


def td_lambda_targets(rew, done_mask, next_values, gamma, lam, alpha=None, entropy=None):
    """rew/done_mask/next_values (T,N) f32 (done_mask is FLOAT here).  shac.py:880-889;
    with `alpha` and `entropy` it is the `_with_entropy` variant, shac.py:904-914."""
    T, N = rew.shape
    A = torch.zeros(N)
    B = torch.zeros(N)
    w = torch.ones(N)
    out = torch.zeros_like(rew)
    for i in range(T - 1, -1, -1):
        d = done_mask[i]
        w = w * lam * (1.0 - d) + d
        r = rew[i] if alpha is None else rew[i] + alpha * entropy[i]
        adj = (1.0 - w) / (1.0 - lam) * r
        A = (1.0 - d) * (lam * gamma * A + gamma * next_values[i] + adj)
        B = gamma * (next_values[i] * d + B * (1.0 - d)) + r
        out[i] = (1.0 - lam) * A + w * B
    return out


def one_step_targets(rew, next_values, gamma, alpha=None, entropy=None):
    """shac.py:879 and :902."""
    if alpha is None:
        return rew + gamma * next_values
    return (rew + alpha * entropy) + gamma * next_values


def critic_dataset_slices(n_rows, batch_size, drop_last=False):
    """row ranges of CriticDataset over the TIME-major flatten (row = t*B + b)
    (mineral/agents/diffrl/utils.py:39-69)."""
    n = n_rows // batch_size if drop_last else ((n_rows - 1) // batch_size) + 1
    return [(i * batch_size, min((i + 1) * batch_size, n_rows)) for i in range(n)]


# --------------------------------------------------------------------------
# a16  SHAC / SAPO critic regression   mineral/agents/diffrl/shac.py:824-875,
#      Critic / EnsembleCritic mineral/agents/diffrl/models.py:149-256,
#      grad_norm mineral/agents/diffrl/utils.py:31-36
# --------------------------------------------------------------------------


def shac_critic_inputs(seed, T, B, O, wild=False):
    """Seeded (obs (T,B,O), target_values (T,B)) of the critic-regression goldens (oracle/make_golden.py:gen_shac_critic);
    regenerated from the seed in the tests so the fixture stays small."""
    g = torch.Generator().manual_seed(100 + seed)
    obs = torch.randn(T, B, O, generator=g)
    tv = 3 * torch.randn(T, B, generator=g)
    if wild:
        tv[3, 5] = 1e4  # one wild target: exercises the clip
    return obs, tv


def critic_forward(params, prefix, x, n_hidden, norm, act):
    """MLP(state_dim, out_dim=1, plain_last=True): n_hidden x (Linear -> [LayerNorm] -> act), then a plain Linear."""
    h = mlp_tower(params, prefix, x, n_hidden, norm, act)
    j = n_hidden * (1 + (1 if norm else 0) + (1 if act else 0))
    return F.linear(h, params[f"{prefix}.mlp.{j}.weight"], params[f"{prefix}.mlp.{j}.bias"])


def shac_critic_loss(params, prefixes, obs, target, n_hidden, norm, act):
    """compute_critic_loss (shac.py:869-875): mean over critics of F.mse_loss(pred.squeeze(-1), target)."""
    losses = [F.mse_loss(critic_forward(params, pf, obs, n_hidden, norm, act).squeeze(-1), target) for pf in prefixes]
    return torch.stack(losses).mean()


def shac_update_critic(params, prefixes, optim, obs_flat, target_flat, batch_size, iterations, n_hidden, norm, act,
                       truncate_grads=True, max_grad_norm=1.0, drop_last=False):
    """update_critic (shac.py:824-867) over the time-major flattened dataset (CriticDataset, no shuffle, keep the
    short last batch).  params: dict name -> leaf tensor (requires_grad) that `optim` steps.  Returns the results dict."""
    results = {"value_loss": [], "grad_norm_before_clip/critic": [], "grad_norm_after_clip/critic": []}
    leaves = list(params.values())
    for _ in range(iterations):
        total, before, after = 0.0, [], []
        slices = critic_dataset_slices(target_flat.shape[0], batch_size, drop_last=drop_last)   # B = len(dataset), shac.py:832-835
        for s, e in slices:
            optim.zero_grad()
            loss = shac_critic_loss(params, prefixes, obs_flat[s:e], target_flat[s:e], n_hidden, norm, act)
            loss.backward()
            for p in leaves:
                p.grad.nan_to_num_(0.0, 0.0, 0.0)
            if truncate_grads:
                before.append(torch.sqrt(sum(torch.sum(p.grad ** 2) for p in leaves)))
                if max_grad_norm is not None:
                    torch.nn.utils.clip_grad_norm_(leaves, max_grad_norm)
                after.append(torch.sqrt(sum(torch.sum(p.grad ** 2) for p in leaves)))
            optim.step()
            total = total + loss
        results["value_loss"].append((total / len(slices)).detach())
        if truncate_grads:
            results["grad_norm_before_clip/critic"].append(torch.mean(torch.stack(before)))
            results["grad_norm_after_clip/critic"].append(torch.mean(torch.stack(after)))
    return results


# --------------------------------------------------------------------------
# 8(f)-2  SAC / DDPG update math     mineral/agents/ddpg/utils.py:8-31,
#         mineral/agents/sac/sac.py:293-297, mineral/agents/ddpg/ddpg.py:292-293
# --------------------------------------------------------------------------


def soft_update(target, current, tau):
    """tar.mul_(1.0 - tau).add_(cur * tau) on one tensor (ddpg/utils.py:8-11); returns the new target."""
    return target.clone().mul_(1.0 - tau).add_(current * tau)


def td_target(q_values, reward, done, gamma, nstep=1, alpha=None, log_prob=None):
    """sac.py:293-297 / ddpg.py:292-293."""
    target_q = torch.min(torch.stack(list(q_values)), dim=0).values
    if alpha is not None:
        target_q = target_q - alpha * log_prob
    return reward + (1 - done) * (gamma ** nstep) * target_q


def distl_projection(next_dist, reward, done, gamma, v_min=-10, v_max=10, num_atoms=51, support=None):
    """C51 projection (ddpg/utils.py:14-31) as explicit per-row loops in the reference's accumulation order."""
    delta_z = (v_max - v_min) / (num_atoms - 1)
    target_z = (reward + (1 - done) * gamma * support).clamp(min=v_min, max=v_max)
    b = (target_z - v_min) / delta_z
    l, u = b.floor().long(), b.ceil().long()
    l[torch.logical_and(u > 0, l == u)] -= 1
    u[torch.logical_and(l < (num_atoms - 1), l == u)] += 1
    out = torch.zeros_like(next_dist)
    lo_w, up_w = next_dist * (u.float() - b), next_dist * (b - l.float())
    for i in range(next_dist.shape[0]):
        for j in range(num_atoms):
            out[i, l[i, j]] += lo_w[i, j]
        for j in range(num_atoms):
            out[i, u[i, j]] += up_w[i, j]
    return out


# --------------------------------------------------------------------------
# K3c/K3d  circular replay          mineral/buffers/simple_replay.py:19-85
# --------------------------------------------------------------------------


# ---- SAC policy head (8(f)-2 agents): Dist + SquashedNormal (mineral/nets/dist.py:37-52, mineral/nets/distributions.py:10-58) ----
def squashed_sigma(logstd, dist_type="squashed_normal", minlogstd=None, maxlogstd=None, minstd=1.0, maxstd=1.0):
    """-> (sigma, d sigma / d logstd).  squashed_normal: exp(clamp(logstd)) (clamp's gradient is 1 on the closed range);
    dreamerv3_squashed_normal: (hi - lo) * sigmoid(logstd + 2) + lo."""
    if dist_type == "squashed_normal":
        ls, inside = logstd, torch.ones_like(logstd, dtype=torch.bool)
        if minlogstd is not None or maxlogstd is not None:
            ls = torch.clamp(logstd, minlogstd, maxlogstd)
            if minlogstd is not None:
                inside &= logstd >= minlogstd
            if maxlogstd is not None:
                inside &= logstd <= maxlogstd
        sigma = ls.exp()
        return sigma, sigma * inside
    s = torch.sigmoid(logstd + 2.0)
    return (maxstd - minstd) * s + minstd, (maxstd - minstd) * s * (1.0 - s)


def squashed_sample(mu, logstd, eps, **dist_kwargs):
    """SAC.get_actions(sample=True, logprob=True) after the heads (mineral/agents/sac/sac.py:126-134): u = mu + sigma * eps
    (Normal.rsample), actions = tanh(u), log_prob = sum_j [Normal.log_prob(u) - 2 (log 2 - u - softplus(-2u))] (B,1)."""
    sigma, _ = squashed_sigma(logstd, **dist_kwargs)
    u = mu + eps * sigma
    base = -((u - mu) ** 2) / (2 * sigma ** 2) - sigma.log() - math.log(math.sqrt(2 * math.pi))
    ladj = 2.0 * (math.log(2.0) - u - F.softplus(-2.0 * u))
    return u.tanh(), (base - ladj).sum(-1, keepdim=True)


def squashed_bwd(mu, logstd, eps, g_act, c_logp, **dist_kwargs):
    """Gradient of  c_logp * sum_rows log_prob + <g_act, actions>  w.r.t. (mu, logstd) through the reparameterised sample,
    in closed form: the Normal.log_prob contributions through u and through (mu, sigma) cancel to -1/sigma.
    (what autograd does for sac.py:318-324 with c_logp = alpha / B and g_act = d(-mean Q_min)/d actions.)"""
    sigma, dsig = squashed_sigma(logstd, **dist_kwargs)
    a = (mu + eps * sigma).tanh()
    g_u = c_logp * 2.0 * a + g_act * (1.0 - a * a)
    return g_u, (g_u * eps - c_logp / sigma) * dsig


def tanh_action(mu, noise=None, noise_bounds=None, out_bounds=None):
    """DDPG actor output + exploration / target-smoothing noise (mineral/agents/ddpg/models.py:55-58, noise.py:22-38)."""
    a = mu.tanh()
    if noise is not None:
        if noise_bounds is not None:
            noise = noise.clamp(noise_bounds[0], noise_bounds[1])
        a = a + noise
        if out_bounds is not None:
            a = a.clamp(out_bounds[0], out_bounds[1])
    return a


def actor_objective(q_values, log_prob, alpha, target_entropy):
    """SAC.update_actor's scalars (sac.py:320-331) and the seeds d actor_loss / d Q_k (-1/B at the first minimum):
    -> (actor_loss, entropy, alpha_loss, seeds list)."""
    qs = torch.stack(q_values)
    q_min, arg = torch.min(qs, dim=0)
    B = q_min.shape[0]
    lp = log_prob if log_prob is not None else torch.zeros_like(q_min)
    actor_loss = (alpha * lp - q_min).mean()
    entropy = (-lp).mean()
    alpha_loss = (alpha * (-lp - target_entropy)).mean()
    seeds = [torch.where(arg == k, -1.0 / B, 0.0).to(q_min.dtype) for k in range(len(q_values))]
    return actor_loss, entropy, alpha_loss, seeds


class RingOracle:
    """Row store with the reference's wrap rule: on overflow the head
    [next_p:C] receives the FIRST C-next_p incoming rows and the tail [0:p]
    receives the LAST p incoming rows (simple_replay.py:42-60)."""

    def __init__(self, obs_dim, act_dim, capacity):
        self.C = int(capacity)
        self.obs = torch.zeros(self.C, obs_dim)
        self.act = torch.zeros(self.C, act_dim)
        self.rew = torch.zeros(self.C, 1)
        self.nobs = torch.zeros(self.C, obs_dim)
        self.done = torch.zeros(self.C, 1, dtype=torch.bool)
        self.next_p, self.full, self.cur = 0, False, 0

    def insert(self, obs, act, rew, nobs, done):
        fields = [
            (self.obs, obs.reshape(-1, self.obs.shape[1])),
            (self.act, act.reshape(-1, self.act.shape[1])),
            (self.rew, rew.reshape(-1, 1)),
            (self.nobs, nobs.reshape(-1, self.nobs.shape[1])),
            (self.done, done.reshape(-1, 1).bool()),
        ]
        n = fields[2][1].shape[0]
        p = self.next_p + n
        if p > self.C:
            self.full = True
            head = self.C - self.next_p
            p -= self.C
            for buf, src in fields:
                buf[self.next_p : self.C] = src[:head]
                buf[0:p] = src[-p:]
        else:
            for buf, src in fields:
                buf[self.next_p : p] = src
        self.next_p = p
        self.cur = self.C if self.full else self.next_p

    def gather(self, idx):
        """rows in `idx` order; done returned as f32 (simple_replay.py:77-85)."""
        return self.obs[idx], self.act[idx], self.rew[idx], self.nobs[idx], self.done[idx].float()


# --------------------------------------------------------------------------
# K3e  n-step FIFO                  mineral/buffers/nstep_replay.py:29-94
# --------------------------------------------------------------------------


class NStepOracle:
    """Sliding window of the last `nstep` transitions per env, persistent across
    calls.  Emits, once the window is full, (obs[oldest], act[oldest], R, next_obs*,
    done*) where the discounted sum stops after the first done in the window,
    next_obs* is the one at the first done (else the newest) and done* is
    `newest done OR any done` (nstep_replay.py:71-94).  Outputs are concatenated
    step-major (nstep_replay.py:60-64).  `done` is FLOAT in and out: the reference's window is
    promoted to float by the first torch.cat with SAC's float dones (sac.py:150) and its argmax
    rejects bool input."""

    def __init__(self, num_envs, obs_dim, act_dim, nstep=3, gamma=0.99):
        self.n, self.N = nstep, num_envs
        self.w_obs = torch.zeros(num_envs, nstep, obs_dim)
        self.w_act = torch.zeros(num_envs, nstep, act_dim)
        self.w_rew = torch.zeros(num_envs, nstep, 1)
        self.w_nobs = torch.zeros(num_envs, nstep, obs_dim)
        self.w_done = torch.zeros(num_envs, nstep, 1)
        self.count = 0
        # f32 tensor built from python doubles gamma**i (nstep_replay.py:27)
        self.disc = torch.tensor([gamma**i for i in range(nstep)]).view(-1, 1)

    @staticmethod
    def _push(win, new):
        return torch.cat((win[:, 1:], new.unsqueeze(1)), dim=1)

    def push(self, obs, act, rew, nobs, done):
        """obs/nobs (N,S,O), act (N,S,A), rew/done (N,S,1)."""
        if self.n == 1:
            return obs, act, rew, nobs, done
        outs = ([], [], [], [], [])
        for s in range(done.shape[1]):
            self.w_obs = self._push(self.w_obs, obs[:, s])
            self.w_nobs = self._push(self.w_nobs, nobs[:, s])
            self.w_done = self._push(self.w_done, done[:, s].float())
            self.w_act = self._push(self.w_act, act[:, s])
            self.w_rew = self._push(self.w_rew, rew[:, s])
            self.count += 1
            if self.count < self.n:
                continue
            d = self.w_done.squeeze(-1) != 0  # (N,n)
            any_done = d.any(dim=1)
            first = d.float().argmax(dim=1)  # first True, 0 if none
            k = torch.arange(self.n)
            live = torch.where(any_done[:, None], k[None, :] <= first[:, None], torch.ones_like(d))
            R = ((self.w_rew * self.disc) * live.unsqueeze(-1)).sum(1)
            pick = torch.where(any_done, first, torch.full_like(first, self.n - 1))
            nob = self.w_nobs[torch.arange(self.N), pick]
            dn = torch.where(any_done[:, None], torch.ones(self.N, 1), self.w_done[:, -1])
            for lst, v in zip(outs, (self.w_obs[:, 0], self.w_act[:, 0], R, nob, dn)):
                lst.append(v.clone())
        return tuple(torch.cat(x) for x in outs)


# --------------------------------------------------------------------------
# K4  PPO losses                    mineral/agents/ppo/ppo.py:449-498
# --------------------------------------------------------------------------


def smooth_clamp(x, lo, hi):
    """ppo.py:449-450."""
    return 1 / (1 + torch.exp((-(x - lo) / (hi - lo) + 0.5) * 4)) * (hi - lo) + lo


def actor_loss(old_nlp, nlp, adv, e_clip, use_smooth_clamp=False):
    """ppo.py:453-463 -> (per-row loss, clip fraction)."""
    ratio = torch.exp(old_nlp - nlp)
    lo, hi = 1.0 - e_clip, 1.0 + e_clip
    clipped_ratio = smooth_clamp(ratio, lo, hi) if use_smooth_clamp else torch.clamp(ratio, lo, hi)
    loss = torch.max(-(adv * ratio), -(adv * clipped_ratio))
    frac = ((ratio - 1.0).abs() > e_clip).float().mean().detach()
    return loss, frac


def critic_loss(old_values, values, e_clip, returns, clip_value_loss):
    """ppo.py:466-476 -> (per-row loss (mb,1), explained variance)."""
    plain = (values - returns) ** 2
    if clip_value_loss:
        vclip = old_values + (values - old_values).clamp(-e_clip, e_clip)
        loss = torch.max(plain, (vclip - returns) ** 2)
    else:
        loss = plain
    ev = (1 - torch.var(returns - values) / (torch.var(returns) + 1e-8)).clamp(0, 1).detach()
    return loss, ev


def bounds_loss(mu, bounds_type, soft_bound=1.1):
    """ppo.py:479-489."""
    if bounds_type == "bound":
        hi = torch.clamp_min(mu - soft_bound, 0.0) ** 2
        lo = torch.clamp_max(mu + soft_bound, 0.0) ** 2
        return (lo + hi).sum(dim=-1)
    if bounds_type == "reg":
        return (mu * mu).sum(dim=-1)
    raise NotImplementedError(bounds_type)


def policy_kl(p0_mu, p0_sigma, p1_mu, p1_sigma, reduce=True):
    """ppo.py:492-498 (mean over rows); diffrl/utils.py:14-22 is the un-reduced twin."""
    kl = torch.log(p1_sigma / p0_sigma + 1e-5) + (p0_sigma**2 + (p1_mu - p0_mu) ** 2) / (2.0 * (p1_sigma**2 + 1e-5)) - 0.5
    kl = kl.sum(dim=-1)
    return kl.mean() if reduce else kl


# --------------------------------------------------------------------------
# a10  ActorCritic forward          mineral/agents/ppo/models.py:93-151,
#      MLP mineral/nets/mlp.py:25-77, Dist mineral/nets/dist.py:31-58
# --------------------------------------------------------------------------

_HALF_LOG_2PI = 0.5 * math.log(2 * math.pi)


def mlp_tower(params, prefix, x, n_layers, norm, act):
    """Linear -> [LayerNorm] -> act, n_layers times.  Parameter names follow the
    reference's nn.Sequential numbering so a reference state_dict plugs in."""
    stride = 1 + (1 if norm else 0) + (1 if act else 0)
    for i in range(n_layers):
        j = i * stride
        x = F.linear(x, params[f"{prefix}.mlp.{j}.weight"], params[f"{prefix}.mlp.{j}.bias"])
        if norm:
            w = params[f"{prefix}.mlp.{j + 1}.weight"]
            x = F.layer_norm(x, (w.shape[0],), w, params[f"{prefix}.mlp.{j + 1}.bias"], 1e-5)
        # the reference builds its activations with inplace=True (mlp.py:37): the in-place ELU backward is evaluated from
        # the OUTPUT, one ulp away from the out-of-place formula, so the oracle does the same
        if act == "ELU":
            x = F.elu(x, inplace=True)
        elif act == "SiLU":
            x = F.silu(x, inplace=True)
        elif act == "ReLU":
            x = F.relu(x, inplace=True)
        elif act is not None:
            raise NotImplementedError(act)
    return x


def dist_params(mu_raw, logstd, dist_type, minstd=1.0, maxstd=1.0):
    """dist.py:32-47 -> (mu, sigma)."""
    if dist_type == "normal":
        return mu_raw, torch.exp(logstd)
    if dist_type == "dreamerv3_normal":
        return torch.tanh(mu_raw), (maxstd - minstd) * torch.sigmoid(logstd + 2.0) + minstd
    raise NotImplementedError(dist_type)


def normal_neglogp(mu, sigma, a):
    """-sum_j log N(a_j; mu_j, sigma_j), torch.distributions.Normal.log_prob form."""
    var = sigma**2
    lp = -((a - mu) ** 2) / (2 * var) - sigma.log() - math.log(math.sqrt(2 * math.pi))
    return -lp.sum(dim=-1)


def normal_entropy(sigma):
    return (0.5 + _HALF_LOG_2PI + torch.log(sigma)).sum(dim=-1)


def actor_critic_forward(params, obs, actions, net):
    """net: dict(n_layers, norm(bool), act, fixed_sigma, dist_type, minstd, maxstd).
    Returns dict(values (mb,1), entropy (mb,), mu, sigma (mb,A), prev_neglogp (mb,))."""
    x = mlp_tower(params, "actor_mlp", obs, net["n_layers"], net["norm"], net["act"])
    mu_raw = F.linear(x, params["mu.weight"], params["mu.bias"])
    if net["fixed_sigma"]:
        logstd = mu_raw * 0 + params["sigma"]
    else:
        logstd = mu_raw * 0 + F.linear(x, params["sigma.weight"], params["sigma.bias"])
    xv = mlp_tower(params, "value_mlp", obs, net["n_layers"], net["norm"], net["act"])
    value = F.linear(xv, params["value.weight"], params["value.bias"])
    mu, sigma = dist_params(mu_raw, logstd, net["dist_type"], net.get("minstd", 1.0), net.get("maxstd", 1.0))
    out = {"values": value, "entropy": normal_entropy(sigma), "mu": mu, "sigma": sigma}
    if actions is not None:
        out["prev_neglogp"] = normal_neglogp(mu, sigma, actions)
    return out


def kl_lr_update(lr, kl, kl_threshold, min_lr=1e-6, max_lr=1e-2):
    """mineral/agents/ppo/utils.py:14-20."""
    new = lr
    if kl > 2.0 * kl_threshold:
        new = max(lr / 1.5, min_lr)
    if kl < 0.5 * kl_threshold:
        new = min(lr * 1.5, max_lr)
    return new


# --------------------------------------------------------------------------
# a8 + a9  the whole update         mineral/agents/ppo/ppo.py:169-183, 248-350
# --------------------------------------------------------------------------


def ppo_update(storage, last_values, params, optim, obs_rms, value_rms, hp, net, lr, record=None):
    """One PPO update on a filled time-major `storage` (obs key 'obs').

    params: dict name->leaf tensor (requires_grad), optim: torch optimizer over
    them, obs_rms/value_rms: [mean,var,count] f64 states, hp: hyper-parameters
    (gamma, tau, e_clip, critic_coef, entropy_coef, bounds_loss_coef, bounds_type,
    clip_value_loss, use_smooth_clamp, mini_epochs, minibatch_size, truncate_grads,
    max_grad_norm, lr_schedule, kl_threshold).  Returns (data, results, lr)."""
    ret, _ = gae_scan(storage["rewards"], storage["values"], storage["dones"], last_values, hp["gamma"], hp["tau"])
    storage["returns"] = ret
    data = pack_rollout(storage)
    data = ppo_value_block(value_rms, data)
    B = data["values"].shape[0]
    mb = hp["minibatch_size"]
    results = {k: [] for k in ("loss", "a_loss", "c_loss", "b_loss", "entropy", "clip_frac", "explained_var", "kl", "grad_norm", "avg_kl", "lr")}
    for _ep in range(hp["mini_epochs"]):
        kls = []
        for i in range(B // mb):
            s, e = minibatch_rows(i, mb)
            obs = data["obses"]["obs"][s:e]
            rms_update(obs_rms, obs)
            x = rms_normalize(obs_rms, obs, 1e-5, (-5.0, 5.0))
            out = actor_critic_forward(params, x, data["actions"][s:e], net)
            a, frac = actor_loss(data["neglogp"][s:e], out["prev_neglogp"], data["advantages"][s:e], hp["e_clip"], hp["use_smooth_clamp"])
            c, ev = critic_loss(data["values"][s:e], out["values"], hp["e_clip"], data["returns"][s:e], hp["clip_value_loss"])
            b = bounds_loss(out["mu"], hp["bounds_type"])
            a, c, ent, b = a.mean(), c.mean(), out["entropy"].mean(), b.mean()
            loss = a + 0.5 * c * hp["critic_coef"] - ent * hp["entropy_coef"] + b * hp["bounds_loss_coef"]
            optim.zero_grad()
            loss.backward()
            if record is not None and "grads" not in record:
                record["grads"] = {k: v.grad.detach().clone() for k, v in params.items()}
                record["first_out"] = {k: v.detach().clone() for k, v in out.items()}
                record["first_x"] = x.detach().clone()
            gn = torch.zeros(())
            if hp["truncate_grads"]:
                gn = torch.nn.utils.clip_grad_norm_(list(params.values()), hp["max_grad_norm"])
            optim.step()
            with torch.no_grad():
                mu, sg = out["mu"].detach(), out["sigma"].detach()
                kl = policy_kl(mu, sg, data["mu"][s:e], data["sigma"][s:e])
                data["mu"][s:e] = mu
                data["sigma"][s:e] = sg
            kls.append(kl)
            for k, v in zip(
                ("loss", "a_loss", "c_loss", "b_loss", "entropy", "clip_frac", "explained_var", "kl", "grad_norm"),
                (loss, a, c, b, ent, frac, ev, kl, gn),
            ):
                results[k].append(v.detach().clone())
        avg_kl = torch.stack(kls).mean()
        results["avg_kl"].append(avg_kl)
        if hp["lr_schedule"] == "kl":
            lr = kl_lr_update(lr, avg_kl.item(), hp["kl_threshold"], hp.get("min_lr", 1e-6), hp.get("max_lr", 1e-2))
        elif hp["lr_schedule"] == "cos":   # ppo.py:339-342, ppo/utils.py:35-37
            lr = hp["init_lr"] * 0.5 * (1.0 + math.cos(math.pi * (hp["agent_steps"] + _ep / hp["mini_epochs"]) / hp["max_agent_steps"]))
        for g in optim.param_groups:
            g["lr"] = lr
        results["lr"].append(lr)
    return data, results, lr


# --------------------------------------------------------------------------
# helpers shared by tests / bench (synthetic inputs)
# --------------------------------------------------------------------------


def synth_rollout(T, N, O, A, seed=0, done_p=0.02):
    """Seeded synthetic time-major PPO storage + last_values (CPU tensors)."""
    g = torch.Generator().manual_seed(seed)
    scale = torch.empty(O).uniform_(0.1, 10.0, generator=g)
    shift = torch.empty(O).uniform_(-3.0, 3.0, generator=g)
    mu = torch.empty(T, N, A).uniform_(-1, 1, generator=g)
    sigma = torch.empty(T, N, A).uniform_(0.1, 1.0, generator=g)
    actions = mu + sigma * torch.randn(T, N, A, generator=g)
    storage = {
        "rewards": torch.randn(T, N, 1, generator=g),
        "values": 3.0 * torch.randn(T, N, 1, generator=g),
        "neglogp": normal_neglogp(mu, sigma, actions),
        "dones": (torch.rand(T, N, generator=g) < done_p).to(torch.uint8),
        "actions": actions,
        "mu": mu,
        "sigma": sigma,
        "returns": torch.zeros(T, N, 1),
        "obses": {"obs": torch.randn(T, N, O, generator=g) * scale + shift},
    }
    last_values = 3.0 * torch.randn(N, 1, generator=g)
    return storage, last_values


def init_actor_critic_params(O, A, units, norm, fixed_sigma, seed=0):
    """Orthogonal init with the reference's gains (models.py:8-17,75-91); names match
    the reference state_dict."""
    g = torch.Generator().manual_seed(seed)
    p = {}

    def lin(name, i, o, gain):
        w = torch.empty(o, i)
        torch.nn.init.orthogonal_(w, gain=gain, generator=g)
        p[f"{name}.weight"], p[f"{name}.bias"] = w, torch.zeros(o)

    stride = 2 + (1 if norm else 0)
    for tower in ("actor_mlp", "value_mlp"):
        i = O
        for li, u in enumerate(units):
            lin(f"{tower}.mlp.{li * stride}", i, u, 1.41421356237)
            if norm:
                p[f"{tower}.mlp.{li * stride + 1}.weight"] = torch.ones(u)
                p[f"{tower}.mlp.{li * stride + 1}.bias"] = torch.zeros(u)
            i = u
    lin("value", units[-1], 1, 1.0)
    lin("mu", units[-1], A, 0.01)
    if fixed_sigma:
        p["sigma"] = torch.zeros(A)
    else:
        lin("sigma", units[-1], A, 0.01)
    return p


def np_permutation(n, seed):
    """`np.random.permutation` as used by CriticDataset.shuffle (diffrl/utils.py:55-58)."""
    return np.random.RandomState(seed).permutation(n)
