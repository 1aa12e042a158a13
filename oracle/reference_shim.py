"""Import the UNMODIFIED reference (jacr13/mineral) from /root/reference on CPU.

TEST INFRASTRUCTURE ONLY.  Used by `oracle/make_golden.py` (and by optional
cross-checks in `tests/` when `/root/reference` exists, i.e. in the build
container only).  Nothing under `mineral_b200/` imports this file.

The reference needs `omegaconf` only for one `OmegaConf.to_container` call
(`mineral/agents/agent.py:7,65`); it is not installed in this image, so a tiny
stand-in module is registered before the first `import mineral.agents`.
"""

import os
import sys
import types

def _find_reference():
    """$MINERAL_REFERENCE_ROOT, else /root/reference (build container), else the staged copy oracle/_ref (GPU box;
    made by oracle/build_ref.py, git-ignored)."""
    here = os.path.dirname(os.path.abspath(__file__))
    for root in (os.environ.get("MINERAL_REFERENCE_ROOT"), "/root/reference", os.path.join(here, "_ref")):
        if root and os.path.isdir(os.path.join(root, "mineral")):
            return root
    return "/root/reference"


REFERENCE_ROOT = _find_reference()


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "mineral"))


class AttrDict(dict):
    """dict with attribute access, standing in for an OmegaConf DictConfig."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def to_attr(x):
    if isinstance(x, dict):
        return AttrDict({k: to_attr(v) for k, v in x.items()})
    return x


def to_plain(x):
    if isinstance(x, dict):
        return {k: to_plain(v) for k, v in x.items()}
    return x


def install():
    """Make `import mineral` resolve to the reference; idempotent."""
    if not reference_available():
        raise RuntimeError(f"reference not found at {REFERENCE_ROOT}")
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    if "omegaconf" not in sys.modules:
        oc = types.ModuleType("omegaconf")

        class OmegaConf:  # noqa: D401 - minimal stand-in
            @staticmethod
            def to_container(cfg, resolve=True, **kw):
                return to_plain(cfg)

        oc.OmegaConf = OmegaConf
        oc.DictConfig = AttrDict
        sys.modules["omegaconf"] = oc


class _Space:
    def __init__(self, shape):
        self.shape = tuple(shape)


class SynthEnv:
    """Deterministic synthetic vector env (no simulator): obs/reward/done streams
    are pre-drawn from a seeded generator so a rollout is reproducible."""

    def __init__(self, num_envs, obs_dim, act_dim, horizon, seed=0, done_p=0.02, timeout_p=0.01):
        import torch

        g = torch.Generator().manual_seed(seed)
        self.num_envs, self.obs_dim, self.act_dim = num_envs, obs_dim, act_dim
        self.observation_space = _Space((obs_dim,))
        self.action_space = _Space((act_dim,))
        scale = torch.empty(obs_dim).uniform_(0.1, 10.0, generator=g)
        shift = torch.empty(obs_dim).uniform_(-3.0, 3.0, generator=g)
        n = horizon + 2
        self._obs = torch.randn(n, num_envs, obs_dim, generator=g) * scale + shift
        self._rew = torch.randn(n, num_envs, generator=g)
        self._done = torch.rand(n, num_envs, generator=g) < done_p
        self._tout = torch.rand(n, num_envs, generator=g) < timeout_p
        self._t = 0

    def reset(self):
        self._t = 0
        return self._obs[0].clone()

    def step(self, actions):
        self._t += 1
        t = self._t
        return (
            self._obs[t].clone(),
            self._rew[t].clone(),
            self._done[t].clone(),
            {"time_outs": self._tout[t].clone()},
        )
