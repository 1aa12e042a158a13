"""Attribute-access dict standing in for an OmegaConf DictConfig (the agents read their configuration both ways:
`cfg.agent.ppo`, `cfg['e_clip']`, `cfg.get('min_lr', 1e-6)`; mineral/agents/ppo/ppo.py:23-26,63-82)."""


class AttrDict(dict):
    def __getattr__(self, key):
        try:
            return self[key]
        except KeyError as e:
            raise AttributeError(key) from e

    def __setattr__(self, key, value):
        self[key] = value


def to_attr(x):
    """Nested dicts -> nested AttrDicts (lists and scalars are kept as they are)."""
    if isinstance(x, dict):
        return AttrDict({k: to_attr(v) for k, v in x.items()})
    return x
