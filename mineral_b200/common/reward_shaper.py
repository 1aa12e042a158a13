class RewardShaper:
    """Scale-only reward shaping (reference: mineral/common/reward_shaper.py:1-15).  Host-side scalar
    bookkeeping; the multiply itself rides on whatever tensor op the caller applies."""

    def __init__(self, fn='scale', scale=1.0):
        if fn != 'scale':
            raise NotImplementedError(fn)
        self.fn, self.scale = fn, scale

    def __call__(self, rewards):
        return rewards * self.scale
