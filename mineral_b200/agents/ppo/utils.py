"""Host-side LR schedules (mineral/agents/ppo/utils.py:7-37).  The KL-adaptive rule also exists as a
device kernel (mnr_kl_lr_update) so the update loop needs no host sync; this class is kept for API
parity and for tests that compare the two."""

import math


class AdaptiveScheduler:
    def __init__(self, kl_threshold=0.008, min_lr=1e-6, max_lr=1e-2):
        self.min_lr, self.max_lr, self.kl_threshold = min_lr, max_lr, kl_threshold

    def update(self, current_lr, kl_dist):
        lr = current_lr
        if kl_dist > 2.0 * self.kl_threshold:
            lr = max(current_lr / 1.5, self.min_lr)
        if kl_dist < 0.5 * self.kl_threshold:
            lr = min(current_lr * 1.5, self.max_lr)
        return lr


class LinearScheduler:
    def __init__(self, start_lr, min_lr=1e-6, max_steps=1000000):
        self.start_lr, self.min_lr, self.max_steps = start_lr, min_lr, max_steps

    def update(self, steps):
        return max(self.min_lr, self.start_lr - self.start_lr * (steps / float(self.max_steps)))


def adjust_learning_rate_cos(init_lr, epoch, mini_epochs, agent_steps, max_agent_steps):
    return init_lr * 0.5 * (1.0 + math.cos(math.pi * (agent_steps + epoch / mini_epochs) / max_agent_steps))
