"""Exploration-noise schedules of DDPG (mineral/agents/ddpg/schedule_util.py:13-53): host-side scalars."""

import math


class LinearSchedule:
    def __init__(self, start_val, end_val, total_iters=5):
        self.start_val, self.end_val, self.total_iters = start_val, end_val, total_iters
        self.count = 0
        self.last_val = start_val

    def step(self):
        if self.count > self.total_iters:
            return self.last_val
        self.last_val = self.count / self.total_iters * (self.end_val - self.start_val) + self.start_val
        self.count += 1
        return self.last_val

    def val(self):
        return self.last_val


class ExponentialSchedule:
    def __init__(self, start_val, gamma, end_val=None):
        self.start_val, self.end_val, self.gamma = start_val, end_val, gamma
        self.total_iters = int((math.log(end_val) - math.log(start_val)) / math.log(gamma)) if end_val is not None else None
        self.count = 0
        self.last_val = start_val

    def step(self):
        if self.total_iters is not None and self.count > self.total_iters:
            return self.last_val
        self.last_val = self.last_val * self.gamma
        self.count += 1
        return self.last_val

    def val(self):
        return self.last_val
