from .nstep_replay import NStepReplay
from .simple_replay import ReplayBuffer
