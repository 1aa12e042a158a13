from .diffrl.shac import SHAC
from .ppo.ppo import PPO
