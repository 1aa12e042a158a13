from .ddpg.ddpg import DDPG
from .diffrl.shac import SHAC
from .ppo.ppo import PPO
from .sac.sac import SAC
