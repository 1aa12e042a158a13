"""mineral_b200: B200-native (sm_100a) implementation of jacr13/mineral's rollout->update data path.

The package mirrors the reference's module layout for the classes on that path
(`mineral_b200.agents.PPO`, `mineral_b200.agents.ppo.experience.ExperienceBuffer`,
`mineral_b200.common.normalizers.RunningMeanStd`, `mineral_b200.buffers.ReplayBuffer`, ...), so a
reference user swaps `mineral` for `mineral_b200` in the import.  All arithmetic on the path runs in
hand-written CUDA kernels behind the C ABI in `include/mineral_b200.h`; importing the package loads
`lib/libmineral_b200.so` and fails if it is not built.
"""

from . import _native  # noqa: F401  (loads the shared library; ImportError if missing)

__version__ = "0.1.0"
