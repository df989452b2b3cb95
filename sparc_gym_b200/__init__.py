"""sparc_gym_b200 -- B200-native batched SPaRC environment.

The hot path of tobiTKM/SPaRC-Gym (reset -> step -> terminal rule validation -> one-hot 'new'
observation) as hand-written sm_100a CUDA behind a C ABI (include/sparc_b200.h), with the
reference's Gymnasium surface on top:

    SparcVecEnv      N lockstep envs on one GPU; SparcVectorEnv = its gymnasium.vector-shaped face (vec_env.py)
    ShardedRunner    env-sharded multi-GPU runner, NCCL all-reduce of episode statistics (runner.py)
    SPaRC_Gym        single-env drop-in for gym.make('SPaRC-Gym' | 'SPaRc-Gym') (gym_env.py)

Importing the package registers the env ids when gymnasium is installed (the reference does the
same on `import SPaRC_Gym`, SPaRC_Gym/__init__.py:10-11).  There is no CPU fallback.
"""
from . import dataset  # noqa: F401
from .gym_env import SPaRC_Gym, make, register_envs

register_envs()

__all__ = ["SPaRC_Gym", "make", "register_envs", "dataset"]
