"""Import-compatibility shim: `import SPaRC_Gym` + `gym.make("SPaRC-Gym", ...)` keep working when the
reference package is replaced by the B200 build (the env class and registration live in
sparc_gym_b200.gym_env)."""
from sparc_gym_b200.gym_env import SPaRC_Gym, make, register_envs

register_envs()
__all__ = ["SPaRC_Gym", "make", "register_envs"]
