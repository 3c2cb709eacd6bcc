"""gym_art_b200: B200-native (sm_100a) batched quadrotor simulator with gym_art's API on the env-step path.

    from gym_art_b200 import QuadVecEnv                          # batched, torch tensors
    from gym_art_b200.quadrotor.quadrotor import QuadrotorEnv    # reference-shaped single-env facade

The arithmetic lives in libquadsim.so (hand-written CUDA, C-ABI in include/quadsim.h); build it with
`python -m gym_art_b200._build`.  Importing the package does not need a GPU; stepping an env does.
"""
from .vec_env import QuadVecEnv, make_config  # noqa: F401

__all__ = ["QuadVecEnv", "make_config"]
