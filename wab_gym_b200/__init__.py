"""wab_gym_b200 — B200-native batched simulator for the Wolves-and-Bushes grid world.

Public surface (mirrors ``/root/reference/wab_env.py``):
  * ``default_game_options`` — same keys/values as the reference (``wab_env.py:11-39``)
  * ``WolvesAndBushesEnv``   — single-env gym-style compat class (``reset()``, ``step(a)``)
  * ``VecEnv``               — N lockstep envs on device-resident torch tensors

The compute path is hand-written sm_100a CUDA behind the C ABI of ``include/wab_b200.h``; there is
no CPU fallback, and importing the env classes fails loudly if the library cannot be loaded.
"""
from .config import GameConfig, default_game_options  # noqa: F401

__all__ = ["GameConfig", "default_game_options", "VecEnv", "ObsBatch", "WolvesAndBushesEnv"]


def __getattr__(name):  # lazy: `import wab_gym_b200.config` must work without torch / CUDA
    if name in ("VecEnv", "ObsBatch"):
        from . import vec_env
        return getattr(vec_env, name)
    if name == "WolvesAndBushesEnv":
        from .env import WolvesAndBushesEnv
        return WolvesAndBushesEnv
    raise AttributeError(name)
