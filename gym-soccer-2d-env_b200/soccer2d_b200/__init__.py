"""soccer2d_b200 - B200-native batched lockstep simulator behind the gym API of
CLSFramework/gym-soccer-2d-env (Soccer2DEnv.step/reset).  Host code is Python/PyTorch over the C ABI of
libsoccer2d.so (include/soccer2d.h); the hot path is hand-written CUDA for sm_100a (csrc/)."""
from . import _abi
from ._abi import Soccer2DError
from .sharding import StatsFuture, allreduce_stats, allreduce_stats_async, shard_range
from .vec_env import REACHBALL_DEFAULTS, Soccer2DVecEnv

__all__ = ["Soccer2DVecEnv", "Soccer2DError", "REACHBALL_DEFAULTS", "allreduce_stats", "allreduce_stats_async", "StatsFuture",
           "shard_range", "_abi"]
