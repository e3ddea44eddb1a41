"""avoid-mpc_b200: B200-native batched collision-avoidance MPC hot path
(k-NN over depth clouds + quadrotor NLP solve) behind the call surface of
SJTU-ViSYS-team/Avoid-MPC.  See DESIGN.md."""
from . import capi, defaults, multi, shard, synth  # noqa: F401
from .capi import AmpcError, Handle  # noqa: F401
