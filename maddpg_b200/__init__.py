"""maddpg_b200: B200-native (sm_100a) implementation of MADDPG's batched rollout/update hot path.

Drop-in surface:
  * ``MADDPGAgentTrainer`` -- maddpg/trainer/maddpg.py:112-196 (action / experience / preupdate / update)
  * ``BatchedMultiAgentEnv`` / ``make_env`` -- multiagent.environment.MultiAgentEnv as driven by
    experiments/train.py:48-61,104,114,128
  * ``DeviceReplayBuffer`` -- maddpg/trainer/replay_buffer.py
All compute is in libmaddpg_b200.so (include/maddpg_b200.h); importing this package fails loudly if
the library has not been built.
"""
from . import _lib  # noqa: F401  (raises ImportError when the CUDA library is missing)
from .env import BatchedMultiAgentEnv, make_env  # noqa: F401
from .replay import DevicePrioritizedReplayMemory, DeviceReplayBuffer, JointReplayRing  # noqa: F401
from .trainer import AgentTrainer, MADDPGAgentTrainer, MADDPGCore  # noqa: F401

__all__ = ["BatchedMultiAgentEnv", "make_env", "DeviceReplayBuffer", "DevicePrioritizedReplayMemory", "JointReplayRing", "AgentTrainer",
           "MADDPGAgentTrainer", "MADDPGCore"]
