"""Device-resident lockstep rollout: the batched form of experiments/train.py:110-133
(act -> env.step -> experience -> reset every max_episode_len steps) with no host round trip.

Per lockstep step: one grouped actor+Gumbel kernel (all agents), one fused env-step kernel, one
replay-insert kernel; every ``max_episode_len`` steps a device reset (all env instances share the
episode counter).  A whole episode can be captured into a CUDA graph."""
import torch

from . import _lib


class BatchedRollout(object):
    def __init__(self, env, core, max_episode_len=25, use_graph=True):
        assert env.obs_dims == core.obs_dims and env.act_dims == core.act_dims
        self.env, self.core = env, core
        self.max_episode_len = int(max_episode_len)
        self.episode_step = 0
        self.total_steps = 0
        self.use_graph = use_graph
        self._graph = None
        self.graph_ok = False
        self.ep_return = torch.zeros((env.num_envs, env.n), dtype=torch.float32, device=env.device)

    def step(self):
        """train.py:112-133 for all env instances."""
        env, core = self.env, self.core
        core.act(env.obs, env.act)
        env.step_device(ring=core.ring)
        self.episode_step += 1
        self.total_steps += 1
        if self.episode_step >= self.max_episode_len:
            env.reset()
            self.episode_step = 0

    def run(self, steps):
        for _ in range(steps):
            self.step()

    @property
    def agent_steps_per_step(self):
        return self.env.num_envs * self.env.n
