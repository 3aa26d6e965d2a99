"""CPU oracle: numpy float64 restatement of the Multi-Agent Particle Environment (MPE).

TEST INFRASTRUCTURE ONLY.  Nothing under ``maddpg_b200/`` may import this module; it is
used by ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` as the checker / timed CPU baseline, never as a product path.

PARITY UNPINNED.  The reference (adolfogonzalez3/maddpg) does not contain the environment:
it imports the third-party package ``multiagent`` (openai/multiagent-particle-envs, *no pinned
version*: reference README.md:8,23-26) at experiments/train.py:49-50 and calls
``MultiAgentEnv(world, reset_world, reward, observation)`` (train.py:58-60), ``env.reset()``
(train.py:104,128) and ``env.step(action_n)`` (train.py:114).  That package is not installed in
the build container and cannot be fetched, and the reference has no test or golden vector for
it, so this file restates the published upstream algorithm (``multiagent/core.py``,
``multiagent/environment.py``, ``multiagent/scenarios/{simple,simple_spread,simple_tag,
simple_world_comm}.py`` at upstream ``master``) as specified in SURVEY.md Appendix A.

Structure mirrors upstream on purpose (one Python object per entity, the O(E^2) pair loop, the
per-agent observation/reward callbacks) so that timing it is a fair stand-in for the
reference's CPU environment path.  State is float64 like upstream; actions are float32 like the
trainer's output (maddpg/trainer/maddpg.py:151-152).
"""
import numpy as np


# --------------------------------------------------------------------------------------------
# spaces (gym 0.10.5 is absent; only the attributes the reference reads are provided:
# maddpg/common/distributions.py:408-422 reads ``.n`` / ``.low`` / ``.high``; train.py:83 reads
# ``observation_space[i].shape``)
# --------------------------------------------------------------------------------------------
class Discrete:
    def __init__(self, n):
        self.n = int(n)

    def __repr__(self):
        return "Discrete(%d)" % self.n


class MultiDiscrete:
    """MPE's own ``multiagent.multi_discrete.MultiDiscrete``: list of [min, max] pairs."""

    def __init__(self, array_of_param_array):
        self.low = np.array([x[0] for x in array_of_param_array])
        self.high = np.array([x[1] for x in array_of_param_array])
        self.num_discrete_space = self.low.shape[0]

    def __repr__(self):
        return "MultiDiscrete(%s)" % [[int(l), int(h)] for l, h in zip(self.low, self.high)]


class Box:
    def __init__(self, shape):
        self.shape = tuple(shape)
        self.low = -np.inf
        self.high = np.inf


# --------------------------------------------------------------------------------------------
# multiagent/core.py
# --------------------------------------------------------------------------------------------
class EntityState:
    def __init__(self):
        self.p_pos = None
        self.p_vel = None


class AgentState(EntityState):
    def __init__(self):
        super().__init__()
        self.c = None


class Action:
    def __init__(self):
        self.u = None
        self.c = None


class Entity:
    def __init__(self):
        self.name = ""
        self.size = 0.050
        self.movable = False
        self.collide = True
        self.max_speed = None
        self.accel = None
        self.state = EntityState()
        self.initial_mass = 1.0

    @property
    def mass(self):
        return self.initial_mass


class Landmark(Entity):
    def __init__(self):
        super().__init__()
        self.boundary = False


class Agent(Entity):
    def __init__(self):
        super().__init__()
        self.movable = True
        self.silent = False
        self.u_noise = None
        self.c_noise = None
        self.u_range = 1.0
        self.state = AgentState()
        self.action = Action()
        self.action_callback = None
        self.adversary = False
        self.leader = False


class World:
    """multiagent.core.World."""

    def __init__(self):
        self.agents = []
        self.landmarks = []
        self.dim_c = 0
        self.dim_p = 2
        self.dt = 0.1
        self.damping = 0.25
        self.contact_force = 1e+2
        self.contact_margin = 1e-3
        self.collaborative = False

    @property
    def entities(self):
        return self.agents + self.landmarks

    @property
    def policy_agents(self):
        return [a for a in self.agents if a.action_callback is None]

    def step(self):
        p_force = [None] * len(self.entities)
        p_force = self.apply_action_force(p_force)
        p_force = self.apply_environment_force(p_force)
        self.integrate_state(p_force)
        for agent in self.agents:
            self.update_agent_state(agent)

    def apply_action_force(self, p_force):
        for i, agent in enumerate(self.agents):
            if agent.movable:
                p_force[i] = agent.action.u + 0.0  # u_noise is None in every in-scope scenario
        return p_force

    def apply_environment_force(self, p_force):
        entities = self.entities
        for a, entity_a in enumerate(entities):
            for b, entity_b in enumerate(entities):
                if b <= a:
                    continue
                f_a, f_b = self.get_collision_force(entity_a, entity_b)
                if f_a is not None:
                    if p_force[a] is None:
                        p_force[a] = 0.0
                    p_force[a] = f_a + p_force[a]
                if f_b is not None:
                    if p_force[b] is None:
                        p_force[b] = 0.0
                    p_force[b] = f_b + p_force[b]
        return p_force

    def integrate_state(self, p_force):
        for i, entity in enumerate(self.entities):
            if not entity.movable:
                continue
            entity.state.p_vel = entity.state.p_vel * (1 - self.damping)
            if p_force[i] is not None:
                entity.state.p_vel = entity.state.p_vel + (p_force[i] / entity.mass) * self.dt
            if entity.max_speed is not None:
                speed = np.sqrt(np.square(entity.state.p_vel[0]) + np.square(entity.state.p_vel[1]))
                if speed > entity.max_speed:
                    entity.state.p_vel = entity.state.p_vel / speed * entity.max_speed
            entity.state.p_pos = entity.state.p_pos + entity.state.p_vel * self.dt

    def update_agent_state(self, agent):
        if agent.silent:
            agent.state.c = np.zeros(self.dim_c)
        else:
            agent.state.c = agent.action.c + 0.0  # c_noise is None

    def get_collision_force(self, entity_a, entity_b):
        if (not entity_a.collide) or (not entity_b.collide):
            return [None, None]
        if entity_a is entity_b:
            return [None, None]
        delta_pos = entity_a.state.p_pos - entity_b.state.p_pos
        dist = np.sqrt(np.sum(np.square(delta_pos)))
        dist_min = entity_a.size + entity_b.size
        k = self.contact_margin
        penetration = np.logaddexp(0, -(dist - dist_min) / k) * k
        force = self.contact_force * delta_pos / dist * penetration
        force_a = +force if entity_a.movable else None
        force_b = -force if entity_b.movable else None
        return [force_a, force_b]


# --------------------------------------------------------------------------------------------
# scenarios
# --------------------------------------------------------------------------------------------
def _dist(a, b):
    return np.sqrt(np.sum(np.square(a.state.p_pos - b.state.p_pos)))


def _is_collision(a, b):
    return True if _dist(a, b) < a.size + b.size else False


def _bound(x):
    if x < 0.9:
        return 0
    if x < 1.0:
        return (x - 0.9) * 10
    return min(np.exp(2 * x - 2), 10)


class BaseScenario:
    """Every scenario draws its reset state from ``self.rng`` (upstream: the global numpy RNG,
    unseeded in train.py).  Parity tests inject states with ``set_world_state`` instead."""

    name = ""

    def __init__(self, rng=None):
        self.rng = rng if rng is not None else np.random

    def benchmark_data(self, agent, world):
        return {}


class SimpleScenario(BaseScenario):
    name = "simple"

    def make_world(self):
        world = World()
        world.agents = [Agent() for _ in range(1)]
        for i, agent in enumerate(world.agents):
            agent.name = "agent %d" % i
            agent.collide = False
            agent.silent = True
        world.landmarks = [Landmark() for _ in range(1)]
        for i, landmark in enumerate(world.landmarks):
            landmark.name = "landmark %d" % i
            landmark.collide = False
            landmark.movable = False
        self.reset_world(world)
        return world

    def reset_world(self, world):
        for agent in world.agents:
            agent.state.p_pos = self.rng.uniform(-1, +1, world.dim_p)
            agent.state.p_vel = np.zeros(world.dim_p)
            agent.state.c = np.zeros(world.dim_c)
        for landmark in world.landmarks:
            landmark.state.p_pos = self.rng.uniform(-1, +1, world.dim_p)
            landmark.state.p_vel = np.zeros(world.dim_p)

    def reward(self, agent, world):
        dist2 = np.sum(np.square(agent.state.p_pos - world.landmarks[0].state.p_pos))
        return -dist2

    def observation(self, agent, world):
        entity_pos = [e.state.p_pos - agent.state.p_pos for e in world.landmarks]
        return np.concatenate([agent.state.p_vel] + entity_pos)


class SimpleSpreadScenario(BaseScenario):
    name = "simple_spread"

    def __init__(self, rng=None, num_agents=3):
        super().__init__(rng)
        self.num_agents = int(num_agents)

    def make_world(self):
        world = World()
        world.dim_c = 2
        world.collaborative = True
        world.agents = [Agent() for _ in range(self.num_agents)]
        for i, agent in enumerate(world.agents):
            agent.name = "agent %d" % i
            agent.collide = True
            agent.silent = True
            agent.size = 0.15
        world.landmarks = [Landmark() for _ in range(self.num_agents)]
        for i, landmark in enumerate(world.landmarks):
            landmark.name = "landmark %d" % i
            landmark.collide = False
            landmark.movable = False
        self.reset_world(world)
        return world

    def reset_world(self, world):
        for agent in world.agents:
            agent.state.p_pos = self.rng.uniform(-1, +1, world.dim_p)
            agent.state.p_vel = np.zeros(world.dim_p)
            agent.state.c = np.zeros(world.dim_c)
        for landmark in world.landmarks:
            landmark.state.p_pos = self.rng.uniform(-1, +1, world.dim_p)
            landmark.state.p_vel = np.zeros(world.dim_p)

    def reward(self, agent, world):
        rew = 0
        for l in world.landmarks:
            dists = [_dist(a, l) for a in world.agents]
            rew -= min(dists)
        if agent.collide:
            for a in world.agents:
                if _is_collision(a, agent):  # includes a is agent: constant -1
                    rew -= 1
        return rew

    def observation(self, agent, world):
        entity_pos = [e.state.p_pos - agent.state.p_pos for e in world.landmarks]
        comm = []
        other_pos = []
        for other in world.agents:
            if other is agent:
                continue
            comm.append(other.state.c)
            other_pos.append(other.state.p_pos - agent.state.p_pos)
        return np.concatenate([agent.state.p_vel] + [agent.state.p_pos] + entity_pos + other_pos + comm)

    def benchmark_data(self, agent, world):
        # upstream simple_spread.benchmark_data: (reward, collisions, sum of min distances, occupied landmarks)
        rew = 0
        collisions = 0
        occupied_landmarks = 0
        min_dists = 0
        for l in world.landmarks:
            dists = [_dist(a, l) for a in world.agents]
            min_dists += min(dists)
            rew -= min(dists)
            if min(dists) < 0.1:
                occupied_landmarks += 1
        if agent.collide:
            for a in world.agents:
                if _is_collision(a, agent):
                    rew -= 1
                    collisions += 1
        return (rew, collisions, min_dists, occupied_landmarks)


class SimpleTagScenario(BaseScenario):
    name = "simple_tag"

    def make_world(self):
        world = World()
        world.dim_c = 2
        num_good_agents = 1
        num_adversaries = 3
        num_agents = num_adversaries + num_good_agents
        num_landmarks = 2
        world.agents = [Agent() for _ in range(num_agents)]
        for i, agent in enumerate(world.agents):
            agent.name = "agent %d" % i
            agent.collide = True
            agent.silent = True
            agent.adversary = True if i < num_adversaries else False
            agent.size = 0.075 if agent.adversary else 0.05
            agent.accel = 3.0 if agent.adversary else 4.0
            agent.max_speed = 1.0 if agent.adversary else 1.3
        world.landmarks = [Landmark() for _ in range(num_landmarks)]
        for i, landmark in enumerate(world.landmarks):
            landmark.name = "landmark %d" % i
            landmark.collide = True
            landmark.movable = False
            landmark.size = 0.2
            landmark.boundary = False
        self.reset_world(world)
        return world

    def reset_world(self, world):
        for agent in world.agents:
            agent.state.p_pos = self.rng.uniform(-1, +1, world.dim_p)
            agent.state.p_vel = np.zeros(world.dim_p)
            agent.state.c = np.zeros(world.dim_c)
        for landmark in world.landmarks:
            if not landmark.boundary:
                landmark.state.p_pos = self.rng.uniform(-0.9, +0.9, world.dim_p)
                landmark.state.p_vel = np.zeros(world.dim_p)

    def good_agents(self, world):
        return [a for a in world.agents if not a.adversary]

    def adversaries(self, world):
        return [a for a in world.agents if a.adversary]

    def benchmark_data(self, agent, world):
        # upstream simple_tag / simple_world_comm: adversaries report their collisions with good agents, good agents 0
        if agent.adversary:
            collisions = 0
            for a in self.good_agents(world):
                if _is_collision(a, agent):
                    collisions += 1
            return collisions
        return 0

    def reward(self, agent, world):
        return self.adversary_reward(agent, world) if agent.adversary else self.agent_reward(agent, world)

    def agent_reward(self, agent, world):
        rew = 0
        if agent.collide:
            for a in self.adversaries(world):
                if _is_collision(a, agent):
                    rew -= 10
        for p in range(world.dim_p):
            rew -= _bound(abs(agent.state.p_pos[p]))
        return rew

    def adversary_reward(self, agent, world):
        rew = 0
        if agent.collide:
            for ag in self.good_agents(world):
                for adv in self.adversaries(world):
                    if _is_collision(ag, adv):
                        rew += 10
        return rew

    def observation(self, agent, world):
        entity_pos = [e.state.p_pos - agent.state.p_pos for e in world.landmarks if not e.boundary]
        other_pos = []
        other_vel = []
        for other in world.agents:
            if other is agent:
                continue
            other_pos.append(other.state.p_pos - agent.state.p_pos)
            if not other.adversary:
                other_vel.append(other.state.p_vel)
        return np.concatenate([agent.state.p_vel] + [agent.state.p_pos] + entity_pos + other_pos + other_vel)


class SimpleWorldCommScenario(BaseScenario):
    name = "simple_world_comm"

    def make_world(self):
        world = World()
        world.dim_c = 4
        num_good_agents = 2
        num_adversaries = 4
        num_agents = num_adversaries + num_good_agents
        num_landmarks = 1
        num_food = 2
        num_forests = 2
        world.agents = [Agent() for _ in range(num_agents)]
        for i, agent in enumerate(world.agents):
            agent.name = "agent %d" % i
            agent.collide = True
            agent.leader = True if i == 0 else False
            agent.silent = True if i > 0 else False
            agent.adversary = True if i < num_adversaries else False
            agent.size = 0.075 if agent.adversary else 0.045
            agent.accel = 3.0 if agent.adversary else 4.0
            agent.max_speed = 1.0 if agent.adversary else 1.3
        world.landmarks = [Landmark() for _ in range(num_landmarks)]
        for i, landmark in enumerate(world.landmarks):
            landmark.name = "landmark %d" % i
            landmark.collide = True
            landmark.movable = False
            landmark.size = 0.2
            landmark.boundary = False
        world.food = [Landmark() for _ in range(num_food)]
        for i, landmark in enumerate(world.food):
            landmark.name = "food %d" % i
            landmark.collide = False
            landmark.movable = False
            landmark.size = 0.03
            landmark.boundary = False
        world.forests = [Landmark() for _ in range(num_forests)]
        for i, landmark in enumerate(world.forests):
            landmark.name = "forest %d" % i
            landmark.collide = False
            landmark.movable = False
            landmark.size = 0.3
            landmark.boundary = False
        world.landmarks += world.food
        world.landmarks += world.forests
        self.reset_world(world)
        return world

    def reset_world(self, world):
        for agent in world.agents:
            agent.state.p_pos = self.rng.uniform(-1, +1, world.dim_p)
            agent.state.p_vel = np.zeros(world.dim_p)
            agent.state.c = np.zeros(world.dim_c)
        for landmark in world.landmarks:
            landmark.state.p_pos = self.rng.uniform(-0.9, +0.9, world.dim_p)
            landmark.state.p_vel = np.zeros(world.dim_p)
        for landmark in world.food:
            landmark.state.p_pos = self.rng.uniform(-0.9, +0.9, world.dim_p)
            landmark.state.p_vel = np.zeros(world.dim_p)
        for landmark in world.forests:
            landmark.state.p_pos = self.rng.uniform(-0.9, +0.9, world.dim_p)
            landmark.state.p_vel = np.zeros(world.dim_p)

    def good_agents(self, world):
        return [a for a in world.agents if not a.adversary]

    def adversaries(self, world):
        return [a for a in world.agents if a.adversary]

    def benchmark_data(self, agent, world):
        # upstream simple_tag / simple_world_comm: adversaries report their collisions with good agents, good agents 0
        if agent.adversary:
            collisions = 0
            for a in self.good_agents(world):
                if _is_collision(a, agent):
                    collisions += 1
            return collisions
        return 0

    def reward(self, agent, world):
        return self.adversary_reward(agent, world) if agent.adversary else self.agent_reward(agent, world)

    def agent_reward(self, agent, world):
        rew = 0
        if agent.collide:
            for a in self.adversaries(world):
                if _is_collision(a, agent):
                    rew -= 5
        for p in range(world.dim_p):
            rew -= 2 * _bound(abs(agent.state.p_pos[p]))
        for food in world.food:
            if _is_collision(agent, food):
                rew += 2
        rew += 0.05 * min([_dist(food, agent) for food in world.food])
        return rew

    def adversary_reward(self, agent, world):
        rew = 0
        agents = self.good_agents(world)
        adversaries = self.adversaries(world)
        rew -= 0.1 * min([_dist(a, agent) for a in agents])
        if agent.collide:
            for ag in agents:
                for adv in adversaries:
                    if _is_collision(ag, adv):
                        rew += 5
        return rew

    def observation(self, agent, world):
        entity_pos = [e.state.p_pos - agent.state.p_pos for e in world.landmarks if not e.boundary]
        in_forest = [np.array([-1]), np.array([-1])]
        inf1 = False
        inf2 = False
        if _is_collision(agent, world.forests[0]):
            in_forest[0] = np.array([1])
            inf1 = True
        if _is_collision(agent, world.forests[1]):
            in_forest[1] = np.array([1])
            inf2 = True
        other_pos = []
        other_vel = []
        for other in world.agents:
            if other is agent:
                continue
            oth_f1 = _is_collision(other, world.forests[0])
            oth_f2 = _is_collision(other, world.forests[1])
            if (inf1 and oth_f1) or (inf2 and oth_f2) or \
                    (not inf1 and not oth_f1 and not inf2 and not oth_f2) or agent.leader:
                other_pos.append(other.state.p_pos - agent.state.p_pos)
                if not other.adversary:
                    other_vel.append(other.state.p_vel)
            else:
                other_pos.append([0, 0])
                if not other.adversary:
                    other_vel.append([0, 0])
        comm = [world.agents[0].state.c]
        if agent.adversary:  # leader and plain adversaries share the field order
            return np.concatenate([agent.state.p_vel] + [agent.state.p_pos] + entity_pos + other_pos
                                  + other_vel + in_forest + comm)
        return np.concatenate([agent.state.p_vel] + [agent.state.p_pos] + entity_pos + other_pos
                              + in_forest + other_vel)


# --------------------------------------------------------------------------------------------
# the other scenarios `train.py --scenario` can name.  Restated from the published upstream
# files multiagent/scenarios/{simple_adversary,simple_push,simple_speaker_listener,simple_crypto,simple_reference}.py (PARITY UNPINNED like the
# rest of this module: the package is absent from /root/reference and from this container).  reset_world is split into the
# goal draw (np.random.choice(world.landmarks), kept first like upstream) and ``apply_goals`` -- the goal pointers and the
# colours upstream derives from them -- so that parity tests can inject the drawn indices.
# --------------------------------------------------------------------------------------------
class _GoalScenario(BaseScenario):
    n_goal = 1

    def draw_goals(self, world):
        return [int(self.rng.randint(len(world.landmarks))) for _ in range(self.n_goal)]

    def reset_world(self, world):
        self.apply_goals(world, self.draw_goals(world))
        for agent in world.agents:
            agent.state.p_pos = self.rng.uniform(-1, +1, world.dim_p)
            agent.state.p_vel = np.zeros(world.dim_p)
            agent.state.c = np.zeros(world.dim_c)
        for landmark in world.landmarks:
            landmark.state.p_pos = self.rng.uniform(-1, +1, world.dim_p)
            landmark.state.p_vel = np.zeros(world.dim_p)

    def good_agents(self, world):
        return [agent for agent in world.agents if not agent.adversary]

    def adversaries(self, world):
        return [agent for agent in world.agents if agent.adversary]


class SimpleAdversaryScenario(_GoalScenario):
    name = "simple_adversary"

    def make_world(self):
        world = World()
        world.dim_c = 2
        num_agents = 3
        num_adversaries = 1
        num_landmarks = num_agents - 1
        world.agents = [Agent() for _ in range(num_agents)]
        for i, agent in enumerate(world.agents):
            agent.name = 'agent %d' % i
            agent.collide = False
            agent.silent = True
            agent.adversary = True if i < num_adversaries else False
            agent.size = 0.15
        world.landmarks = [Landmark() for _ in range(num_landmarks)]
        for i, landmark in enumerate(world.landmarks):
            landmark.name = 'landmark %d' % i
            landmark.collide = False
            landmark.movable = False
            landmark.size = 0.08
        self.reset_world(world)
        return world

    def apply_goals(self, world, goals):
        world.goals = list(goals)
        goal = world.landmarks[goals[0]]
        for agent in world.agents:
            agent.goal_a = goal

    def reward(self, agent, world):
        return self.adversary_reward(agent, world) if agent.adversary else self.agent_reward(agent, world)

    def agent_reward(self, agent, world):
        # shaped_reward = shaped_adv_reward = True upstream
        adv_rew = sum([np.sqrt(np.sum(np.square(a.state.p_pos - a.goal_a.state.p_pos))) for a in self.adversaries(world)])
        pos_rew = -min([np.sqrt(np.sum(np.square(a.state.p_pos - a.goal_a.state.p_pos))) for a in self.good_agents(world)])
        return pos_rew + adv_rew

    def adversary_reward(self, agent, world):
        return -np.sum(np.square(agent.state.p_pos - agent.goal_a.state.p_pos))

    def observation(self, agent, world):
        entity_pos = [entity.state.p_pos - agent.state.p_pos for entity in world.landmarks]
        other_pos = [other.state.p_pos - agent.state.p_pos for other in world.agents if other is not agent]
        if not agent.adversary:
            return np.concatenate([agent.goal_a.state.p_pos - agent.state.p_pos] + entity_pos + other_pos)
        return np.concatenate(entity_pos + other_pos)


class SimplePushScenario(_GoalScenario):
    name = "simple_push"

    def make_world(self):
        world = World()
        world.dim_c = 2
        num_agents = 2
        num_adversaries = 1
        num_landmarks = 2
        world.agents = [Agent() for _ in range(num_agents)]
        for i, agent in enumerate(world.agents):
            agent.name = 'agent %d' % i
            agent.collide = True
            agent.silent = True
            agent.adversary = True if i < num_adversaries else False
        world.landmarks = [Landmark() for _ in range(num_landmarks)]
        for i, landmark in enumerate(world.landmarks):
            landmark.name = 'landmark %d' % i
            landmark.collide = False
            landmark.movable = False
        self.reset_world(world)
        return world

    def apply_goals(self, world, goals):
        world.goals = list(goals)
        for i, landmark in enumerate(world.landmarks):
            landmark.color = np.array([0.1, 0.1, 0.1])
            landmark.color[i + 1] += 0.8
            landmark.index = i
        goal = world.landmarks[goals[0]]
        for agent in world.agents:
            agent.goal_a = goal
            agent.color = np.array([0.25, 0.25, 0.25])
            if agent.adversary:
                agent.color = np.array([0.75, 0.25, 0.25])
            else:
                agent.color[goal.index + 1] += 0.5

    def reward(self, agent, world):
        return self.adversary_reward(agent, world) if agent.adversary else self.agent_reward(agent, world)

    def agent_reward(self, agent, world):
        return -np.sqrt(np.sum(np.square(agent.state.p_pos - agent.goal_a.state.p_pos)))

    def adversary_reward(self, agent, world):
        agent_dist = [np.sqrt(np.sum(np.square(a.state.p_pos - a.goal_a.state.p_pos))) for a in world.agents if not a.adversary]
        pos_rew = min(agent_dist)
        neg_rew = np.sqrt(np.sum(np.square(agent.goal_a.state.p_pos - agent.state.p_pos)))
        return pos_rew - neg_rew

    def observation(self, agent, world):
        entity_pos = [entity.state.p_pos - agent.state.p_pos for entity in world.landmarks]
        entity_color = [entity.color for entity in world.landmarks]
        other_pos = [other.state.p_pos - agent.state.p_pos for other in world.agents if other is not agent]
        if not agent.adversary:
            return np.concatenate([agent.state.p_vel] + [agent.goal_a.state.p_pos - agent.state.p_pos] + [agent.color] +
                                  entity_pos + entity_color + other_pos)
        return np.concatenate([agent.state.p_vel] + entity_pos + other_pos)


class SimpleSpeakerListenerScenario(_GoalScenario):
    name = "simple_speaker_listener"

    def make_world(self):
        world = World()
        world.dim_c = 3
        num_landmarks = 3
        world.collaborative = True
        world.agents = [Agent() for _ in range(2)]
        for i, agent in enumerate(world.agents):
            agent.name = 'agent %d' % i
            agent.collide = False
            agent.size = 0.075
        world.agents[0].movable = False  # speaker
        world.agents[1].silent = True    # listener
        world.landmarks = [Landmark() for _ in range(num_landmarks)]
        for i, landmark in enumerate(world.landmarks):
            landmark.name = 'landmark %d' % i
            landmark.collide = False
            landmark.movable = False
            landmark.size = 0.04
        self.reset_world(world)
        return world

    def apply_goals(self, world, goals):
        world.goals = list(goals)
        for agent in world.agents:
            agent.goal_a = None
            agent.goal_b = None
        world.agents[0].goal_a = world.agents[1]
        world.agents[0].goal_b = world.landmarks[goals[0]]
        world.landmarks[0].color = np.array([0.65, 0.15, 0.15])
        world.landmarks[1].color = np.array([0.15, 0.65, 0.15])
        world.landmarks[2].color = np.array([0.15, 0.15, 0.65])

    def reward(self, agent, world):
        a = world.agents[0]
        dist2 = np.sum(np.square(a.goal_a.state.p_pos - a.goal_b.state.p_pos))
        return -dist2

    def observation(self, agent, world):
        goal_color = np.zeros(3)
        if agent.goal_b is not None:
            goal_color = agent.goal_b.color
        entity_pos = [entity.state.p_pos - agent.state.p_pos for entity in world.landmarks]
        comm = []
        for other in world.agents:
            if other is agent or (other.state.c is None):
                continue
            comm.append(other.state.c)
        if not agent.movable:   # speaker
            return np.concatenate([goal_color])
        return np.concatenate([agent.state.p_vel] + entity_pos + comm)  # listener (agent.silent)


class SimpleReferenceScenario(_GoalScenario):
    name = "simple_reference"
    n_goal = 2  # goal_b of agent 0, goal_b of agent 1

    def make_world(self):
        world = World()
        world.dim_c = 10
        world.collaborative = True
        world.agents = [Agent() for _ in range(2)]
        for i, agent in enumerate(world.agents):
            agent.name = 'agent %d' % i
            agent.collide = False
        world.landmarks = [Landmark() for _ in range(3)]
        for i, landmark in enumerate(world.landmarks):
            landmark.name = 'landmark %d' % i
            landmark.collide = False
            landmark.movable = False
        self.reset_world(world)
        return world

    def apply_goals(self, world, goals):
        world.goals = list(goals)
        world.agents[0].goal_a = world.agents[1]
        world.agents[0].goal_b = world.landmarks[goals[0]]
        world.agents[1].goal_a = world.agents[0]
        world.agents[1].goal_b = world.landmarks[goals[1]]
        world.landmarks[0].color = np.array([0.75, 0.25, 0.25])
        world.landmarks[1].color = np.array([0.25, 0.75, 0.25])
        world.landmarks[2].color = np.array([0.25, 0.25, 0.75])

    def reward(self, agent, world):
        if agent.goal_a is None or agent.goal_b is None:
            return 0.0
        dist2 = np.sum(np.square(agent.goal_a.state.p_pos - agent.goal_b.state.p_pos))
        return -dist2

    def observation(self, agent, world):
        goal_color = [np.zeros(3), np.zeros(3)]
        if agent.goal_b is not None:
            goal_color[1] = agent.goal_b.color
        entity_pos = [entity.state.p_pos - agent.state.p_pos for entity in world.landmarks]
        comm = [other.state.c for other in world.agents if other is not agent]
        return np.concatenate([agent.state.p_vel] + entity_pos + [goal_color[1]] + comm)


class SimpleCryptoScenario(_GoalScenario):
    name = "simple_crypto"
    n_goal = 2  # goal landmark, key landmark

    def make_world(self):
        world = World()
        num_agents = 3
        num_adversaries = 1
        num_landmarks = 2
        world.dim_c = 4
        world.agents = [Agent() for _ in range(num_agents)]
        for i, agent in enumerate(world.agents):
            agent.name = 'agent %d' % i
            agent.collide = False
            agent.adversary = True if i < num_adversaries else False
            agent.speaker = True if i == 2 else False
            agent.movable = False
        world.landmarks = [Landmark() for _ in range(num_landmarks)]
        for i, landmark in enumerate(world.landmarks):
            landmark.name = 'landmark %d' % i
            landmark.collide = False
            landmark.movable = False
        self.reset_world(world)
        return world

    def apply_goals(self, world, goals):
        world.goals = list(goals)
        color_list = [np.zeros(world.dim_c) for _ in world.landmarks]
        for i, color in enumerate(color_list):
            color[i] += 1
        for color, landmark in zip(color_list, world.landmarks):
            landmark.color = color
        goal = world.landmarks[goals[0]]
        for agent in world.agents:
            agent.key = None
            agent.goal_a = goal
        world.agents[2].key = world.landmarks[goals[1]].color

    def good_listeners(self, world):
        return [agent for agent in world.agents if not agent.adversary and not agent.speaker]

    def reward(self, agent, world):
        return self.adversary_reward(agent, world) if agent.adversary else self.agent_reward(agent, world)

    def agent_reward(self, agent, world):
        good_rew, adv_rew = 0, 0
        for a in self.good_listeners(world):
            if (a.state.c == np.zeros(world.dim_c)).all():
                continue
            good_rew -= np.sum(np.square(a.state.c - agent.goal_a.color))
        for a in self.adversaries(world):
            if (a.state.c == np.zeros(world.dim_c)).all():
                continue
            adv_rew += np.sum(np.square(a.state.c - agent.goal_a.color))
        return adv_rew + good_rew

    def adversary_reward(self, agent, world):
        rew = 0
        if not (agent.state.c == np.zeros(world.dim_c)).all():
            rew -= np.sum(np.square(agent.state.c - agent.goal_a.color))
        return rew

    def observation(self, agent, world):
        goal_color = np.zeros(world.dim_c)
        if agent.goal_a is not None:
            goal_color = agent.goal_a.color
        comm = []
        for other in world.agents:
            if other is agent or (other.state.c is None) or not other.speaker:
                continue
            comm.append(other.state.c)
        key = world.agents[2].key
        if agent.speaker:
            return np.concatenate([goal_color] + [key])
        if not agent.adversary:
            return np.concatenate([key] + comm)
        return np.concatenate(comm)


def make_scenario(name, rng=None, num_agents=None):
    """``scenarios.load(name + ".py").Scenario()`` (train.py:53)."""
    if name == "simple":
        return SimpleScenario(rng)
    if name == "simple_spread":
        return SimpleSpreadScenario(rng, 3 if num_agents is None else num_agents)
    if name == "simple_tag":
        return SimpleTagScenario(rng)
    if name == "simple_world_comm":
        return SimpleWorldCommScenario(rng)
    extra = {"simple_adversary": SimpleAdversaryScenario, "simple_push": SimplePushScenario,
             "simple_speaker_listener": SimpleSpeakerListenerScenario, "simple_crypto": SimpleCryptoScenario,
             "simple_reference": SimpleReferenceScenario}
    if name in extra:
        return extra[name](rng)
    raise NotImplementedError(name)


# --------------------------------------------------------------------------------------------
# multiagent/environment.py
# --------------------------------------------------------------------------------------------
class MultiAgentEnv:
    """multiagent.environment.MultiAgentEnv."""

    def __init__(self, world, reset_callback=None, reward_callback=None, observation_callback=None,
                 info_callback=None, done_callback=None):
        self.world = world
        self.agents = self.world.policy_agents
        self.n = len(world.policy_agents)
        self.reset_callback = reset_callback
        self.reward_callback = reward_callback
        self.observation_callback = observation_callback
        self.info_callback = info_callback
        self.done_callback = done_callback
        self.discrete_action_space = True
        self.discrete_action_input = False
        self.force_discrete_action = False
        self.shared_reward = world.collaborative
        self.time = 0
        self.action_space = []
        self.observation_space = []
        for agent in self.agents:
            total_action_space = []
            if agent.movable:
                total_action_space.append(Discrete(world.dim_p * 2 + 1))
            if not agent.silent:
                total_action_space.append(Discrete(world.dim_c))
            if len(total_action_space) > 1:
                self.action_space.append(MultiDiscrete([[0, s.n - 1] for s in total_action_space]))
            else:
                self.action_space.append(total_action_space[0])
            obs_dim = len(observation_callback(agent, self.world))
            self.observation_space.append(Box((obs_dim,)))
            agent.action.c = np.zeros(self.world.dim_c)

    def step(self, action_n):
        obs_n, reward_n, done_n, info_n = [], [], [], {'n': []}
        self.agents = self.world.policy_agents
        for i, agent in enumerate(self.agents):
            self._set_action(action_n[i], agent, self.action_space[i])
        self.world.step()
        for agent in self.agents:
            obs_n.append(self._get_obs(agent))
            reward_n.append(self._get_reward(agent))
            done_n.append(self._get_done(agent))
            info_n['n'].append(self._get_info(agent))
        reward = np.sum(reward_n)
        if self.shared_reward:
            reward_n = [reward] * self.n
        return obs_n, reward_n, done_n, info_n

    def reset(self):
        self.reset_callback(self.world)
        self.agents = self.world.policy_agents
        return [self._get_obs(agent) for agent in self.agents]

    def _get_info(self, agent):
        return {} if self.info_callback is None else self.info_callback(agent, self.world)

    def _get_obs(self, agent):
        return self.observation_callback(agent, self.world)

    def _get_done(self, agent):
        return False if self.done_callback is None else self.done_callback(agent, self.world)

    def _get_reward(self, agent):
        return self.reward_callback(agent, self.world)

    def _set_action(self, action, agent, action_space):
        agent.action.u = np.zeros(self.world.dim_p)
        agent.action.c = np.zeros(self.world.dim_c)
        if isinstance(action_space, MultiDiscrete):
            act = []
            size = action_space.high - action_space.low + 1
            index = 0
            for s in size:
                act.append(action[index:(index + s)])
                index += s
            action = act
        else:
            action = [action]
        if agent.movable:
            # float32 difference (the trainer returns float32) stored into a float64 vector
            agent.action.u[0] += action[0][1] - action[0][2]
            agent.action.u[1] += action[0][3] - action[0][4]
            sensitivity = 5.0
            if agent.accel is not None:
                sensitivity = agent.accel
            agent.action.u *= sensitivity
            action = action[1:]
        if not agent.silent:
            agent.action.c = action[0]
            action = action[1:]
        assert len(action) == 0

    def render(self, mode="human"):
        raise NotImplementedError("rendering is out of scope")


def make_env(scenario_name, rng=None, num_agents=None, benchmark=False):
    """experiments/train.py:48-61."""
    scenario = make_scenario(scenario_name, rng, num_agents)
    world = scenario.make_world()
    if benchmark:
        env = MultiAgentEnv(world, scenario.reset_world, scenario.reward, scenario.observation,
                            scenario.benchmark_data)
    else:
        env = MultiAgentEnv(world, scenario.reset_world, scenario.reward, scenario.observation)
    env.scenario = scenario
    return env


# --------------------------------------------------------------------------------------------
# state injection / extraction helpers used by the parity tests (not part of upstream)
# --------------------------------------------------------------------------------------------
def get_world_state(world):
    """Returns dict(agent_pos (A,2), agent_vel (A,2), agent_c (A,dim_c), landmark_pos (L,2))."""
    return dict(
        agent_pos=np.array([a.state.p_pos for a in world.agents], dtype=np.float64),
        agent_vel=np.array([a.state.p_vel for a in world.agents], dtype=np.float64),
        agent_c=np.array([a.state.c for a in world.agents], dtype=np.float64).reshape(len(world.agents), world.dim_c),
        landmark_pos=np.array([l.state.p_pos for l in world.landmarks], dtype=np.float64),
        goal=np.array(getattr(world, "goals", []), dtype=np.int64),
    )


def set_world_state(world, agent_pos, agent_vel, landmark_pos, agent_c=None, goal=None, scenario=None):
    if goal is not None and len(goal):
        scenario.apply_goals(world, [int(g) for g in goal])
    for i, a in enumerate(world.agents):
        a.state.p_pos = np.array(agent_pos[i], dtype=np.float64)
        a.state.p_vel = np.array(agent_vel[i], dtype=np.float64)
        a.state.c = np.zeros(world.dim_c) if agent_c is None else np.array(agent_c[i], dtype=np.float64)
    for i, l in enumerate(world.landmarks):
        l.state.p_pos = np.array(landmark_pos[i], dtype=np.float64)
        l.state.p_vel = np.zeros(world.dim_p)


class BatchedOracleEnv:
    """E independent oracle envs stepped one after the other (the CPU picture of what the GPU
    kernel does in lockstep).  Arrays are (E, ...) float64 state / float32 actions."""

    def __init__(self, scenario_name, num_envs, num_agents=None, seed=0):
        self.rng = np.random.RandomState(seed)
        self.envs = [make_env(scenario_name, self.rng, num_agents) for _ in range(num_envs)]
        e0 = self.envs[0]
        self.n = e0.n
        self.num_envs = num_envs
        self.obs_dims = [s.shape[0] for s in e0.observation_space]
        self.act_dims = [int(np.sum(s.high - s.low + 1)) if isinstance(s, MultiDiscrete) else s.n
                         for s in e0.action_space]
        self.action_space = e0.action_space
        self.observation_space = e0.observation_space

    def get_state(self):
        st = [get_world_state(e.world) for e in self.envs]
        return {k: np.stack([s[k] for s in st]) for k in st[0]}

    def set_state(self, agent_pos, agent_vel, landmark_pos, agent_c=None, goal=None):
        for e, env in enumerate(self.envs):
            set_world_state(env.world, agent_pos[e], agent_vel[e], landmark_pos[e],
                            None if agent_c is None else agent_c[e], None if goal is None else goal[e], env.scenario)

    def observe(self):
        return [np.stack([env._get_obs(env.agents[i]) for env in self.envs]) for i in range(self.n)]

    def reset(self):
        outs = [env.reset() for env in self.envs]
        return [np.stack([o[i] for o in outs]) for i in range(self.n)]

    def step(self, action_n):
        """action_n: list over agents of (E, K_i) float32.  Returns obs_n [(E,D_i) f64],
        rew (E, A) f64, done (E, A) bool."""
        obs = [[] for _ in range(self.n)]
        rew = np.zeros((self.num_envs, self.n))
        done = np.zeros((self.num_envs, self.n), dtype=bool)
        for e, env in enumerate(self.envs):
            o, r, d, _ = env.step([np.asarray(action_n[i][e], dtype=np.float32) for i in range(self.n)])
            for i in range(self.n):
                obs[i].append(o[i])
            rew[e] = r
            done[e] = d
        return [np.stack(o) for o in obs], rew, done
