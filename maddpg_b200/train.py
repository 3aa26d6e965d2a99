"""experiments/train.py on the B200 kernels: same flags, same loop (reference train.py:78-189), with the
environment and the trainers replaced by the device-backed drop-ins of this package.

    python -m maddpg_b200.train --scenario simple_spread --num-episodes 1000            # reference shape (1 env)
    python -m maddpg_b200.train --scenario simple_spread --num-envs 4096 --num-episodes 40960   # batched

With ``--num-envs 1`` every call below has the reference's shapes (numpy in / numpy out).  With
``--num-envs E > 1`` the same loop runs on CUDA tensors with a leading env axis: one loop iteration is E
transitions, ``terminal`` is shared by all env instances, and rewards are summed over env instances for the
episode statistics.
"""
import argparse
import os
import pickle
import time

import numpy as np
import torch

from .env import make_env as _make_env
from .trainer import MADDPGAgentTrainer


def parse_args(argv=None):
    # experiments/train.py:11-37 (same names and defaults) + device-side knobs
    parser = argparse.ArgumentParser("Reinforcement Learning experiments for multiagent environments")
    parser.add_argument("--scenario", type=str, default="simple", help="name of the scenario script")
    parser.add_argument("--max-episode-len", type=int, default=25, help="maximum episode length")
    parser.add_argument("--num-episodes", type=int, default=60000, help="number of episodes")
    parser.add_argument("--num-adversaries", type=int, default=0, help="number of adversaries")
    parser.add_argument("--good-policy", type=str, default="maddpg", help="policy for good agents")
    parser.add_argument("--adv-policy", type=str, default="maddpg", help="policy of adversaries")
    parser.add_argument("--lr", type=float, default=1e-2, help="learning rate for Adam optimizer")
    parser.add_argument("--gamma", type=float, default=0.95, help="discount factor")
    parser.add_argument("--batch-size", type=int, default=1024, help="number of episodes to optimize at the same time")
    parser.add_argument("--num-units", type=int, default=64, help="number of units in the mlp")
    parser.add_argument("--exp-name", type=str, default=None, help="name of the experiment")
    parser.add_argument("--save-dir", type=str, default="/tmp/policy/", help="directory in which training state and model should be saved")
    parser.add_argument("--save-rate", type=int, default=1000, help="save model once every time this many episodes are completed")
    parser.add_argument("--load-dir", type=str, default="", help="directory in which training state and model are loaded")
    parser.add_argument("--restore", action="store_true", default=False)
    parser.add_argument("--display", action="store_true", default=False)
    parser.add_argument("--benchmark", action="store_true", default=False)
    parser.add_argument("--benchmark-iters", type=int, default=100000, help="number of iterations run for benchmarking")
    parser.add_argument("--benchmark-dir", type=str, default="./benchmark_files/", help="directory where benchmark data is saved")
    parser.add_argument("--plots-dir", type=str, default="./learning_curves/", help="directory where plot data is saved")
    # new (not in the reference)
    parser.add_argument("--num-envs", type=int, default=1, help="lockstep env instances on the GPU")
    parser.add_argument("--num-agents", type=int, default=None, help="simple_spread only: N agents = N landmarks")
    parser.add_argument("--device", type=str, default="cuda")
    parser.add_argument("--seed", type=int, default=0)
    parser.add_argument("--replay-capacity", type=int, default=int(1e6))
    return parser.parse_args(argv)


def mlp_model(input, num_outputs, scope, reuse=False, num_units=64, rnn_cell=None):
    """Signature placeholder of train.py:39-46: the 3-layer ReLU MLP lives in the CUDA kernels
    (csrc/mdp_mlp.cuh); trainers accept this callable for signature parity and never call it."""
    raise NotImplementedError("the MLP is evaluated by libmaddpg_b200; this callable only marks the architecture")


def make_env(scenario_name, arglist, benchmark=False):
    # train.py:48-61
    return _make_env(scenario_name, arglist, benchmark, num_agents=getattr(arglist, "num_agents", None),
                     device=getattr(arglist, "device", "cuda"))


def get_trainers(env, num_adversaries, obs_shape_n, arglist):
    # train.py:63-75 (adversaries first; local_q_func == ddpg)
    trainers = []
    model = mlp_model
    trainer = MADDPGAgentTrainer
    for i in range(num_adversaries):
        trainers.append(trainer("agent_%d" % i, model, obs_shape_n, env.action_space, i, arglist,
                                local_q_func=(arglist.adv_policy == "ddpg")))
    for i in range(num_adversaries, env.n):
        trainers.append(trainer("agent_%d" % i, model, obs_shape_n, env.action_space, i, arglist,
                                local_q_func=(arglist.good_policy == "ddpg")))
    return trainers


def save_state(save_dir, trainers):
    """U.save_state (tf_util.py:267-273): all variables incl. Adam slots (the replay ring is not saved,
    like the reference)."""
    core = trainers[0].core
    os.makedirs(save_dir, exist_ok=True)
    torch.save({"params": core.params.cpu(), "adam_m": core.adam_m.cpu(), "adam_v": core.adam_v.cpu(),
                "adam_t": core.adam_t.cpu(), "obs_dims": core.obs_dims, "act_dims": core.act_dims,
                "num_units": core.num_units}, os.path.join(save_dir, "maddpg_b200.pt"))


def load_state(load_dir, trainers):
    """U.load_state (tf_util.py:259-265)."""
    core = trainers[0].core
    st = torch.load(os.path.join(load_dir, "maddpg_b200.pt"), map_location="cpu")
    assert st["obs_dims"] == core.obs_dims and st["act_dims"] == core.act_dims and st["num_units"] == core.num_units
    core.params.copy_(st["params"])
    core.adam_m.copy_(st["adam_m"])
    core.adam_v.copy_(st["adam_v"])
    core.adam_t.copy_(st["adam_t"])


def _rew_scalar(r):
    return float(r.sum().item()) if isinstance(r, torch.Tensor) else float(np.sum(r))


def train(arglist):
    # train.py:78-189
    env = make_env(arglist.scenario, arglist, arglist.benchmark)
    obs_shape_n = [env.observation_space[i].shape for i in range(env.n)]
    num_adversaries = min(env.n, arglist.num_adversaries)
    trainers = get_trainers(env, num_adversaries, obs_shape_n, arglist)
    print("Using good policy {} and adv policy {}".format(arglist.good_policy, arglist.adv_policy))
    if arglist.load_dir == "":
        arglist.load_dir = arglist.save_dir
    if arglist.display or arglist.restore or arglist.benchmark:
        print("Loading previous state...")
        load_state(arglist.load_dir, trainers)

    episode_rewards = [0.0]
    agent_rewards = [[0.0] for _ in range(env.n)]
    final_ep_rewards = []
    final_ep_ag_rewards = []
    agent_info = [[[]]]
    obs_n = env.reset()
    episode_step = 0
    train_step = 0
    t_start = time.time()

    print("Starting iterations...")
    while True:
        action_n = [agent.action(obs) for agent, obs in zip(trainers, obs_n)]
        new_obs_n, rew_n, done_n, info_n = env.step(action_n)
        episode_step += 1
        done = all(bool(np.all(d.cpu().numpy())) if isinstance(d, torch.Tensor) else bool(np.all(d)) for d in done_n)
        terminal = (episode_step >= arglist.max_episode_len)
        for i, agent in enumerate(trainers):
            agent.experience(obs_n[i], action_n[i], rew_n[i], new_obs_n[i], done_n[i], terminal)
        obs_n = new_obs_n

        for i, rew in enumerate(rew_n):
            r = _rew_scalar(rew)
            episode_rewards[-1] += r
            agent_rewards[i][-1] += r

        if done or terminal:
            obs_n = env.reset()
            episode_step = 0
            episode_rewards.append(0)
            for a in agent_rewards:
                a.append(0)
            agent_info.append([[]])

        train_step += 1

        if arglist.benchmark:
            for i, info in enumerate(info_n):
                agent_info[-1][i].append(info_n["n"])
            if train_step > arglist.benchmark_iters and (done or terminal):
                file_name = arglist.benchmark_dir + arglist.exp_name + ".pkl"
                print("Finished benchmarking, now saving...")
                with open(file_name, "wb") as fp:
                    pickle.dump(agent_info[:-1], fp)
                break
            continue

        if arglist.display:
            time.sleep(0.1)
            env.render()
            continue

        loss = None
        for agent in trainers:
            agent.preupdate()
        for agent in trainers:
            loss = agent.update(trainers, train_step)

        if terminal and (len(episode_rewards) % arglist.save_rate == 0):
            save_state(arglist.save_dir, trainers)
            if num_adversaries == 0:
                print("steps: {}, episodes: {}, mean episode reward: {}, time: {}".format(
                    train_step, len(episode_rewards), np.mean(episode_rewards[-arglist.save_rate:]),
                    round(time.time() - t_start, 3)))
            else:
                print("steps: {}, episodes: {}, mean episode reward: {}, agent episode reward: {}, time: {}".format(
                    train_step, len(episode_rewards), np.mean(episode_rewards[-arglist.save_rate:]),
                    [np.mean(rew[-arglist.save_rate:]) for rew in agent_rewards], round(time.time() - t_start, 3)))
            t_start = time.time()
            final_ep_rewards.append(np.mean(episode_rewards[-arglist.save_rate:]))
            for rew in agent_rewards:
                final_ep_ag_rewards.append(np.mean(rew[-arglist.save_rate:]))

        if len(episode_rewards) > arglist.num_episodes:
            os.makedirs(arglist.plots_dir, exist_ok=True)
            rew_file_name = arglist.plots_dir + str(arglist.exp_name) + "_rewards.pkl"
            with open(rew_file_name, "wb") as fp:
                pickle.dump(final_ep_rewards, fp)
            agrew_file_name = arglist.plots_dir + str(arglist.exp_name) + "_agrewards.pkl"
            with open(agrew_file_name, "wb") as fp:
                pickle.dump(final_ep_ag_rewards, fp)
            print("...Finished total of {} episodes.".format(len(episode_rewards)))
            break
    return trainers, episode_rewards


if __name__ == "__main__":
    train(parse_args())
