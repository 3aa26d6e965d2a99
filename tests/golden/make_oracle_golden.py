"""Generates the committed oracle fixtures (run anywhere; needs only numpy + oracle/):

    python tests/golden/make_oracle_golden.py

* env_<case>.npz    seeded initial state + action tape + the oracle's observations/rewards/final state for the
                    rollouts of tests/helpers.ENV_CASES (the 4 scenarios + simple_spread N=24)
* update_<case>.npz one sequential update round (every agent once) of tests/helpers.TRAINER_CASES: statistics,
                    TD targets, gradients and post-update parameter checksums

The env and trainer oracles are restatements with no reference-side golden vectors (parity unpinned, see the
module headers); these files pin the oracle against itself and give the CUDA tests a box-independent target."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests.helpers import ENV_CASES, TRAINER_CASES, env_case, oracle_update_round, run_oracle_rollout, trainer_case  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
ENV_SEED, UPD_SEED = 21, 6


def main():
    force = "--force" in sys.argv
    for name in ENV_CASES:
        if os.path.exists(os.path.join(HERE, "env_%s.npz" % name)) and not force:
            continue  # committed fixtures stay byte-identical; --force regenerates everything
        case = env_case(name, seed=ENV_SEED)
        ref = run_oracle_rollout(case)
        tape = np.stack([np.concatenate(a, axis=1) for a in case["tape"]])
        np.savez_compressed(os.path.join(HERE, "env_%s.npz" % name), agent_pos=case["agent_pos"], agent_vel=case["agent_vel"],
                            landmark_pos=case["landmark_pos"], tape=tape.astype(np.float32), obs0=ref["obs0"].astype(np.float32),
                            obs=ref["obs"].astype(np.float32), rew=ref["rew"].astype(np.float32),
                            final_pos=ref["final"]["agent_pos"], final_vel=ref["final"]["agent_vel"],
                            goal=np.zeros((case["E"], 0), np.int64) if case["goal"] is None else case["goal"])
    for name in TRAINER_CASES:
        if os.path.exists(os.path.join(HERE, "update_%s.npz" % name)) and not force:
            continue
        ref = oracle_update_round(trainer_case(name, seed=UPD_SEED))
        out = {}
        for j, r in enumerate(ref):
            out["stats_%d" % j] = np.asarray(r["stats"], np.float64)
            out["y_%d" % j] = r["y"]
            for key in ("q", "p", "target_q", "target_p"):
                out["%s_sum_%d" % (key, j)] = np.asarray([float(np.sum(x, dtype=np.float64)) for x in r[key]])
                out["%s_abs_%d" % (key, j)] = np.asarray([float(np.sum(np.abs(x), dtype=np.float64)) for x in r[key]])
            out["qgrad_norm_%d" % j] = np.asarray([float(np.linalg.norm(g)) for g in r["q_grads"]])
            out["pgrad_norm_%d" % j] = np.asarray([float(np.linalg.norm(g)) for g in r["p_grads"]])
        np.savez_compressed(os.path.join(HERE, "update_%s.npz" % name), **out)
    print("wrote", sorted(f for f in os.listdir(HERE) if f.endswith(".npz")))


if __name__ == "__main__":
    main()
