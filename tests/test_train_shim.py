"""maddpg_b200.train's stand-in modules satisfy every import of the reference's experiments/train.py and get out of the way
again (no GPU needed: the script is stopped by argparse's --help right after its imports)."""
import io
import sys
from contextlib import redirect_stdout

import pytest


def test_reference_train_imports_resolve_to_the_stand_ins():
    from maddpg_b200 import train as T
    try:
        path = T.reference_train_path()
    except FileNotFoundError:
        pytest.skip("the reference's experiments/train.py is not available on this machine")
    out = io.StringIO()
    with pytest.raises(SystemExit) as e, redirect_stdout(out):
        T.run_reference_train(["--help"], path)
    assert e.value.code == 0
    text = out.getvalue()
    assert "--scenario" in text and "--num-adversaries" in text and "--benchmark-iters" in text  # train.py:11-37, verbatim
    for name in ("tensorflow", "maddpg.common.tf_util", "maddpg.trainer.maddpg", "multiagent.environment"):
        assert name not in sys.modules


def test_stand_ins_expose_what_train_py_calls():
    from maddpg_b200 import train as T
    saved = T.install_stubs()
    try:
        import maddpg.common.tf_util as U
        import multiagent.scenarios as scenarios
        import tensorflow as tf
        from maddpg.trainer.maddpg import MADDPGAgentTrainer
        from multiagent.environment import MultiAgentEnv  # noqa: F401
        import tensorflow.contrib.layers as layers
        assert issubclass(MADDPGAgentTrainer, T.MADDPGAgentTrainer)
        with U.single_threaded_session():
            U.initialize()
        assert tf.train.Saver() is not None and callable(layers.fully_connected)
        sc = scenarios.load("simple_spread.py").Scenario()
        assert sc.make_world().scenario_name == "simple_spread"
        with pytest.raises(NotImplementedError):
            scenarios.load("simple_football.py")
    finally:
        T.remove_stubs(saved)
