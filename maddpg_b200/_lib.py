"""ctypes binding of libmaddpg_b200.so (the C ABI declared in include/maddpg_b200.h).

The library is built in-tree by ``maddpg_b200/csrc/build.sh`` (``__graft_entry__.build()``) into
``maddpg_b200/_lib/``.  There is NO fallback: if the shared object is missing or a symbol cannot be
resolved the import raises, and every wrapper raises on a non-zero status.
"""
import ctypes as C
import os

MAX_AGENTS = 32
MAX_HEADS = 2

MDP_OK, MDP_EINVAL, MDP_ECUDA, MDP_ENOTSUP = 0, -1, -2, -3
SCENARIO_IDS = {"simple": 0, "simple_spread": 1, "simple_tag": 2, "simple_world_comm": 3,
                "simple_adversary": 4, "simple_push": 5, "simple_speaker_listener": 6, "simple_crypto": 7,
                "simple_reference": 8}
NET_P, NET_TARGET_P, NET_Q, NET_TARGET_Q = 0, 1, 2, 3

_I32A = C.c_int32 * MAX_AGENTS
_I32AH = (C.c_int32 * MAX_HEADS) * MAX_AGENTS
_I64A = C.c_int64 * MAX_AGENTS
_I64A4 = (C.c_int64 * 4) * MAX_AGENTS
_I64A2 = (C.c_int64 * 2) * MAX_AGENTS
_I32A4 = (C.c_int32 * 4) * MAX_AGENTS


class EnvCfg(C.Structure):
    _fields_ = [("scenario", C.c_int32), ("num_agents", C.c_int32), ("state_f64", C.c_int32)]


class EnvDims(C.Structure):
    _fields_ = [("n_agents", C.c_int32), ("n_landmarks", C.c_int32), ("comm_dim", C.c_int32),
                ("collaborative", C.c_int32),
                ("obs_dim", _I32A), ("act_dim", _I32A), ("obs_off", _I32A), ("act_off", _I32A),
                ("n_heads", _I32A), ("head_dim", _I32AH),
                ("obs_sum", C.c_int32), ("act_sum", C.c_int32), ("obs_stride", C.c_int32), ("act_stride", C.c_int32),
                ("state_comps", C.c_int32), ("state_elem_size", C.c_int32), ("env_bytes_per_step", C.c_int32),
                ("n_goal", C.c_int32), ("comm_off", _I32A), ("comm_len", _I32A), ("movable", _I32A)]


class RingLayout(C.Structure):
    _fields_ = [("n_agents", C.c_int32),
                ("obs_dim", _I32A), ("act_dim", _I32A), ("obs_off", _I32A), ("act_off", _I32A),
                ("obs_sum", C.c_int32), ("act_sum", C.c_int32), ("x_dim", C.c_int32),
                ("nx_off", C.c_int32), ("rw_off", C.c_int32), ("dn_off", C.c_int32), ("row_stride", C.c_int32)]


class CoreCfg(C.Structure):
    _fields_ = [("n_agents", C.c_int32), ("num_units", C.c_int32),
                ("obs_dim", _I32A), ("act_dim", _I32A), ("n_heads", _I32A), ("head_dim", _I32AH),
                ("local_q", _I32A),
                ("lr", C.c_double), ("gamma", C.c_double), ("polyak", C.c_double), ("grad_clip", C.c_double),
                ("actor_reg", C.c_double), ("beta1", C.c_double), ("beta2", C.c_double), ("adam_eps", C.c_double)]


class CoreLayout(C.Structure):
    _fields_ = [("total_params", C.c_int64), ("total_train", C.c_int64),
                ("net_off", _I64A4), ("train_off", _I64A2),
                ("net_in", _I32A4), ("net_out", _I32A4), ("net_size", _I64A4),
                ("update_flops_critic", _I64A), ("update_flops_actor", _I64A)]


# every symbol include/maddpg_b200.h declares: name -> (restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    "mdp_env_create": (C.c_int, [C.POINTER(EnvCfg), C.POINTER(_P)]),
    "mdp_env_get_dims": (C.c_int, [_P, C.POINTER(EnvDims)]),
    "mdp_env_destroy": (None, [_P]),
    "mdp_env_force_generic": (C.c_int, [_P, C.c_int32]),
    "mdp_env_reset": (C.c_int, [_P, C.c_int32, _P, _P, C.c_uint64, C.c_uint64, _P, _P]),
    "mdp_env_step": (C.c_int, [_P, C.c_int32, _P, _P, _P, _P, _P, _P, _P, C.c_int64, C.c_int32, C.c_int64, _P]),
    "mdp_ring_make_layout": (C.c_int, [C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(RingLayout)]),
    "mdp_replay_insert": (C.c_int, [C.POINTER(RingLayout), _P, C.c_int64, C.c_int64, C.c_int32, C.c_int32,
                                    _P, C.c_int32, _P, C.c_int32, _P, C.c_int32, _P, C.c_int32, _P, C.c_int32, _P]),
    "mdp_replay_gather": (C.c_int, [_P, C.c_int64, C.c_int32, _P, C.c_int32, _P, C.c_int32, _P]),
    "mdp_core_create": (C.c_int, [C.POINTER(CoreCfg), C.POINTER(_P)]),
    "mdp_core_get_layout": (C.c_int, [_P, C.POINTER(CoreLayout)]),
    "mdp_core_destroy": (None, [_P]),
    "mdp_core_set_tensor_cores": (C.c_int, [_P, C.c_int32]),
    "mdp_core_set_fused_update": (C.c_int, [_P, C.c_int32]),
    "mdp_core_bind_peers": (C.c_int, [_P, C.c_int32, C.c_int32, _P, _P, _P, _P]),
    "mdp_core_bind": (C.c_int, [_P, _P, _P, _P, _P, _P, _P]),
    "mdp_actor_act": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P, C.c_int32, _P, C.c_int32, _P,
                                C.c_uint64, C.c_uint64, _P, _P]),
    "mdp_critic_q": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_int32, _P, C.c_int32, _P, _P]),
    "mdp_td_target": (C.c_int, [_P, C.c_int32, C.POINTER(RingLayout), C.c_int32, _P, _P, _P, C.c_int32, C.c_uint64,
                                C.c_uint64, _P, _P, _P]),
    "mdp_td_target_all": (C.c_int, [_P, C.POINTER(RingLayout), C.c_int32, _P, _P, C.c_int64, C.c_uint64, C.c_uint64, _P, _P]),
    "mdp_critic_grads": (C.c_int, [_P, C.c_int32, C.POINTER(RingLayout), C.c_int32, _P, _P, _P, _P, _P]),
    "mdp_critic_grads_all": (C.c_int, [_P, C.POINTER(RingLayout), C.c_int32, _P, _P, C.c_int64, _P, _P]),
    "mdp_actor_grads": (C.c_int, [_P, C.c_int32, C.POINTER(RingLayout), C.c_int32, _P, _P, _P, C.c_int32, C.c_uint64,
                                  C.c_uint64, _P]),
    "mdp_clip_adam_polyak": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_float, C.c_int32, _P]),
    "mdp_update_agent": (C.c_int, [_P, C.c_int32, C.POINTER(RingLayout), C.c_int32, _P, _P, _P, _P, C.c_int32,
                                   C.c_uint64, C.c_uint64, _P, _P]),
    "mdp_update_prepare": (C.c_int, [_P, C.c_int32, C.c_int32, _P, C.c_int32, C.c_int64, C.c_uint64, C.c_uint64, _P]),
    "mdp_update_all": (C.c_int, [_P, C.POINTER(RingLayout), C.c_int32, _P, _P, C.c_int64, C.c_uint64, C.c_uint64, _P,
                                 C.c_float, _P]),
    "mdp_clip_adam_polyak_all": (C.c_int, [_P, C.c_int32, C.c_float, C.c_int32, _P]),
    "mdp_rollout_episode": (C.c_int, [_P, _P, C.c_int32, _P, _P, _P, C.c_int64, C.c_int32, C.c_int64, C.c_int32,
                                      C.c_uint64, C.c_uint64, C.c_int32, C.c_uint64, C.c_uint64, _P, _P]),
    "mdp_philox_uniform": (C.c_int, [C.c_uint64, C.c_uint64, C.c_uint32, C.c_int64, C.c_int32, C.c_int32, _P]),
    "mdp_rollout_episodes": (C.c_int, [_P, _P, C.c_int32, _P, _P, _P, C.c_int64, C.c_int32, C.c_int64, C.c_int32, C.c_int32,
                                       C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, _P, _P]),
    "mdp_host_step_layout": (C.c_int, [_P, C.c_int32, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "mdp_host_step": (C.c_int, [_P, _P, C.c_int32, _P, _P, _P, _P, _P, _P, C.c_int64, C.c_int32, C.c_int64,
                                C.c_uint64, C.c_uint64, _P]),
    "mdp_host_step_pipelined": (C.c_int, [_P, _P, C.c_int32, C.c_int32, _P, _P, _P, _P, _P, _P, C.c_int64, C.c_int32,
                                          C.c_int64, C.c_uint64, C.c_uint64, _P]),
    "mdp_host_copy_mode": (C.c_int, [_P, C.c_int32]),
    "mdp_env_benchmark": (C.c_int, [_P, C.c_int32, _P, _P, _P]),
    "mdp_env_set_ctl": (C.c_int, [_P, _P]),
    "mdp_core_set_ctl": (C.c_int, [_P, _P]),
    "mdp_ctl_advance": (C.c_int, [_P, C.c_uint64, C.c_int64, C.c_int64, C.c_uint64, _P]),
    "mdp_replay_make_index": (C.c_int, [_P, C.c_int32, C.c_int64, C.c_uint64, C.c_uint64, _P, _P]),
    "mdp_sumtree_layout": (C.c_int, [C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_int32), C.POINTER(C.c_int64)]),
    "mdp_sumtree_flush": (C.c_int, [_P, C.c_int64, C.c_int64, C.c_int64, C.c_double, _P, _P]),
    "mdp_sumtree_sample": (C.c_int, [_P, C.c_int64, C.c_int64, C.c_int64, C.c_double, C.c_int32, _P, C.c_double, _P, _P, _P,
                                     _P, _P, _P]),
    "mdp_sumtree_update": (C.c_int, [_P, C.c_int64, _P, C.c_int32, _P, _P, C.c_double, C.c_double, C.c_double, _P, _P, _P]),
    "mdp_td3_policy_act": (C.c_int, [_P, C.c_int32, C.c_int32, _P, C.c_int32, _P, C.c_int32, C.c_float, C.c_float, C.c_uint64,
                                     C.c_uint64, _P, _P, _P, C.c_int32, C.c_int32, _P]),
    "mdp_td3_q_target": (C.c_int, [_P, _P, C.c_int32, C.POINTER(RingLayout), C.c_int32, _P, C.c_int32, _P, C.c_int32, _P, _P,
                                   C.c_int32, C.c_float, _P, _P, _P]),
    "mdp_td3_policy_grads": (C.c_int, [_P, _P, C.c_int32, C.c_float, C.POINTER(RingLayout), C.c_int32, _P, _P, C.c_int32, _P, _P,
                                       C.c_int32, C.c_int32, _P]),
    "mdp_td3_polyak": (C.c_int, [_P, C.c_int32, C.c_double, _P]),
    "mdp_stream_synchronize": (C.c_int, [_P]),
    "mdp_last_error": (C.c_char_p, []),
    "mdp_version": (C.c_char_p, []),
    "mdp_launch_count": (C.c_int64, []),
}

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_lib", os.environ.get("MDP_LIB_NAME", "libmaddpg_b200.so"))


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "maddpg_b200: %s is missing -- build it with `bash maddpg_b200/csrc/build.sh` "
            "(or __graft_entry__.build()); there is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


class MdpError(RuntimeError):
    pass


def check(rc, what=""):
    if rc != 0:
        msg = lib.mdp_last_error().decode("utf-8", "replace")
        exc = ValueError if rc == MDP_EINVAL else NotImplementedError if rc == MDP_ENOTSUP else MdpError
        raise exc("%s failed (%d): %s" % (what or "libmaddpg_b200 call", rc, msg))


def launch_count():
    return int(lib.mdp_launch_count())


def ptr(t):
    """Device (or host) address of a torch tensor / None as a c_void_p."""
    if t is None:
        return None
    return C.c_void_p(t.data_ptr())


_raw_stream = None


def current_stream():
    """The raw cudaStream_t of torch's current stream on the current device as a c_void_p.  ``torch.cuda.current_stream()``
    builds a Stream object through several layers of device-index resolution (~14 us per call, measured under cProfile in the
    reference's loop: a third of an iteration of experiments/train.py on these kernels); the two C entry points it ends in
    cost well under a microsecond."""
    global _raw_stream
    if _raw_stream is None:
        import torch
        torch.cuda.current_stream()  # initialises the CUDA context
        _raw_stream = (torch._C._cuda_getCurrentRawStream, torch._C._cuda_getDevice)
    return C.c_void_p(_raw_stream[0](_raw_stream[1]()))


def synchronize_current_stream():
    """cudaStreamSynchronize on torch's current stream without building a Stream object."""
    check(lib.mdp_stream_synchronize(current_stream()), "mdp_stream_synchronize")
