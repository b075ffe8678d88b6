"""ctypes binding of libcleanba_b200.so (include/cleanba_b200.h).  No fallback: if the CUDA library is missing or a
call fails, a CleanbaError is raised -- the product path never degrades to a CPU implementation."""
import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_int32, c_longlong, c_uint8, c_uint32, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
# CLEANBA_B200_LIB: another build of the SAME library (developer A/B runs of two builds on one box); default = the in-tree build
LIB_PATH = os.environ.get("CLEANBA_B200_LIB") or os.path.join(_HERE, "libcleanba_b200.so")

CB_ALGO_PPO, CB_ALGO_IMPALA = 0, 1
CB_CONV_TCGEN05, CB_CONV_SIMT = 0, 1
CB_MODEL_IMPALA_RESNET, CB_MODEL_NATURE_CNN = 0, 1
MODELS = {"impala_resnet": CB_MODEL_IMPALA_RESNET, "nature_cnn": CB_MODEL_NATURE_CNN}


class CleanbaError(RuntimeError):
    pass


class cb_config(ctypes.Structure):
    _fields_ = [("device", c_int), ("algo", c_int), ("max_batch", c_int), ("train", c_int),
                ("num_actions", c_int), ("conv_backend", c_int), ("model", c_int)]


# every symbol declared in include/cleanba_b200.h: (restype, argtypes)
_P = c_void_p
SIGNATURES = {
    "cb_last_error": (c_char_p, []),
    "cb_version": (c_int, []),
    "cb_create": (c_int, [POINTER(cb_config), POINTER(_P)]),
    "cb_destroy": (None, [_P]),
    "cb_num_params": (c_longlong, [c_int]),
    "cb_num_leaves": (c_int, []),
    "cb_leaf_info": (c_int, [c_int, c_int, c_char_p, c_int, POINTER(c_longlong), POINTER(c_int), POINTER(c_int)]),
    "cb_num_params_model": (c_longlong, [c_int, c_int]),
    "cb_num_leaves_model": (c_int, [c_int]),
    "cb_leaf_info_model": (c_int, [c_int, c_int, c_int, c_char_p, c_int, POINTER(c_longlong), POINTER(c_int), POINTER(c_int)]),
    "cb_hidden_width": (c_int, [_P]),
    "cb_set_params": (c_int, [_P, _P, _P]),
    "cb_get_params": (c_int, [_P, _P, _P]),
    "cb_params_ptr": (_P, [_P]),
    "cb_refresh_weights": (c_int, [_P, _P]),
    "cb_publish_params": (c_int, [_P, _P, _P]),
    "cb_get_opt_state": (c_int, [_P, _P, _P, POINTER(c_longlong), _P]),
    "cb_set_opt_state": (c_int, [_P, _P, _P, c_longlong, _P]),
    "cb_actor_step": (c_int, [_P, _P, c_int, _P, _P, _P, _P, _P, _P]),
    "cb_actor_step_cursor": (c_int, [_P, _P, c_int, _P, _P]),
    "cb_policy_value": (c_int, [_P, _P, _P, c_int, _P, _P, _P]),
    "cb_gae": (c_int, [_P, _P, _P, _P, _P, _P, c_int, c_int, c_float, c_float, c_int, _P, _P, _P]),
    "cb_split_key": (c_int, [_P, _P, _P, _P]),
    "cb_permutation": (c_int, [_P, _P, c_int, _P, _P]),
    "cb_ppo_grad": (c_int, [_P, _P, _P, c_int, _P, _P, _P, _P, c_float, c_float, c_float, _P, _P, _P]),
    "cb_impala_grad": (c_int, [_P, _P, _P, c_int, c_int, _P, _P, _P, _P, _P, c_float, c_float, c_float, _P, _P, _P]),
    "cb_optimizer_step": (c_int, [_P, _P, c_float, c_float, c_float, _P, _P]),
    "cb_grad_accumulate": (c_int, [_P, _P, _P, c_int, _P]),
    "cb_optimizer_step_peers": (c_int, [_P, POINTER(_P), c_int, c_float, c_float, c_float, _P, _P]),
    "cb_enable_peer_access": (c_int, [_P, c_int]),
    "cb_set_grad_milestone": (c_int, [_P, _P, POINTER(c_longlong)]),
    "cb_graph_steps": (c_int, [_P, c_int]),
    "cb_set_actor_tail": (c_int, [_P, c_int]),
    "cb_graph_replays": (c_longlong, [_P]),
    "cb_reduce_peers": (c_int, [_P, POINTER(_P), c_int, _P, _P]),
    "cb_memcpy_2d": (c_int, [_P, ctypes.c_size_t, _P, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_size_t, _P]),
    "cb_set_sm_budget": (c_int, [_P, c_int]),
    "cb_launch_count": (c_longlong, []),
    "cb_profile": (c_int, [_P, c_int]),
    "cb_profile_report": (c_int, [_P, c_char_p, c_int]),
    "cb_debug_tensor": (c_longlong, [_P, c_char_p, _P, c_longlong]),
}

_lib = None


def load():
    """Load the shared library (once).  Raises CleanbaError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CleanbaError(f"{LIB_PATH} not found: build it with `python -m cleanba_b200.build` "
                           "(nvcc, sm_100a). There is no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here means the .so does not match the header
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise CleanbaError(load().cb_last_error().decode("utf-8", "replace"))


def leaves(num_actions=18, model=CB_MODEL_IMPALA_RESNET):
    """[(flax path, offset, shape)] of the flat parameter vector of the given trunk."""
    lib = load()
    out = []
    for i in range(lib.cb_num_leaves_model(model)):
        name = ctypes.create_string_buffer(256)
        off, nd = c_longlong(), c_int()
        shape = (c_int * 4)()
        check(lib.cb_leaf_info_model(model, i, num_actions, name, 256, ctypes.byref(off), ctypes.byref(nd), shape))
        out.append((name.value.decode(), off.value, tuple(shape[: nd.value])))
    return out
