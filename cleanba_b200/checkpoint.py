"""Checkpoint interop with the reference.

`.cleanrl_model` (cleanba/cleanba_ppo.py:753-771, cleanba/cleanba_impala.py:729-747) is
`flax.serialization.to_bytes([vars(args), [network_params, actor_params, critic_params]])`.  flax (0.6.8,
flax/serialization.py) first turns the target into a "state dict" -- lists/tuples become dicts keyed "0", "1", ...,
FrozenDicts become plain dicts -- and then msgpack-packs it with three extension types:

    ExtType(1, packb((shape, dtype.name, raw C-order bytes)))   numpy / jax arrays
    ExtType(2, packb((real, imag)))                             python complex
    ExtType(3, <same payload as 1>)                             numpy scalars

Neither flax nor jax is installable here, so this module restates that wire format on top of `msgpack` alone; the file
written by `save_cleanrl_model` loads with `flax.serialization.from_bytes` in the reference's eval script
(cleanrl_utils/evals/ppo_envpool_jax_eval.py:35-38) and `load_cleanrl_model` reads files the reference wrote.  The
parameter tree uses flax's own names, which are the leaf names of the C ABI (`cb_leaf_info`):
`network_params/params/ConvSequence_i/...`, `actor_params/params/Dense_0/...`, `critic_params/params/Dense_0/...`.

The reference cannot resume (it saves neither the optimizer state nor the step); `save_train_state` /
`load_train_state` add that as a sidecar `.npz` (parameters, optimizer moments and count, PRNG key, policy version).
"""
from __future__ import annotations

import os
from typing import Any, Dict, Mapping, Tuple

import msgpack
import numpy as np

EXT_NDARRAY, EXT_COMPLEX, EXT_NPSCALAR = 1, 2, 3
_TOP = ("network_params", "actor_params", "critic_params")   # order inside the saved list (cleanba_ppo.py:763-767)


# ------------------------------------------------------------------------------------------------ wire format
def _ndarray_payload(arr: np.ndarray) -> bytes:
    if arr.dtype.hasobject:
        raise ValueError("object arrays cannot be serialised")
    return msgpack.packb((tuple(arr.shape), arr.dtype.name, arr.tobytes("C")), use_bin_type=True)


def _ext_pack(x):
    if isinstance(x, np.ndarray):
        return msgpack.ExtType(EXT_NDARRAY, _ndarray_payload(x))
    if isinstance(x, np.generic):
        return msgpack.ExtType(EXT_NPSCALAR, _ndarray_payload(np.asarray(x)))
    if isinstance(x, complex):
        return msgpack.ExtType(EXT_COMPLEX, msgpack.packb((x.real, x.imag)))
    raise TypeError(f"cannot serialise {type(x)}")


def _ext_unpack(code: int, data: bytes):
    if code in (EXT_NDARRAY, EXT_NPSCALAR):
        shape, dtype_name, buf = msgpack.unpackb(data, raw=True)
        dtype = np.dtype(dtype_name.decode() if isinstance(dtype_name, bytes) else dtype_name)
        arr = np.frombuffer(buf, dtype=dtype).reshape(tuple(shape)).copy()
        return arr if code == EXT_NDARRAY else arr[()]
    if code == EXT_COMPLEX:
        re, im = msgpack.unpackb(data)
        return complex(re, im)
    return msgpack.ExtType(code, data)


def to_state_dict(x: Any) -> Any:
    """flax.serialization.to_state_dict for the containers that occur here (dict / list / tuple / leaves)."""
    if isinstance(x, Mapping):
        return {str(k): to_state_dict(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return {str(i): to_state_dict(v) for i, v in enumerate(x)}
    return x


def msgpack_serialize(state: Any) -> bytes:
    return msgpack.packb(state, default=_ext_pack, strict_types=True, use_bin_type=True)


def msgpack_restore(data: bytes) -> Any:
    return msgpack.unpackb(data, ext_hook=_ext_unpack, raw=False, strict_map_key=False)


# ------------------------------------------------------------------------------------------------ parameter tree
def _leaves(num_actions: int, model: int = 0):
    from . import lib
    return lib.leaves(num_actions, model)


def _model_of_size(n: int, num_actions: int) -> int:
    """The two trunks have different parameter counts: the flat vector's length identifies its model."""
    from . import lib
    for m in (lib.CB_MODEL_IMPALA_RESNET, lib.CB_MODEL_NATURE_CNN):
        if int(lib.load().cb_num_params_model(m, num_actions)) == n:
            return m
    raise ValueError(f"no built model has {n} parameters")


def flat_to_tree(flat: np.ndarray, num_actions: int = 18) -> Dict[str, Any]:
    """Flat fp32 vector (C-ABI leaf order) -> {"network_params": {"params": {...}}, "actor_params": ..., "critic_params": ...}."""
    flat = np.asarray(flat, dtype=np.float32).ravel()
    tree: Dict[str, Any] = {}
    end = 0
    for name, offset, shape in _leaves(num_actions, _model_of_size(flat.size, num_actions)):
        size = int(np.prod(shape))
        node = tree
        parts = name.split("/")
        for p in parts[:-1]:
            node = node.setdefault(p, {})
        node[parts[-1]] = flat[offset:offset + size].reshape(shape).copy()
        end = max(end, offset + size)
    if end != flat.size:
        raise ValueError(f"parameter vector has {flat.size} elements, the model has {end}")
    return tree


def tree_to_flat(tree: Mapping[str, Any], num_actions: int = 18) -> np.ndarray:
    from . import lib
    nature = "Conv_0" in tree.get("network_params", {}).get("params", {})       # flax names of the Nature-CNN trunk
    leaves = _leaves(num_actions, lib.CB_MODEL_NATURE_CNN if nature else lib.CB_MODEL_IMPALA_RESNET)
    total = max(off + int(np.prod(shape)) for _, off, shape in leaves)
    flat = np.empty(total, np.float32)
    for name, offset, shape in leaves:
        node: Any = tree
        for p in name.split("/"):
            if p not in node:
                raise KeyError(f"checkpoint has no leaf {name!r}")
            node = node[p]
        arr = np.asarray(node, dtype=np.float32)
        if tuple(arr.shape) != tuple(shape):
            raise ValueError(f"leaf {name!r}: checkpoint shape {tuple(arr.shape)} != model shape {tuple(shape)}")
        flat[offset:offset + arr.size] = arr.ravel()
    return flat


def _plain_args(args: Any) -> Dict[str, Any]:
    d = dict(vars(args)) if not isinstance(args, Mapping) else dict(args)
    out = {}
    for k, v in d.items():
        if k.startswith("_"):      # runtime handles (e.g. the tracer) are not hyper-parameters
            continue
        if isinstance(v, (np.ndarray, np.generic, str, int, float, bool, type(None), list, tuple, dict)):
            out[k] = v
        else:
            out[k] = str(v)
    return out


# ------------------------------------------------------------------------------------------------ .cleanrl_model
def save_cleanrl_model(path: str, args: Any, flat_params: np.ndarray, num_actions: int = 18) -> str:
    """Write `[vars(args), [network_params, actor_params, critic_params]]` exactly as cleanba_ppo.py:756-770 does."""
    tree = flat_to_tree(flat_params, num_actions)
    target = [_plain_args(args), [tree[k] for k in _TOP]]
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, "wb") as f:
        f.write(msgpack_serialize(to_state_dict(target)))
    return path


def load_cleanrl_model(path: str, num_actions: int = 18) -> Tuple[Dict[str, Any], np.ndarray]:
    """-> (args dict, flat fp32 parameter vector in C-ABI order).  Accepts files written by the reference."""
    with open(path, "rb") as f:
        state = msgpack_restore(f.read())
    if not (isinstance(state, dict) and "0" in state and "1" in state):
        raise ValueError(f"{path}: not a cleanrl model ([args, [network, actor, critic]])")
    params = state["1"]
    tree = {k: params[str(i)] for i, k in enumerate(_TOP)}
    return state["0"], tree_to_flat(tree, num_actions)


# ------------------------------------------------------------------------------------------------ resume sidecar
def save_train_state(path: str, params: np.ndarray, m: np.ndarray, v: np.ndarray, count: int, key: np.ndarray,
                     learner_policy_version: int, global_step: int = 0) -> str:
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    tmp = path + ".tmp.npz"
    np.savez(tmp, params=np.asarray(params, np.float32), m=np.asarray(m, np.float32), v=np.asarray(v, np.float32),
             count=np.int64(count), key=np.asarray(key, np.uint32), learner_policy_version=np.int64(learner_policy_version),
             global_step=np.int64(global_step))
    os.replace(tmp, path)
    return path


def load_train_state(path: str) -> Dict[str, Any]:
    with np.load(path) as z:
        return dict(params=z["params"], m=z["m"], v=z["v"], count=int(z["count"]), key=z["key"],
                    learner_policy_version=int(z["learner_policy_version"]), global_step=int(z["global_step"]))
