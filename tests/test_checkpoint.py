"""`.cleanrl_model` interop (cleanba/cleanba_ppo.py:753-771): the flax msgpack wire format restated without flax.
The byte-level expectations below are written out from the format (flax/serialization.py of flax 0.6.8: state dict with
"0","1",... keys for lists; ExtType 1 = packb((shape, dtype.name, bytes))), not produced by this module."""
import struct

import msgpack
import numpy as np
import pytest

from cleanba_b200 import checkpoint as ck


def test_ndarray_wire_bytes_match_the_flax_format():
    arr = np.arange(6, dtype=np.float32).reshape(2, 3)
    got = ck.msgpack_serialize({"a": arr})
    # payload: fixarray(3) [ fixarray(2) [2, 3], fixstr "float32", bin8(24) raw ]
    payload = b"\x93" + b"\x92\x02\x03" + b"\xa7float32" + b"\xc4\x18" + arr.tobytes()
    assert len(payload) == 38
    # top: fixmap(1) { fixstr "a": ext8(len 38, type 1) payload }
    expect = b"\x81" + b"\xa1a" + b"\xc7" + bytes([len(payload)]) + b"\x01" + payload
    assert got == expect
    back = ck.msgpack_restore(got)
    assert back["a"].dtype == np.float32 and back["a"].shape == (2, 3) and np.array_equal(back["a"], arr)


def test_scalars_lists_and_numpy_scalars():
    state = ck.to_state_dict([{"seed": 1, "lr": 2.5e-4, "env_id": "Breakout-v5", "ids": [0, 1], "none": None, "flag": True},
                              [np.float32(1.5)]])
    assert set(state.keys()) == {"0", "1"} and state["0"]["ids"] == {"0": 0, "1": 1}
    data = ck.msgpack_serialize(state)
    # the numpy scalar travels as ExtType 3 with the ndarray payload of a 0-d array
    raw = msgpack.unpackb(data, raw=False, strict_map_key=False, ext_hook=lambda c, d: ("ext", c, d))
    tag, code, payload = raw["1"]["0"]
    assert (tag, code) == ("ext", 3)
    shape, dtype, buf = msgpack.unpackb(payload, raw=False)
    assert shape == [] and dtype == "float32" and struct.unpack("<f", buf)[0] == 1.5
    back = ck.msgpack_restore(data)
    assert back["0"]["env_id"] == "Breakout-v5" and back["0"]["none"] is None and back["0"]["flag"] is True
    assert back["1"]["0"] == np.float32(1.5)


def test_flat_tree_roundtrip_and_flax_names():
    n = 1094115
    flat = np.random.default_rng(0).standard_normal(n).astype(np.float32)
    tree = ck.flat_to_tree(flat)
    assert set(tree) == {"network_params", "actor_params", "critic_params"}
    net = tree["network_params"]["params"]
    assert set(net) == {"ConvSequence_0", "ConvSequence_1", "ConvSequence_2", "Dense_0"}
    assert net["ConvSequence_0"]["Conv_0"]["kernel"].shape == (3, 3, 4, 16)          # HWIO (flax nn.Conv)
    assert net["ConvSequence_1"]["ResidualBlock_1"]["Conv_1"]["kernel"].shape == (3, 3, 32, 32)
    assert net["Dense_0"]["kernel"].shape == (3872, 256)
    assert tree["actor_params"]["params"]["Dense_0"]["kernel"].shape == (256, 18)
    assert tree["critic_params"]["params"]["Dense_0"]["bias"].shape == (1,)
    assert np.array_equal(ck.tree_to_flat(tree), flat)
    with pytest.raises(ValueError):
        ck.flat_to_tree(flat[:-1])
    tree["actor_params"]["params"]["Dense_0"]["kernel"] = np.zeros((256, 6), np.float32)
    with pytest.raises(ValueError):
        ck.tree_to_flat(tree)


def test_cleanrl_model_file_roundtrip(tmp_path):
    from cleanba_b200.params import init_params
    from cleanba_b200.sebulba import Args
    flat = init_params(3)
    args = Args(seed=3, actor_device_ids=[0], learner_device_ids=[1, 2])
    path = ck.save_cleanrl_model(str(tmp_path / "runs" / "x" / "cleanba_ppo.cleanrl_model"), args, flat)
    saved_args, back = ck.load_cleanrl_model(path)
    assert np.array_equal(back, flat)
    assert saved_args["seed"] == 3 and saved_args["learner_device_ids"] == {"0": 1, "1": 2} and saved_args["env_id"] == "Breakout-v5"
    # the top level is the state dict of [args, [network, actor, critic]]
    raw = ck.msgpack_restore(open(path, "rb").read())
    assert set(raw) == {"0", "1"} and set(raw["1"]) == {"0", "1", "2"}
    assert set(raw["1"]["1"]["params"]["Dense_0"]) == {"bias", "kernel"}
    with open(tmp_path / "bad.cleanrl_model", "wb") as f:
        f.write(ck.msgpack_serialize({"x": 1}))
    with pytest.raises(ValueError):
        ck.load_cleanrl_model(str(tmp_path / "bad.cleanrl_model"))


def test_train_state_sidecar(tmp_path):
    p = np.arange(8, dtype=np.float32)
    path = ck.save_train_state(str(tmp_path / "s.npz"), p, p * 2, p * 3, 7, np.array([1, 2], np.uint32), 5, 1234)
    st = ck.load_train_state(path)
    assert st["count"] == 7 and st["learner_policy_version"] == 5 and st["global_step"] == 1234
    assert np.array_equal(st["m"], p * 2) and np.array_equal(st["key"], np.array([1, 2], np.uint32))
