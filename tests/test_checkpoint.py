"""MXNet .params reader / writer (rangedet_b200/checkpoint.py, mirror of utils/load_model.py:18-39): byte layout
against a hand-assembled file, round trip of the whole model dictionary, arg/aux split, error behaviour."""
import struct

import numpy as np
import pytest

from rangedet_b200 import checkpoint as ck


def test_reads_a_hand_assembled_v2_file(tmp_path):
    a = np.arange(6, dtype=np.float32).reshape(2, 3)
    b = np.array([1, 2, 3], dtype=np.float16)
    raw = struct.pack("<QQQ", 0x112, 0, 2)
    raw += struct.pack("<IiI2q", 0xF993FAC9, 0, 2, 2, 3) + struct.pack("<iii", 2, 0, 0) + a.tobytes()     # gpu(0), f32
    raw += struct.pack("<IiI1q", 0xF993FAC9, 0, 1, 3) + struct.pack("<iii", 1, 0, 2) + b.tobytes()        # cpu, f16
    raw += struct.pack("<Q", 2)
    for n in (b"arg:conv_weight", b"aux:bn_moving_mean"):
        raw += struct.pack("<Q", len(n)) + n
    p = tmp_path / "m-0007.params"
    p.write_bytes(raw)
    d = ck.nd_load(str(p))
    assert list(d) == ["arg:conv_weight", "aux:bn_moving_mean"]
    assert np.array_equal(d["arg:conv_weight"], a) and d["aux:bn_moving_mean"].dtype == np.float16
    arg, aux = ck.load_checkpoint(str(tmp_path / "m"), 7)
    assert list(arg) == ["conv_weight"] and list(aux) == ["bn_moving_mean"]
    assert ck.get_latest_ckpt_epoch(str(tmp_path / "m")) == 7


def test_v3_header_and_errors(tmp_path):
    a = np.ones((4,), np.int32)
    raw = struct.pack("<QQQ", 0x112, 0, 1) + struct.pack("<Iii1q", 0xF993FACA, 0, 1, 4) + struct.pack("<iii", 1, 0, 4) + a.tobytes()
    raw += struct.pack("<Q", 0)
    p = tmp_path / "x.params"
    p.write_bytes(raw)
    out = ck.nd_load(str(p))
    assert isinstance(out, list) and np.array_equal(out[0], a)
    (tmp_path / "bad.params").write_bytes(struct.pack("<QQQ", 0x113, 0, 0))
    with pytest.raises(ValueError):
        ck.nd_load(str(tmp_path / "bad.params"))
    (tmp_path / "sparse.params").write_bytes(struct.pack("<QQQ", 0x112, 0, 1) + struct.pack("<Ii", 0xF993FAC9, 1))
    with pytest.raises(ValueError):
        ck.nd_load(str(tmp_path / "sparse.params"))
    with pytest.raises(FileNotFoundError):
        ck.get_latest_ckpt_epoch(str(tmp_path / "nothing"))


def test_model_round_trip(tmp_path):
    from rangedet_b200.model_params import make_params, num_parameters
    P = make_params(seed=0, device="cpu")
    arg, aux = ck.from_model_params(P)
    assert all(k.endswith(("_moving_mean", "_moving_var")) for k in aux) and len(aux) > 100
    assert sum(v.size for v in arg.values()) == num_parameters(P)
    ck.save_checkpoint(str(tmp_path / "rangedet"), 18, arg, aux)
    arg2, aux2 = ck.load_checkpoint(str(tmp_path / "rangedet"), ck.get_latest_ckpt_epoch(str(tmp_path / "rangedet")))
    Q = ck.to_model_params(arg2, aux2, device="cpu")
    assert sorted(Q) == sorted(P)
    for k in P:
        assert np.array_equal(Q[k].numpy(), P[k].numpy()), k


def test_v3_none_and_scalar_arrays(tmp_path):
    """np_shape (v3) files: NDArray::Save writes the shape and returns for a none array, which v3 marks with
    ndim -1; ndim 0 is a real scalar there (one element follows).  In v2 ndim 0 is the none marker."""
    s = np.array(2.5, np.float32)
    v = np.array([7, 8], np.int64)
    raw = struct.pack("<QQQ", 0x112, 0, 4)
    raw += struct.pack("<Iii", 0xF993FACA, 0, -1)                                                     # v3 none
    raw += struct.pack("<Iii", 0xF993FACA, 0, 0) + struct.pack("<iii", 1, 0, 0) + s.tobytes()         # v3 scalar
    raw += struct.pack("<IiI", 0xF993FAC9, 0, 0)                                                      # v2 none
    raw += struct.pack("<Iii1q", 0xF993FACA, 0, 1, 2) + struct.pack("<iii", 1, 0, 6) + v.tobytes()    # v3 int64 vector
    raw += struct.pack("<Q", 0)
    p = tmp_path / "n.params"
    p.write_bytes(raw)
    out = ck.nd_load(str(p))
    assert len(out) == 4 and out[0].size == 0 and out[2].size == 0
    assert out[1].shape == () and float(out[1]) == 2.5
    assert np.array_equal(out[3], v)
    bad = struct.pack("<QQQ", 0x112, 0, 1) + struct.pack("<Iii", 0xF993FACA, 0, -2)
    (tmp_path / "bad.params").write_bytes(bad)
    with pytest.raises(ValueError):
        ck.nd_load(str(tmp_path / "bad.params"))
