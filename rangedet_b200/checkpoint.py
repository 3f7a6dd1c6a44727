"""MXNet `.params` checkpoints <-> the parameter dictionary of this package.

Host-side mirror of utils/load_model.py:18-39 (`load_checkpoint(prefix, epoch)` -> `(arg_params, aux_params)`,
`get_latest_ckpt_epoch`): the reference reads `<prefix>-<epoch:04d>.params` with `mx.nd.load` and splits the keys
`arg:<name>` / `aux:<name>`.  MXNet is not a dependency here, so the file format is read directly.  Format as written
by `mx.nd.save` (MXNet 1.x / 2.0 `NDArray::Save`, dense arrays only; restated from the published format -- no MXNet
build is available here to produce a real file, so **parity unpinned**: the reader is pinned only against the writer
below, byte layout per the comments):

    uint64 0x112 (list magic) | uint64 0 | uint64 n_arrays
    n_arrays x { uint32 magic 0xF993FAC9 (v2) or 0xF993FACA (v3, numpy shape semantics)
                 int32 storage type (0 = dense)
                 uint32 ndim (v2) / int32 ndim (v3) | int64 dims[ndim]
                 int32 dev_type | int32 dev_id
                 int32 type_flag (0 f32, 1 f64, 2 f16, 3 u8, 4 i32, 5 i8, 6 i64)
                 raw little-endian data }
    uint64 n_names | n_names x { uint64 length | bytes }

Parameter names are the reference's (`res1_unit1_conv1_weight`, `rpn_cls_conv_0_lvl_0_bn_moving_mean`, ...), which
rangedet_b200.dla / rangedet_b200.train use unchanged, so `to_model_params` is a dtype / device conversion.
"""
import glob
import struct

import numpy as np

LIST_MAGIC = 0x112
V2_MAGIC, V3_MAGIC = 0xF993FAC9, 0xF993FACA
DTYPES = {0: np.float32, 1: np.float64, 2: np.float16, 3: np.uint8, 4: np.int32, 5: np.int8, 6: np.int64}
FLAGS = {np.dtype(v): k for k, v in DTYPES.items()}


def nd_load(path):
    """-> dict name -> numpy array (or a list when the file carries no names), like mx.nd.load."""
    buf = open(path, "rb").read()
    off = 0

    def rd(fmt):
        nonlocal off
        v = struct.unpack_from("<" + fmt, buf, off)
        off += struct.calcsize("<" + fmt)
        return v if len(v) > 1 else v[0]

    magic, _ = rd("QQ")
    if magic != LIST_MAGIC:
        raise ValueError("%s: not an MXNet NDArray list file (magic %#x)" % (path, magic))
    arrays = []
    for _ in range(rd("Q")):
        m = rd("I")
        if m not in (V2_MAGIC, V3_MAGIC):
            raise ValueError("%s: unsupported NDArray magic %#x (legacy v1 files are not handled)" % (path, m))
        stype = rd("i")
        if stype != 0:
            raise ValueError("%s: sparse storage type %d is not supported" % (path, stype))
        ndim = rd("I") if m == V2_MAGIC else rd("i")
        # NDArray::Save writes the shape and returns when is_none(): an unknown shape is ndim 0 in the legacy (v2)
        # convention and ndim -1 under np_shape (v3, where ndim 0 is a genuine scalar holding one element)
        if (m == V2_MAGIC and ndim == 0) or (m == V3_MAGIC and ndim == -1):
            arrays.append(np.zeros((0,), np.float32))
            continue
        if ndim < 0 or ndim > 32:
            raise ValueError("%s: bad ndim %d" % (path, ndim))
        shape = tuple(struct.unpack_from("<%dq" % ndim, buf, off)) if ndim else ()
        off += 8 * ndim
        rd("ii")   # context (dev_type, dev_id): ignored, the caller chooses the device
        flag = rd("i")
        if flag not in DTYPES:
            raise ValueError("%s: unknown type flag %d" % (path, flag))
        dt = np.dtype(DTYPES[flag]).newbyteorder("<")
        n = int(np.prod(shape)) if shape else 1
        arrays.append(np.frombuffer(buf, dt, n, off).reshape(shape).copy())
        off += n * dt.itemsize
    names = []
    for _ in range(rd("Q")):
        ln = rd("Q")
        names.append(buf[off:off + ln].decode("utf-8"))
        off += ln
    if not names:
        return arrays
    if len(names) != len(arrays):
        raise ValueError("%s: %d names for %d arrays" % (path, len(names), len(arrays)))
    return dict(zip(names, arrays))


def nd_save(path, data):
    """dict name -> array (or list of arrays) -> `.params` file in the v2 format above (as mx.nd.save writes it)."""
    names = list(data) if isinstance(data, dict) else []
    arrays = [np.ascontiguousarray(data[k]) for k in names] if names else [np.ascontiguousarray(a) for a in data]
    out = [struct.pack("<QQQ", LIST_MAGIC, 0, len(arrays))]
    for a in arrays:
        if a.dtype not in FLAGS:
            raise TypeError("unsupported dtype %s" % a.dtype)
        out.append(struct.pack("<Ii", V2_MAGIC, 0))
        out.append(struct.pack("<I%dq" % a.ndim, a.ndim, *a.shape))
        if a.ndim == 0:
            raise ValueError("0-d arrays cannot be stored in the v2 format")
        out.append(struct.pack("<iii", 1, 0, FLAGS[a.dtype]))   # cpu(0)
        out.append(a.astype(a.dtype.newbyteorder("<"), copy=False).tobytes())
    out.append(struct.pack("<Q", len(names)))
    for k in names:
        b = k.encode("utf-8")
        out.append(struct.pack("<Q", len(b)) + b)
    with open(path, "wb") as f:
        f.write(b"".join(out))


def get_latest_ckpt_epoch(prefix):   # utils/load_model.py:5-15
    ckpts = glob.glob(prefix + "*.params")
    if not ckpts:
        raise FileNotFoundError("can not find params starting with %s" % prefix)
    return max(int(p[p.rfind(".params") - 4:p.rfind(".params")]) for p in ckpts)


def load_checkpoint(prefix, epoch):   # utils/load_model.py:18-39
    """-> (arg_params, aux_params): dicts of name -> numpy array."""
    save_dict = nd_load("%s-%04d.params" % (prefix, epoch))
    arg_params, aux_params = {}, {}
    for k, v in save_dict.items():
        tp, name = k.split(":", 1)
        if tp == "arg":
            arg_params[name] = v
        if tp == "aux":
            aux_params[name] = v
    return arg_params, aux_params


def save_checkpoint(prefix, epoch, arg_params, aux_params):
    """Inverse of load_checkpoint (what mx.model.save_checkpoint writes next to the -symbol.json)."""
    d = {"arg:" + k: np.asarray(v) for k, v in arg_params.items()}
    d.update({"aux:" + k: np.asarray(v) for k, v in aux_params.items()})
    nd_save("%s-%04d.params" % (prefix, epoch), d)


def to_model_params(arg_params, aux_params, device="cuda"):
    """-> the fp32 torch parameter dictionary rangedet_b200.dla / rangedet_b200.train consume (same names; fp16
    checkpoints of the shipped fp16 configs are widened)."""
    import torch
    P = {}
    for d in (arg_params, aux_params):
        for k, v in d.items():
            P[k] = torch.from_numpy(np.asarray(v, dtype=np.float32).copy()).to(device)
    return P


def from_model_params(P):
    """torch parameter dictionary -> (arg_params, aux_params) numpy dicts (aux = BatchNorm moving statistics)."""
    arg, aux = {}, {}
    for k, v in P.items():
        (aux if k.endswith(("_moving_mean", "_moving_var")) else arg)[k] = v.detach().float().cpu().numpy()
    return arg, aux
