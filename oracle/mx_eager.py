"""An EAGER stand-in for the slice of MXNet's symbol API that the reference's graph code calls, on torch tensors.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Purpose: execute the reference's own graph-building Python --
rangedet/symbol/backbone/meta_kernel.py, dla_backbone.py, rangedet/symbol/head/builder.py, loss.py and the real
mxnext/{simple,complicate,initializer}.py -- UNMODIFIED in this container (MXNet itself is not installable: no
network), so that the torch restatements under oracle/ are checked against the reference's CODE and not only
against our reading of it.  What remains assumed is the semantics of each individual MXNet operator, restated here
one line each from MXNet's operator documentation (cited per op); the composition -- which op, which order, which
names, which shapes, which padding -- is the reference's.  Gradients come from torch autograd through these ops;
`MakeLoss` back-propagates its constant `grad_scale`.

Only usable where /root/reference exists; nothing on the GPU box imports this module (golden vectors generated
here are committed under tests/golden/).
"""
import contextlib
import importlib
import os
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

REF = os.environ.get("RD_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REF, "rangedet", "symbol", "head", "builder.py"))


class Env(object):
    """Bindings of one eager run: graph inputs + parameters + BN aux states by name."""
    current = None

    def __init__(self, bindings, training=True):
        self.b, self.training = bindings, training
        self.used = []

    def get(self, name):
        if name not in self.b:
            raise KeyError("eager mxnet: no binding for variable %r" % name)
        self.used.append(name)
        return self.b[name]


class S(object):
    """A 'symbol' that already has a value (or a variable resolved by name in the current Env)."""

    def __init__(self, t=None, name=None):
        self._t, self.name = t, name

    @property
    def t(self):
        return self._t if self._t is not None else Env.current.get(self.name)

    def _bin(self, o, f):
        return S(f(self.t, o.t if isinstance(o, S) else o))

    def __add__(self, o): return self._bin(o, lambda a, b: a + b)
    def __radd__(self, o): return self._bin(o, lambda a, b: b + a)
    def __sub__(self, o): return self._bin(o, lambda a, b: a - b)
    def __rsub__(self, o): return self._bin(o, lambda a, b: b - a)
    def __mul__(self, o): return self._bin(o, lambda a, b: a * b)
    def __rmul__(self, o): return self._bin(o, lambda a, b: b * a)
    def __truediv__(self, o): return self._bin(o, lambda a, b: a / b)
    def __rtruediv__(self, o): return self._bin(o, lambda a, b: b / a)
    def __neg__(self): return S(-self.t)


def _t(x):
    return x.t if isinstance(x, S) else x


def _pair(v):
    return tuple(v) if isinstance(v, (tuple, list)) else (v, v)


# ---- operators (mx.sym.*) -----------------------------------------------------------------------------------
def var(name, **kw):
    return S(None, name)


def Convolution(data, name=None, weight=None, bias=None, num_filter=None, kernel=None, stride=(1, 1), pad=(0, 0),
                dilate=(1, 1), num_group=1, workspace=None, no_bias=False, **kw):
    """mx.sym.Convolution: cross-correlation, weight (num_filter, C/group, kh, kw); parameters `<name>_weight/_bias`."""
    w = _t(weight) if weight is not None else Env.current.get(name + "_weight")
    b = None if no_bias else (_t(bias) if bias is not None else Env.current.get(name + "_bias"))
    assert w.shape[0] == num_filter and tuple(w.shape[2:]) == _pair(kernel), (name, tuple(w.shape), num_filter, kernel)
    return S(F.conv2d(_t(data), w, b, _pair(stride), _pair(pad), _pair(dilate), num_group), name)


def Deconvolution(data, name=None, weight=None, bias=None, num_filter=None, kernel=None, stride=(1, 1), pad=(0, 0),
                  dilate=(1, 1), num_group=1, workspace=None, no_bias=True, adj=(0, 0), **kw):
    """mx.sym.Deconvolution: transposed convolution, weight (C_in, num_filter/group, kh, kw),
    out = (in - 1) * stride - 2 * pad + kernel + adj."""
    w = _t(weight) if weight is not None else Env.current.get(name + "_weight")
    b = None if no_bias else (_t(bias) if bias is not None else Env.current.get(name + "_bias"))
    assert w.shape[1] * num_group == num_filter and tuple(w.shape[2:]) == _pair(kernel), (name, tuple(w.shape))
    return S(F.conv_transpose2d(_t(data), w, b, _pair(stride), _pair(pad), _pair(adj), num_group, _pair(dilate)), name)


def BatchNorm(data, gamma=None, beta=None, moving_mean=None, moving_var=None, name=None, fix_gamma=True,
              use_global_stats=False, momentum=0.9, eps=1e-3, axis=1, **kw):
    """mx.sym.BatchNorm over axis 1: training (and not use_global_stats) normalises with the BIASED batch variance and
    updates moving = moving * momentum + batch * (1 - momentum) (biased variance); otherwise uses the moving stats."""
    e = Env.current
    g = _t(gamma) if gamma is not None else e.get(name + "_gamma")
    b = _t(beta) if beta is not None else e.get(name + "_beta")
    mm = _t(moving_mean) if moving_mean is not None else e.get(name + "_moving_mean")
    mv = _t(moving_var) if moving_var is not None else e.get(name + "_moving_var")
    if fix_gamma:
        g = torch.ones_like(g)
    x = _t(data)
    if e.training and not use_global_stats:
        dims = [d for d in range(x.dim()) if d != 1]
        mean, varb = x.mean(dims), x.var(dims, unbiased=False)
        with torch.no_grad():
            mm.mul_(momentum).add_(mean.detach() * (1 - momentum))
            mv.mul_(momentum).add_(varb.detach() * (1 - momentum))
    else:
        mean, varb = mm, mv
    shp = [1, -1] + [1] * (x.dim() - 2)
    return S((x - mean.view(shp)) / torch.sqrt(varb.view(shp) + eps) * g.view(shp) + b.view(shp), name)


def Activation(data, name=None, act_type="relu", **kw):
    x = _t(data)
    f = {"relu": torch.relu, "sigmoid": torch.sigmoid, "tanh": torch.tanh,
         "softrelu": lambda v: F.softplus(v, threshold=1e9)}[act_type]   # softrelu: log(1 + exp(x))
    return S(f(x), name)


def reshape(data, shape=None, name=None, **kw):
    """mx.sym.reshape: 0 copies the input dimension, -1 infers; the other special codes are not used by the reference."""
    x = _t(data)
    assert all(s >= -1 for s in shape), shape
    return S(x.reshape([x.shape[i] if s == 0 else s for i, s in enumerate(shape)]), name)


def im2col(data, kernel=None, stride=(1, 1), dilate=(1, 1), pad=(0, 0), name=None, **kw):
    """mx.sym.im2col: (N, C, H, W) -> (N, C * kh * kw, L), channel-major then kernel position, zero padding."""
    return S(F.unfold(_t(data), _pair(kernel), _pair(dilate), _pair(pad), _pair(stride)), name)


def _cmp(f):
    return lambda lhs=None, rhs=None, name=None, **kw: S(f(_t(lhs), _t(rhs)).to(_t(lhs).dtype), name)


def MakeLoss(data, grad_scale=1.0, name=None, normalization="null", **kw):
    """mx.sym.MakeLoss: forward = data; backward = grad_scale per element (whatever arrives from above is ignored)."""
    assert normalization == "null"

    class _ML(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x):
            return x.clone()

        @staticmethod
        def backward(ctx, g):
            return torch.full_like(g, float(grad_scale))

    return S(_ML.apply(_t(data)), name)


def smooth_l1(data, scalar=1.0, name=None, **kw):
    """mx.sym.smooth_l1: 0.5 (s x)^2 if |x| < 1/s^2 else |x| - 0.5/s^2."""
    x, s2 = _t(data), float(scalar) ** 2
    return S(torch.where(x.abs() < 1.0 / s2, 0.5 * x * x * s2, x.abs() - 0.5 / s2), name)


def clip(data, a_min=None, a_max=None, name=None, **kw):
    """mx.sym.clip: gradient passes where a_min <= x <= a_max (torch.clamp does the same)."""
    return S(torch.clamp(_t(data), a_min, a_max), name)


def cast(data, dtype=None, name=None, **kw):
    # the harness runs the fp32 graphs (fp16=False); a cast to float16 would only change storage precision
    return S(_t(data).to({np.float32: torch.float32, np.float16: torch.float16}.get(dtype, torch.float32)), name)


def concat(*args, dim=1, name=None, **kw):
    return S(torch.cat([_t(a) for a in args], dim), name)


def slice_axis(data, axis=None, begin=None, end=None, name=None, **kw):
    x = _t(data)
    return S(x.narrow(axis, begin, (x.shape[axis] if end is None else end) - begin), name)


def zeros(shape=None, name=None, **kw):
    return S(torch.zeros(shape), name)


SYM_OPS = dict(
    var=var, Variable=var, Symbol=S, Group=lambda syms: list(syms), Convolution=Convolution, Deconvolution=Deconvolution,
    BatchNorm=BatchNorm, Activation=Activation, reshape=reshape, im2col=im2col, MakeLoss=MakeLoss, smooth_l1=smooth_l1,
    clip=clip, cast=cast, concat=concat, slice_axis=slice_axis, zeros=zeros,
    expand_dims=lambda data, axis=None, name=None, **kw: S(_t(data).unsqueeze(axis), name),
    squeeze=lambda data, axis=None, name=None, **kw: S(_t(data).squeeze(axis), name),
    transpose=lambda data, axes=None, name=None, **kw: S(_t(data).permute(*axes), name),
    elemwise_add=lambda lhs, rhs, name=None, **kw: S(_t(lhs) + _t(rhs), name),
    elemwise_sub=lambda lhs, rhs, name=None, **kw: S(_t(lhs) - _t(rhs), name),
    elemwise_mul=lambda lhs, rhs, name=None, **kw: S(_t(lhs) * _t(rhs), name),
    elemwise_div=lambda lhs, rhs, name=None, **kw: S(_t(lhs) / _t(rhs), name),
    broadcast_add=lambda lhs=None, rhs=None, name=None, **kw: S(_t(lhs) + _t(rhs), name),
    broadcast_minus=lambda lhs=None, rhs=None, name=None, **kw: S(_t(lhs) - _t(rhs), name),
    broadcast_sub=lambda lhs=None, rhs=None, name=None, **kw: S(_t(lhs) - _t(rhs), name),
    broadcast_mul=lambda lhs=None, rhs=None, name=None, **kw: S(_t(lhs) * _t(rhs), name),
    broadcast_div=lambda lhs=None, rhs=None, name=None, **kw: S(_t(lhs) / _t(rhs), name),
    broadcast_greater_equal=_cmp(torch.ge), broadcast_greater=_cmp(torch.gt), broadcast_equal=_cmp(torch.eq),
    sigmoid=lambda data, name=None, **kw: S(torch.sigmoid(_t(data)), name),
    exp=lambda data, name=None, **kw: S(torch.exp(_t(data)), name),
    log=lambda data, name=None, **kw: S(torch.log(_t(data)), name),
    abs=lambda data, name=None, **kw: S(torch.abs(_t(data)), name),
    power=lambda base, exp, name=None, **kw: S(torch.pow(_t(base), _t(exp)), name),
    sum=lambda data, axis=None, name=None, **kw: S(_t(data).sum() if axis is None else _t(data).sum(axis), name),
    stop_gradient=lambda data, name=None, **kw: S(_t(data).detach(), name),
    BlockGrad=lambda data, name=None, **kw: S(_t(data).detach(), name),
)


# ---- CustomOp registry + contrib ops (numpy side: the reference's own Python / compiled C++) -------------------
CUSTOM = {}


class CustomOp(object):
    def assign(self, dst, req, src):
        dst[...] = src


class CustomOpProp(object):
    def __init__(self, need_top_grad=False):
        self.need_top_grad_ = need_top_grad


def register(name):
    def deco(cls):
        CUSTOM[name] = cls
        return cls
    return deco


def Custom(*args, op_type=None, name=None, **kwargs):
    """mx.sym.Custom: tensors are matched to prop.list_arguments(), everything else goes to the Prop constructor as a
    string (that is how MXNet passes CustomOp attributes); outputs follow infer_shape; no gradient flows back."""
    from .ref_py import A
    cls = CUSTOM[op_type]
    tensors = {k: v for k, v in kwargs.items() if isinstance(v, S)}
    attrs = {k: str(v) for k, v in kwargs.items() if not isinstance(v, S)}
    prop = cls(**attrs)
    ins = [A(tensors[k].t.detach().numpy().astype(np.float32)) for k in prop.list_arguments()]
    _, out_shapes = prop.infer_shape([list(a.shape) for a in ins])[:2]
    outs = [A(np.zeros(tuple(s), np.float32)) for s in out_shapes]
    op = prop.create_operator(None, [a.shape for a in ins], [a.dtype for a in ins])
    op.forward(Env.current.training, ["write"] * len(outs), ins, outs, [])
    res = [S(torch.from_numpy(np.asarray(o).copy())) for o in outs]
    return res[0] if len(res) == 1 else res


def Decode3DBbox(bbox_deltas, pc_laser_frame, is_bin=False, name=None, **kw):
    """_contrib_Decode3DBbox through the reference's own functor compiled by oracle/build_ref.py (zero gradient)."""
    from . import reference
    d, p = _t(bbox_deltas).detach().numpy(), _t(pc_laser_frame).detach().numpy()
    return S(torch.from_numpy(reference().decode_3d_bbox(np.ascontiguousarray(d, np.float32),
                                                           np.ascontiguousarray(p, np.float32), is_bin=bool(is_bin))), name)


class _NS(types.ModuleType):
    """Namespace module: known ops, and a loud failure for anything the stand-in does not implement."""

    def __init__(self, name, ops):
        super().__init__(name)
        self.__dict__.update(ops)

    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)

        def missing(*a, **kw):
            raise NotImplementedError("eager mxnet stand-in: %s.%s is not implemented" % (self.__name__, k))
        return missing


class Initializer(object):
    def __init__(self, *a, **kw):
        pass


def _modules():
    from . import ref_py, reference
    mx = types.ModuleType("mxnet")
    sym = _NS("mxnet.symbol", SYM_OPS)
    sym.contrib = _NS("mxnet.symbol.contrib", dict(Decode3DBbox=Decode3DBbox))
    sym.Custom = Custom
    init = _NS("mxnet.initializer", dict(Initializer=Initializer, Normal=type("Normal", (Initializer,), {}),
                                         One=type("One", (Initializer,), {}), Zero=type("Zero", (Initializer,), {}),
                                         Constant=type("Constant", (Initializer,), {})))
    op = types.ModuleType("mxnet.operator")
    op.CustomOp, op.CustomOpProp, op.register = CustomOp, CustomOpProp, register
    stubs = ref_py._stub_modules(reference().rotated_iou)      # mxnet.numpy / mxnet.ndarray used inside the CustomOps
    nd, mnp = stubs["mxnet.ndarray"], stubs["mxnet.numpy"]
    mx.sym = mx.symbol = sym
    mx.init = mx.initializer = init
    mx.operator, mx.nd, mx.ndarray, mx.numpy = op, nd, nd, mnp
    return {"mxnet": mx, "mxnet.symbol": sym, "mxnet.initializer": init, "mxnet.operator": op, "mxnet.ndarray": nd,
            "mxnet.ndarray.contrib": stubs["mxnet.ndarray.contrib"], "mxnet.numpy": mnp}


_REF_PKGS = ("mxnext", "rangedet", "operator_py")


@contextlib.contextmanager
def reference_modules():
    """Inside the block `import mxnext`, `import rangedet.symbol...`, `import operator_py...` load the reference's
    files from /root/reference on top of the eager stand-in; afterwards sys.modules / sys.path are restored."""
    assert available(), "needs /root/reference"
    stubs = _modules()
    saved = {k: sys.modules.get(k) for k in list(stubs)}
    for k in list(sys.modules):
        if k.split(".")[0] in _REF_PKGS:
            saved[k] = sys.modules.pop(k)
    sys.modules.update(stubs)
    sys.path.insert(0, REF)
    try:
        yield importlib.import_module
    finally:
        sys.path.remove(REF)
        for k in list(sys.modules):
            if k.split(".")[0] in _REF_PKGS or k in stubs:
                sys.modules.pop(k, None)
        for k, v in saved.items():
            if v is not None:
                sys.modules[k] = v


@contextlib.contextmanager
def bound(bindings, training=True):
    prev = Env.current
    Env.current = Env(bindings, training)
    try:
        yield Env.current
    finally:
        Env.current = prev
