"""In-situ per-operator timing of the hot path: every kernel-launching call of `rangedet_b200.ops` is bracketed by a
pair of CUDA events recorded on the launching stream (torch's current stream, which is the stream the C-ABI call
receives), with NO host synchronisation between calls -- the kernels run back to back in the order and cache state of
the real step; only at the end are the events read.  Used by bench.py for the roofline table (algorithmic flops /
bytes per call from the tensor shapes) and by scripts/.  CUDA-graph replays cannot be bracketed per
kernel, so the step being measured must run eagerly (train.GraphedTrainStep(capture=False) uses the same buffers,
flat plumbing and kernels as the captured step).

    with OpTimer() as t:
        step.train_step(data, coord)
    rows = t.rows()          # [{op, key, n, ms, flops, bytes}], torch.cuda.synchronize() inside
"""
from collections import OrderedDict

import torch

from . import ops

_E = 2   # bytes per stored activation element (bf16 / fp16)


def _nhwc(t):
    N, Hp, Wp, C = t.shape
    return N, Hp - 2, Wp - 2, C


def _conv(x, w, *a, **k):
    N, H, W, Ci = _nhwc(x)
    taps, Co, _ = w.shape
    s = k.get("stride_w", 1)
    res = k.get("residual_pad") is not None
    by = _E * (N * H * W * Ci + N * H * (W // s) * Co * (2 if res else 1) + taps * Co * Ci)
    return ("%dx%d %d->%d @%d s%d" % (3 if taps == 9 else 1, 3 if taps == 9 else 1, Ci, Co, W, s),
            2.0 * N * H * (W // s) * Ci * Co * taps, by)


def _conv_stats(x, w, out=None, stride_w=1, ws=None):
    key, fl, by = _conv(x, w, stride_w=stride_w)
    return key + " +stats", fl, by


def _conv_bwdstats(x, w, bn_z, bn_coef, bn_mask_mode, out=None, ws=None):
    key, fl, by = _conv(x, w)
    N, H, W, Co = _nhwc(bn_z)
    return key + " +bwd sums", fl, by + _E * N * H * W * Co      # + the z tile of the BatchNorm below


def _finalize(partial, nslots, N, H, W, C, *a, **k):
    return ("C%d" % C, 0.0, 4.0 * nslots * 2 * C)


def _slice(x, w, out, c_off, **k):
    N, H, W, Ci = _nhwc(x)
    taps, Co, _ = w.shape
    return ("slice %d->%d @%d" % (Ci, Co, W), 2.0 * N * H * W * Ci * Co * taps, _E * N * H * W * (Ci + Co))


def _deconv(x, w, *a, **k):
    N, H, W, Ci = _nhwc(x)
    taps, Co, _ = w.shape
    S = 4 if taps == 24 else 2
    res = k.get("residual_pad") is not None
    return ("deconv k%d %d->%d @%d" % (taps // 3, Ci, Co, W), 2.0 * N * H * W * Ci * Co * taps,
            _E * (N * H * W * Ci + N * H * W * S * Co * (2 if res else 1)))


def _wgrad(a, b, ksize, stride_w=1, out=None):
    N, H, W, CA = _nhwc(a)
    CB = b.shape[3]
    return ("k%d %dx%d @%d s%d" % (ksize, CA, CB, W, stride_w), 2.0 * N * H * W * CA * CB * ksize * ksize,
            _E * N * H * W * (CA + CB * stride_w) + 4 * ksize * ksize * CA * CB)


def _stats(z, *a, **k):
    N, H, W, C = _nhwc(z)
    return ("C%d @%d" % (C, W), 0.0, _E * N * H * W * C)


def _bn_fwd(z, coef, relu=True, res_before=None, res_after=None, out=None):
    N, H, W, C = _nhwc(z)
    nres = (res_before is not None) + (res_after is not None)
    return ("C%d @%d res%d" % (C, W, nres), 0.0, _E * N * H * W * C * (2 + nres))


def _bn_bwd(dy, z, coef, mask_mode, y_mask=None, dz_halo_w=1, dz_out=None, want_g=False, g_out=None, dgb_out=None, sums=None):
    N, H, W, C = _nhwc(z)
    rd = 2 + (1 if (mask_mode == 1 and y_mask is not None) else 0)    # dy, z (, y): read by the reduction AND the apply pass
    passes = 1 if sums is not None else 2                              # sums from the conv above: apply pass only
    return ("C%d @%d mask%d%s" % (C, W, mask_mode, " apply" if sums is not None else ""), 0.0,
            _E * N * H * W * C * (passes * rd + 1 + (1 if want_g else 0)))


def _sums(x, out=None):
    N, H, W, C = _nhwc(x)
    return ("C%d @%d" % (C, W), 0.0, _E * N * H * W * C)


def _add(x0, x1, out=None):
    N, H, W, C = _nhwc(x0)
    return ("C%d @%d" % (C, W), 0.0, 3 * _E * N * H * W * C)


def _to_nchw(src, channels=None, tap_major=False, out=None):
    N, H, W, Cs = _nhwc(src)
    C = channels or Cs
    return ("C%d @%d" % (C, W), 0.0, N * H * W * C * (_E + 4))


def _to_nhwc(src, out, tap_major=False):
    N, C, H, W = src.shape
    return ("C%d @%d" % (C, W), 0.0, N * H * W * C * (_E + 4))


def _meta_fwd(data, coord, *a, **k):
    B, C, H, W = data.shape
    return ("C%d @%d" % (C, W), 39.2e3 * B * H * W, B * H * W * (C * 4 + 12 + 9 * C * _E))


def _meta_bwd(go, data, coord, *a, **k):
    B, C, H, W = data.shape
    # op boundary of meta_kernel.py:232-239 (fp32): grad_out once + data + coord in, grad_data out (2828 B/px)
    return ("C%d @%d" % (C, W), 3 * 39.2e3 * B * H * W, B * H * W * (9 * C * 4 + C * 4 + 12 + C * 4))


def _meta_bwd_nhwc(go_pad, data, coord, *a, **k):
    B, C, H, W = data.shape
    # grad_out as the NHWC 2-byte tensor, read by both kernels; data + coord in, grad_data out
    return ("C%d @%d nhwc" % (C, W), 3 * 39.2e3 * B * H * W, B * H * W * (9 * C * _E + C * 4 + 12 + C * 4))


def _copy_ch(src, src_off, dst, dst_off, nchan):
    N, Hp, Wp, _ = src.shape
    return ("C%d" % nchan, 0.0, 2.0 * _E * N * Hp * Wp * nchan)


def _loss(cls_logit, reg_delta, *a, **k):
    B, _, H, W = reg_delta.shape
    return ("@%d" % W, 0.0, 224.0 * B * H * W)


def _loss_nhwc(cls_pad, reg_pad, *a, **k):
    B, Hp, Wp, _ = reg_pad.shape
    return ("@%d nhwc" % (Wp - 2), 0.0, 190.0 * B * (Hp - 2) * (Wp - 2))


def _gather(src, idx, out):
    return ("n%d" % idx.numel(), 0.0, idx.numel() * (4 + 4 + out.element_size()))


def _sgd(weight, *a):
    return ("n%d" % weight.numel(), 0.0, weight.numel() * 4 * 6)


# op name -> (family, metadata function); family "tensor" rows are judged against the bf16 tensor peak, "hbm" rows
# against the measured copy bandwidth
OPS = OrderedDict([
    ("conv2d_nhwc", ("conv", _conv)), ("conv2d_nhwc_stats", ("conv", _conv_stats)), ("bn_train_finalize", ("bn", _finalize)),
    ("conv2d_nhwc_bwdstats", ("conv", _conv_bwdstats)),
    ("conv2d_nhwc_slice", ("conv", _slice)), ("deconv2d_nhwc", ("conv", _deconv)),
    ("conv2d_wgrad", ("wgrad", _wgrad)),
    ("bn_train_stats", ("bn", _stats)), ("bn_act_fwd", ("bn", _bn_fwd)), ("bn_act_bwd", ("bn", _bn_bwd)),
    ("channel_sums", ("bn", _sums)), ("add_nhwc", ("bn", _add)),
    ("nhwc_to_nchw", ("layout", _to_nchw)), ("nchw_to_nhwc", ("layout", _to_nhwc)), ("copy_channels", ("layout", _copy_ch)),
    ("meta_kernel_forward_nhwc", ("meta", _meta_fwd)), ("meta_kernel_backward", ("meta", _meta_bwd)),
    ("meta_kernel_backward_nhwc", ("meta", _meta_bwd_nhwc)),
    ("rpn_loss", ("loss", _loss)), ("rpn_loss_nhwc", ("loss", _loss_nhwc)),
    ("gather_to_bf16", ("optim", _gather)), ("gather_f32", ("optim", _gather)), ("sgd_mom_update", ("optim", _sgd)),
])
BOUND = {"conv": "tensor", "wgrad": "tensor", "bn": "hbm", "layout": "hbm", "meta": "hbm", "loss": "latency", "optim": "hbm"}


class OpTimer(object):
    def __init__(self):
        self.events = []       # (op, key, flops, bytes, e0, e1)
        self._orig = {}

    def __enter__(self):
        for name, (_, meta) in OPS.items():
            if not hasattr(ops, name):
                continue
            orig = getattr(ops, name)
            self._orig[name] = orig
            setattr(ops, name, self._wrap(name, orig, meta))
        return self

    def _wrap(self, name, orig, meta):
        def f(*a, **k):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = orig(*a, **k)
            e1.record()
            key, fl, by = meta(*a, **k)
            self.events.append((name, key, fl, by, e0, e1))
            return r
        return f

    def __exit__(self, *exc):
        for name, orig in self._orig.items():
            setattr(ops, name, orig)
        self._orig = {}
        return False

    def rows(self):
        """One row per (op, shape key): calls, summed milliseconds, algorithmic flops and bytes."""
        torch.cuda.synchronize()
        agg = OrderedDict()
        for name, key, fl, by, e0, e1 in self.events:
            r = agg.setdefault((name, key), dict(op=name, key=key, family=OPS[name][0], n=0, ms=0.0, flops=0.0, bytes=0.0))
            r["n"] += 1
            r["ms"] += e0.elapsed_time(e1)
            r["flops"] += fl
            r["bytes"] += by
        return sorted(agg.values(), key=lambda r: -r["ms"])

    @staticmethod
    def families(rows):
        fam = OrderedDict()
        for r in rows:
            f = fam.setdefault(r["family"], dict(family=r["family"], bound=BOUND[r["family"]], n=0, ms=0.0, flops=0.0, bytes=0.0))
            for k in ("n", "ms", "flops", "bytes"):
                f[k] += r[k]
        for f in fam.values():
            s = f["ms"] * 1e-3
            f["TFLOPs"] = f["flops"] / s / 1e12 if s > 0 else 0.0
            f["GBps"] = f["bytes"] / s / 1e9 if s > 0 else 0.0
        return sorted(fam.values(), key=lambda f: -f["ms"])
