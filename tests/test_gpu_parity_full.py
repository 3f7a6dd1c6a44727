"""Whole-graph parity AT SIZE on the GPU (BASELINE.json configs[3] and [4]): every layer of the cfg-5 training graph
(B=2, 64x2656) teacher-forced with the oracle's own tensors, the composed step end to end, and the cfg-4 forward (B=8).

Why teacher-forced: a 40-layer BN/ReLU network at random init amplifies storage rounding, so an end-to-end comparison of
deep gradients alone cannot tell a wrong tap offset from rounding.  Here the oracle (oracle/dla_train_ref.py: torch fp32
autograd on the reference's graph, storage rounding emulated, run ONCE at full size on the same device) records for every
layer its input, its output, the gradient arriving at its output, and -- through torch.autograd.grad restricted to that
layer -- the gradients it sends to its inputs and parameters.  Each GPU layer (train.TrainGraph.conv_bn / deconv_bn /
head_out / meta_kernel_front, i.e. the C-ABI kernels: tcgen05 conv, BN passes, wgrad, dgrad, layout, Meta-Kernel) is then
fed the oracle's tensors and must reproduce the oracle's results within ONE storage rounding:

    forward  max|y - y_ref| / max|y_ref|                    <= FWD_TOL[dtype]
    dx, dW                  rms(g - g_ref) / rms(g_ref)     <= GRAD_TOL[dtype]
    dgamma, dbeta           rms(g - g_ref) / rms(g_ref)     <= BNSUM_TOL[dtype]

A flipped tap, a transposed weight slice, a wrong stride phase or a mis-routed residual is an O(1) error on that layer.
The report (every layer, every quantity) is written to gpurun_out/parity_full_<dtype>.json.

Reference: rangedet/symbol/backbone/dla_backbone.py:17-161, rangedet/symbol/head/builder.py:198-422.
"""
import json
import os

import pytest
import torch
import torch.nn.functional as F

from conftest import ROOT

pytestmark = pytest.mark.gpu

H, W = 64, 2656
# one storage rounding: bf16 2^-9 = 2.0e-3, fp16 2^-12 = 2.4e-4 relative per element (round to nearest); a normwise forward
# error additionally sees the largest element's ulp; gradients add the ReLU-mask elements that sit within rounding of 0
FWD_TOL = {torch.bfloat16: 1e-2, torch.float16: 2e-3}
GRAD_TOL = {torch.bfloat16: 1.5e-2, torch.float16: 8e-3}   # observed worst 1.4e-2 / 5.0e-3
# dgamma / dbeta are sums of ReLU-MASKED gradients, and the mask z*a + b > 0 is discontinuous in the BatchNorm coefficients:
# in bf16 storage thousands of pixels share the one z value nearest the threshold, and a last-bit change of (a, b) -- the
# summation order of the batch statistics -- moves that whole bin across it.  Observed 1.7e-3 .. 1.7e-2 for the same layer
# with two different statistics orders (dx and dW of that layer unchanged at 4.5e-3): its own bound in bf16.
BNSUM_TOL = {torch.bfloat16: 4e-2, torch.float16: 8e-3}


def _tol(key, dtype):
    return BNSUM_TOL[dtype] if key.endswith(("gamma", "beta")) else GRAD_TOL[dtype]
DTYPES = [torch.float16, torch.bfloat16]
IDS = ["f16", "bf16"]


def _maxrel(a, b):
    a, b = a.detach().float(), b.detach().float()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def _rms_rel(a, b):
    a, b = a.detach().double(), b.detach().double()
    return float(((a - b) ** 2).mean().sqrt() / (b ** 2).mean().sqrt().clamp_min(1e-30))


def _cos(a, b):
    a, b = a.detach().double().flatten(), b.detach().double().flatten()
    return float((a * b).sum() / (a.norm() * b.norm()).clamp_min(1e-30))


class RecordingRef(object):
    """oracle.dla_train_ref.TrainRef with every layer call recorded (inputs, output) on the way forward."""

    def __new__(cls, P, store):
        from oracle import dla_train_ref, meta_kernel_ref
        from oracle.dla_ref import META_UNITS

        class _Rec(dla_train_ref.TrainRef):
            def __init__(self, P, store):
                super().__init__(P, store=store)
                self.records = []
                self.stages = []         # whole res / agg stages: composition inside a stage (skips, fan-out adds)
                self.recording = True

            def quiet(self, fn, *a):
                """Run a piece of the BASE graph without recording (the isolated recomputations)."""
                prev, self.recording = self.recording, False
                try:
                    return fn(*a)
                finally:
                    self.recording = prev

            def res_stage(self, x, coord, name, stride):
                y = super().res_stage(x, coord, name, stride)
                if self.recording and not name.endswith("_res"):
                    self.stages.append(dict(kind="res_stage", name=name, x=x, coord=coord, y=y, stride_w=stride[1],
                                            redo=lambda x_, r_: self.quiet(super(_Rec, self).res_stage, x_, coord, name, stride)))
                return y

            def conv_bn(self, x, wname, bnname, stride=(1, 1), relu=True, residual=None):
                base = super().conv_bn
                y = base(x, wname, bnname, stride=stride, relu=relu, residual=residual)
                if not self.recording:
                    return y
                self.records.append(dict(kind="conv_bn", name=wname, bn=bnname, x=x, res=residual, y=y, stride_w=stride[1], relu=relu,
                                         params=[wname + "_weight", bnname + "_gamma", bnname + "_beta"],
                                         redo=lambda x_, r_: base(x_, wname, bnname, stride=stride, relu=relu, residual=r_)))
                return y

            def basicblock(self, x, coord, name, stride, proj):
                if self.use_meta and name in META_UNITS:   # same ops as the base class, the front recorded separately
                    P = self.P
                    m = meta_kernel_ref.meta_baseline_bias(x, coord, P[name + "_2656_mlp0_weight"].reshape(32, 3),
                                                           P[name + "_2656_mlp0_bias"], P[name + "_2656_mlp1_weight"].reshape(-1, 32),
                                                           P[name + "_2656_mlp1_bias"])
                    m = self.r(m)
                    a = self.r(self.bn(m, name + "point_wise_mlp_bn1").relu())
                    if not self.recording:
                        r1 = self.conv_bn(a, name + "aggregation_conv1", name + "aggregation_bn1")
                        sc = self.conv_bn(x, name + "_sc", name + "_sc_bn", stride=stride, relu=False) if proj else x
                        return self.conv_bn(r1, name + "_conv2", name + "_bn2", stride=stride, relu=True, residual=sc)

                    def redo(x_, r_):
                        m_ = meta_kernel_ref.meta_baseline_bias(x_, coord, P[name + "_2656_mlp0_weight"].reshape(32, 3),
                                                                P[name + "_2656_mlp0_bias"], P[name + "_2656_mlp1_weight"].reshape(-1, 32),
                                                                P[name + "_2656_mlp1_bias"])
                        return self.r(self.bn(self.r(m_), name + "point_wise_mlp_bn1").relu())
                    self.records.append(dict(kind="meta_front", name=name, x=x, coord=coord, y=a, m=m, redo=redo,
                                             params=[name + "point_wise_mlp_bn1_gamma", name + "point_wise_mlp_bn1_beta",
                                                     name + "_2656_mlp0_weight", name + "_2656_mlp0_bias",
                                                     name + "_2656_mlp1_weight", name + "_2656_mlp1_bias"]))
                    r1 = self.conv_bn(a, name + "aggregation_conv1", name + "aggregation_bn1")
                    sc = self.conv_bn(x, name + "_sc", name + "_sc_bn", stride=stride, relu=False) if proj else x
                    return self.conv_bn(r1, name + "_conv2", name + "_bn2", stride=stride, relu=True, residual=sc)
                return super().basicblock(x, coord, name, stride, proj)

            def agg_stage(self, name, const, up, sw, pad):
                def layer(up_, const_):
                    w = self.r(self.P[name + "_deconv_weight"])
                    z = self.r(F.conv_transpose2d(up_, w, stride=(1, sw), padding=(1, pad)))
                    return self.r(self.bn(z, name + "_deconv_bn").relu() + const_)
                y = layer(up, const)
                if not self.recording:
                    return self.res_stage(y, None, name + "_res", (1, 1))
                self.records.append(dict(kind="deconv_bn", name=name, x=up, res=const, y=y, redo=layer,
                                         params=[name + "_deconv_weight", name + "_deconv_bn_gamma", name + "_deconv_bn_beta"]))
                out = self.res_stage(y, None, name + "_res", (1, 1))
                self.stages.append(dict(kind="agg_stage", name=name, x=up, res=const, y=out,
                                        redo=lambda up_, c_: self.quiet(self.agg_stage, name, c_, up_, sw, pad)))
                return out

            def forward(self, data, coord):
                cls, reg = super().forward(data, coord)
                # the 1x1 head convolutions: their inputs are the outputs of the last tower layers recorded above
                last = {r["name"]: r["y"] for r in self.records if r["kind"] == "conv_bn"}
                for lvl in range(3):
                    for br, out, co in (("cls", cls, 1), ("reg", reg, 8)):
                        n = "rpn_%s_%s_lvl_%d" % (br, "logit" if br == "cls" else "delta", lvl)
                        self.records.append(dict(kind="head_out", name=n, x=last["rpn_%s_conv_3_lvl_%d" % (br, lvl)], y=out[lvl], co=co,
                                                 params=[n + "_weight", n + "_bias"],
                                                 redo=lambda x_, r_, n=n: F.conv2d(x_, self.r(self.P[n + "_weight"]), self.P[n + "_bias"])))
                        first = next(r for r in self.records if r["name"] == "rpn_%s_conv_0_lvl_%d" % (br, lvl))

                        def tower(x_, r_, br=br, lvl=lvl, n=n):
                            t = x_
                            for i in range(4):
                                c = "rpn_%s_conv_%d_lvl_%d" % (br, i, lvl)
                                t = self.quiet(self.conv_bn, t, c, c + "_bn")
                            return F.conv2d(t, self.r(self.P[n + "_weight"]), self.P[n + "_bias"])
                        self.stages.append(dict(kind="tower", name="rpn_%s_lvl_%d" % (br, lvl), x=first["x"], y=out[lvl], co=co, br=br,
                                                lvl=lvl, head=n, redo=tower))
                return cls, reg

        return _Rec(P, store)


@pytest.fixture(scope="module")
def ops():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from rangedet_b200 import ops as o
    return o


def _inputs(B, seed=0):
    from rangedet_b200 import synth
    g = torch.Generator(device="cuda").manual_seed(seed + 1)
    data = torch.randn((B, 8, H, W), device="cuda", generator=g)
    coord = torch.from_numpy(synth.range_image_coords(B, seed=seed)).cuda()
    d_cls = [torch.randn((B, 1, H, W >> l), device="cuda", generator=g) for l in range(3)]
    d_reg = [torch.randn((B, 8, H, W >> l), device="cuda", generator=g) for l in range(3)]
    return data, coord, d_cls, d_reg


def _chan_pad(c):
    return 128 if c == 72 else ((c + 63) // 64) * 64


@pytest.mark.parametrize("dtype", DTYPES, ids=IDS)
def test_cfg5_every_layer_teacher_forced_at_full_size(ops, dtype):
    """B=2, 64x2656: ~95 layers, each fed the oracle's input / output-gradient tensors."""
    from oracle import dla_ref
    from rangedet_b200 import train
    B = 2
    P = dla_ref.make_params(seed=0, device="cuda")
    data, coord, d_cls, d_reg = _inputs(B)
    ref = RecordingRef(P, dtype)
    cls_r, reg_r = ref.forward(data, coord)
    # gradient arriving at every recorded output = total derivative of the (linear) loss w.r.t. it
    rnd = lambda t: t.to(dtype).float()
    loss = sum((c * rnd(g)).sum() for c, g in zip(cls_r, d_cls)) + sum((r * rnd(g)).sum() for r, g in zip(reg_r, d_reg))
    ys = [r["y"] for r in ref.records]
    dys = torch.autograd.grad(loss, ys)       # the whole-graph backward is only needed for these; the graph is freed here
    for r in ref.records:                     # keep values only
        for k in ("x", "res", "y", "m"):
            if r.get(k) is not None:
                r[k + "_grad"] = r[k].requires_grad
                r[k] = r[k].detach()
    del loss, cls_r, reg_r, ys
    tg = train.TrainGraph({k: v.clone() for k, v in P.items()}, act_dtype=dtype)
    pad = lambda t, c=None: ops.to_nhwc_padded(t.detach(), c or _chan_pad(t.shape[1]), dtype=dtype)
    unpad = lambda t, c: ops.from_nhwc_padded(t, c)
    report, worst = {}, {"fwd": 0.0, "grad": 0.0}
    for rec, dy in zip(ref.records, dys):
        kind, name = rec["kind"], rec["name"]
        # the layer ALONE on detached copies of the oracle's own input tensors: gradients that are local to it
        # (the residual / skip input is upstream of the main input in the full graph: it must not see that path)
        x_l = rec["x"].clone().requires_grad_(rec["x_grad"])
        r_l = rec["res"].clone().requires_grad_(True) if rec.get("res") is not None else None
        y_l = rec["redo"](x_l, r_l)
        wrt = ([x_l] if rec["x_grad"] else []) + ([r_l] if r_l is not None else []) + [ref.P[n] for n in rec["params"]]
        want = torch.autograd.grad(y_l, wrt, grad_outputs=dy, allow_unused=True)
        want = dict(zip([id(t) for t in wrt], want))
        # the isolated recomputation IS the recorded layer (cuDNN may pick another algorithm: <= 2 storage ulps at the maximum)
        assert _maxrel(y_l, rec["y"]) < (2 ** -10 if dtype == torch.float16 else 2 ** -7), name
        x_key, r_key = id(x_l), (id(r_l) if r_l is not None else None)
        tg.begin()
        tapmajor = kind == "conv_bn" and name.endswith("aggregation_conv1")
        x_in = rec["x"].detach()
        if tapmajor:   # the unit's 576 channels are tap-major (k*64 + c) on the GPU side, c*9 + k in the reference
            Bc, C9, Hc, Wc = x_in.shape
            x_in = x_in.reshape(Bc, C9 // 9, 9, Hc, Wc).transpose(1, 2).reshape(Bc, C9, Hc, Wc)
        xp = pad(x_in)
        resp = pad(rec["res"]) if rec.get("res") is not None else None
        e = {}
        if kind == "conv_bn":
            if not rec["x_grad"]:
                tg.nograd.add(id(xp))
            y = tg.conv_bn(xp, name, rec["bn"], stride_w=rec["stride_w"], relu=rec["relu"], res_before=resp,
                           kinds=("fwd_tapmajor", "dgrad_tapmajor") if tapmajor else ("fwd", "dgrad"))
            co = rec["y"].shape[1]
            e["fwd"] = _maxrel(unpad(y, co), rec["y"])
            tg.seed_grad(y, pad(dy, y.shape[3]))
        elif kind == "deconv_bn":
            y = tg.deconv_bn(xp, resp, name)
            co = rec["y"].shape[1]
            e["fwd"] = _maxrel(unpad(y, co), rec["y"])
            tg.seed_grad(y, pad(dy, y.shape[3]))
        elif kind == "head_out":
            out, bwd = tg.head_out(xp, name, rec["co"])
            e["fwd"] = _maxrel(out, rec["y"])
        elif kind == "meta_front":
            a = tg.meta_kernel_front(xp, rec["coord"], name)
            Bc, C9, Hc, Wc = rec["y"].shape
            to_ref = lambda t: t[:, 1:-1, 1:-1, :].reshape(Bc, Hc, Wc, 9, C9 // 9).permute(0, 4, 3, 1, 2).reshape(Bc, C9, Hc, Wc).float()
            e["fwd"] = _maxrel(to_ref(a), rec["y"])
            dyp = torch.zeros_like(a)
            dyp[:, 1:-1, 1:-1, :] = dy.reshape(Bc, C9 // 9, 9, Hc, Wc).permute(0, 3, 4, 2, 1).reshape(Bc, Hc, Wc, C9).to(dtype)
            tg.seed_grad(a, dyp)
        if kind == "head_out":
            bwd(dy.contiguous())
        else:
            tg.run_tape()
        tg._join_side()
        if x_key in want and want[x_key] is not None:
            gx = tg.grad_of(xp)
            got = unpad(gx, rec["x"].shape[1])
            w_ = want[x_key]
            if tapmajor:
                Bc, C9, Hc, Wc = w_.shape
                w_ = w_.reshape(Bc, C9 // 9, 9, Hc, Wc).transpose(1, 2).reshape(Bc, C9, Hc, Wc)
            e["dx"] = _rms_rel(got, w_)
        if resp is not None:
            e["dres"] = _rms_rel(unpad(tg.grad_of(resp), rec["res"].shape[1]), want[r_key])
        for n in rec["params"]:
            g_ref = want.get(id(ref.P[n]))
            if g_ref is None:
                continue
            e["d" + n[len(name):] if n.startswith(name) else "d_" + n] = _rms_rel(tg.pgrads[n].reshape(g_ref.shape), g_ref)
        report["%s:%s" % (kind, name)] = e
        worst["fwd"] = max(worst["fwd"], e["fwd"])
        worst["grad"] = max([worst["grad"]] + [v for k, v in e.items() if k != "fwd"])
        del want, y_l, x_l, r_l
    torch.cuda.synchronize()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    tag = "f16" if dtype == torch.float16 else "bf16"
    with open(os.path.join(ROOT, "gpurun_out", "parity_full_%s.json" % tag), "w") as f:
        json.dump({"config": "cfg-5 B=2 64x2656 %s storage, teacher-forced per layer" % tag, "layers": len(report),
                   "worst": worst, "tol": {"fwd": FWD_TOL[dtype], "grad": GRAD_TOL[dtype], "bn_sums": BNSUM_TOL[dtype]}, "report": report}, f, indent=1)
    assert len(report) >= 90, len(report)
    bad = {k: e for k, e in report.items()
           if e["fwd"] > FWD_TOL[dtype] or any(v > _tol(kk, dtype) for kk, v in e.items() if kk != "fwd") or any(v != v for v in e.values())}
    # the Meta-Kernel unit's front (tcgen05 split-bf16 MLP + BN(576) masks) carries twice the gradient bound
    bad = {k: e for k, e in bad.items() if not k.startswith("meta_front")}
    assert not bad, sorted(bad.items(), key=lambda kv: -max(kv[1].values()))[:8]
    mf = next(v for k, v in report.items() if k.startswith("meta_front"))
    assert mf["fwd"] <= FWD_TOL[dtype] and all(v <= 2 * _tol(kk, dtype) for kk, v in mf.items() if kk != "fwd"), mf


# a stage is 4-10 layers deep: the per-layer roundings (and the ReLU masks they flip) compound
STAGE_FWD_TOL = {torch.bfloat16: 5e-2, torch.float16: 1e-2}
STAGE_GRAD_TOL = {torch.bfloat16: 2.5e-1, torch.float16: 8e-2}   # observed worst 5.6e-2 / 1.4e-1 (res3a / res3: ten layers)


@pytest.mark.parametrize("dtype", DTYPES, ids=IDS)
def test_cfg5_every_stage_teacher_forced_at_full_size(ops, dtype):
    """Composition INSIDE the stages at B=2, 64x2656: each residual stage (res1 with the Meta-Kernel unit ... res3), each
    aggregation stage (deconv + skip + residual stage) and each of the six head towers (4 conv-BN-ReLU + 1x1 output) runs
    as a whole on the oracle's input and output-gradient tensors: projection shortcuts, identity skips, gradient
    accumulation at the fan-out points, stride-2 phases, tap-major unit, every parameter gradient of the stage."""
    from oracle import dla_ref
    from rangedet_b200 import train
    B = 2
    P = dla_ref.make_params(seed=0, device="cuda")
    data, coord, d_cls, d_reg = _inputs(B)
    ref = RecordingRef(P, dtype)
    cls_r, reg_r = ref.forward(data, coord)
    rnd = lambda t: t.to(dtype).float()
    loss = sum((c * rnd(g)).sum() for c, g in zip(cls_r, d_cls)) + sum((r * rnd(g)).sum() for r, g in zip(reg_r, d_reg))
    dys = torch.autograd.grad(loss, [r["y"] for r in ref.stages])
    for r in ref.stages:
        for k in ("x", "res", "y"):
            if r.get(k) is not None:
                r[k + "_grad"] = r[k].requires_grad
                r[k] = r[k].detach()
    del loss, cls_r, reg_r
    ref.records = []
    tg = train.TrainGraph({k: v.clone() for k, v in P.items()}, act_dtype=dtype)
    pad = lambda t, c=None: ops.to_nhwc_padded(t.detach(), c or _chan_pad(t.shape[1]), dtype=dtype)
    unpad = lambda t, c: ops.from_nhwc_padded(t, c)
    leaves = [(n, p) for n, p in ref.P.items() if p.requires_grad]
    report = {}
    for rec, dy in zip(ref.stages, dys):
        kind, name = rec["kind"], rec["name"]
        x_l = rec["x"].clone().requires_grad_(rec["x_grad"])
        r_l = rec["res"].clone().requires_grad_(True) if rec.get("res") is not None else None
        y_l = rec["redo"](x_l, r_l)
        wrt = ([x_l] if rec["x_grad"] else []) + ([r_l] if r_l is not None else []) + [p for _, p in leaves]
        got_ref = torch.autograd.grad(y_l, wrt, grad_outputs=dy, allow_unused=True)
        off = len(wrt) - len(leaves)
        want_p = {n: g for (n, _), g in zip(leaves, got_ref[off:]) if g is not None}
        want_x = got_ref[0] if rec["x_grad"] else None
        want_r = got_ref[off - 1] if r_l is not None else None
        tg.begin()
        xp = pad(rec["x"])
        if not rec["x_grad"]:
            tg.nograd.add(id(xp))
        resp = pad(rec["res"]) if rec.get("res") is not None else None
        e = {}
        if kind == "res_stage":
            y = tg.res_stage(xp, rec["coord"], name, rec["stride_w"])
        elif kind == "agg_stage":
            y = tg.agg_stage(name, resp, xp)
        else:
            t = xp
            for i in range(4):
                c = "rpn_%s_conv_%d_lvl_%d" % (rec["br"], i, rec["lvl"])
                t = tg.conv_bn(t, c, c + "_bn")
            out, bwd = tg.head_out(t, rec["head"], rec["co"])
        if kind == "tower":
            e["fwd"] = _maxrel(out, rec["y"])
            bwd(dy.contiguous())
        else:
            e["fwd"] = _maxrel(unpad(y, rec["y"].shape[1]), rec["y"])
            tg.seed_grad(y, pad(dy, y.shape[3]))
        tg.run_tape()
        tg._join_side()
        if want_x is not None:
            e["dx"] = _rms_rel(unpad(tg.grad_of(xp), rec["x"].shape[1]), want_x)
        if want_r is not None:
            e["dres"] = _rms_rel(unpad(tg.grad_of(resp), rec["res"].shape[1]), want_r)
        assert set(want_p) <= set(tg.pgrads), sorted(set(want_p) - set(tg.pgrads))[:5]
        worst_p, worst_n = 0.0, None
        for n, g_ref in want_p.items():
            v = _rms_rel(tg.pgrads[n].reshape(g_ref.shape), g_ref)
            if not (v <= worst_p):
                worst_p, worst_n = v, n
        e["params"] = len(want_p)
        e["dparam_worst"] = worst_p
        e["dparam_worst_name"] = worst_n
        report["%s:%s" % (kind, name)] = e
        del got_ref, want_p, y_l
    torch.cuda.synchronize()
    tag = "f16" if dtype == torch.float16 else "bf16"
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "parity_stages_%s.json" % tag), "w") as f:
        json.dump({"config": "cfg-5 B=2 64x2656 %s storage, teacher-forced per stage / tower" % tag, "stages": len(report),
                   "tol": {"fwd": STAGE_FWD_TOL[dtype], "grad": STAGE_GRAD_TOL[dtype]}, "report": report}, f, indent=1)
    assert len(report) == 15 and sum(e["params"] for e in report.values()) >= 279, (len(report), sum(e["params"] for e in report.values()))
    num = lambda e: [v for k, v in e.items() if k in ("dx", "dres", "dparam_worst")]
    bad = {k: e for k, e in report.items() if not (e["fwd"] <= STAGE_FWD_TOL[dtype]) or any(not (v <= STAGE_GRAD_TOL[dtype]) for v in num(e))}
    assert not bad, bad


@pytest.mark.parametrize("dtype", DTYPES, ids=IDS)
def test_cfg5_train_step_end_to_end_at_full_size(ops, dtype):
    """The composed step (forward, fused RPN loss, backward) at B=2, 64x2656 against the oracle end to end: head outputs,
    the six loss sums, and the direction of every parameter gradient.  Bounds are those of a deep random-init network
    under storage rounding (the per-layer test above carries the tight ones)."""
    import numpy as np
    from oracle import dla_ref, dla_train_ref, loss_ref
    from rangedet_b200 import synth, train
    B = 2
    P = dla_ref.make_params(seed=0, device="cuda")
    data, coord, _, _ = _inputs(B)
    T = synth.rpn_targets(B, seed=500)
    step = train.GraphedTrainStep({k: v.clone() for k, v in P.items()}, B, H, W, lr=0.0, capture=False, act_dtype=dtype)
    step.set_targets(T)
    step.forward(data, coord)
    step._bwd()
    torch.cuda.synchronize()
    cls_g, reg_g = [c.clone() for c in step.out[0]], [r.clone() for r in step.out[1]]
    # oracle: forward, loss head evaluated at ITS OWN head outputs (IoU target through the C restatement), backward
    ref = dla_train_ref.TrainRef(P, store=dtype)
    cls_r, reg_r = ref.forward(data, coord)
    cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    d_cls, d_reg, losses_r, losses_g = [], [], [], []
    for lvl, s in enumerate((1, 2, 4)):
        # the IoU target of the oracle's own regression output, computed by the GPU decode + IoU kernels (bit-identical to
        # the C restatement: tests/test_gpu_parity.py) -- the CPU loop would take minutes at 340 k x 200 pairs per level
        delta = reg_r[lvl].detach().reshape(B, 8, -1).transpose(1, 2).contiguous()
        dec = ops.decode_3d_bbox(delta, cu(T["pc_vehicle_frame_s%d" % s]))
        iou = ops.batch_rotated_iou(dec, cu(T["gt_bbox_veh_for_iou_pred"]), "bev").reshape(B, 1, H, W // s)
        r = loss_ref.rpn_loss_level(cls_r[lvl], reg_r[lvl], None, None, cu(T["range_image_mask_s%d" % s]), cu(T["rpn_reg_target_s%d" % s]),
                                    cu(T["rpn_reg_weight_s%d" % s]), cu(T["reg_normalize_weight_s%d" % s]), iou_target_override=iou)
        d_cls.append(r["d_cls"])
        d_reg.append(r["d_reg"])
        losses_r += [float(r["cls_loss"].sum()), float(r["reg_loss"].sum())]
        losses_g += [float(step.loss_out[lvl]["cls_loss"].sum()), float(step.loss_out[lvl]["reg_loss"].sum())]
    rnd = lambda t: t.to(dtype).float()
    torch.autograd.backward(cls_r + reg_r, [rnd(g) for g in d_cls + d_reg])
    res = {"head": {}, "loss_ref": losses_r, "loss_gpu": losses_g, "grads": {}}
    for l in range(3):
        res["head"]["cls_%d" % l] = _rms_rel(cls_g[l], cls_r[l])
        res["head"]["reg_%d" % l] = _rms_rel(reg_g[l], reg_r[l])
    for k in step.names:
        g_ref = ref.P[k].grad
        if g_ref is None:
            continue
        got = step.gviews[k]
        res["grads"][k] = {"rms": _rms_rel(got, g_ref), "cos": _cos(got, g_ref)}
    tag = "f16" if dtype == torch.float16 else "bf16"
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "parity_e2e_%s.json" % tag), "w") as f:
        json.dump(res, f, indent=1)
    # Observed (profiles/r02_parity_*.json): f16 head outputs 1.4-2.9e-2 rms, loss sums 1e-4-3e-4, gradient cosines min 0.91 /
    # median 0.97; bf16 9-18e-2, 1e-3, 0.48 / 0.81.  Two correct implementations of a ~40-layer random-init ReLU network
    # that differ by one storage rounding per layer disagree on a fraction of the ReLU masks, hence this much in deep
    # gradients; the TIGHT checks are the teacher-forced tests above (every layer within one rounding, every stage /
    # tower within a few).  A composition error -- a mis-routed skip, a missing
    # gradient accumulation, a stale buffer -- drives the affected cosines to ~0 and the loss sums off by O(1).
    tol_head = {torch.float16: 5e-2, torch.bfloat16: 2.5e-1}[dtype]
    tol_loss = {torch.float16: 1e-3, torch.bfloat16: 5e-3}[dtype]
    cos_min, cos_med = {torch.float16: (0.85, 0.95), torch.bfloat16: (0.35, 0.7)}[dtype]
    assert max(res["head"].values()) < tol_head, res["head"]
    for a, b in zip(losses_g, losses_r):
        assert abs(a - b) <= tol_loss * abs(b) + 1e-6, (losses_g, losses_r)
    cs = sorted(v["cos"] for v in res["grads"].values())
    low = {k: v for k, v in res["grads"].items() if not (v["cos"] >= cos_min)}
    assert len(res["grads"]) >= 270 and not low, sorted(low.items(), key=lambda kv: kv[1]["cos"])[:8]
    assert cs[len(cs) // 2] >= cos_med, cs[len(cs) // 2]


def test_cfg4_forward_b8_at_full_size(ops):
    """BASELINE.json configs[3]: backbone + Meta-Kernel + head forward, bf16, B=8, 64x2656 (inference form: folded moving
    statistics) against oracle/dla_ref.py with the same storage rounding; features of the three levels and the six outputs."""
    from oracle import dla_ref
    from rangedet_b200 import dla
    B = 8
    P = dla_ref.make_params(seed=0, device="cuda")
    data, coord, _, _ = _inputs(B, seed=3)
    bb, hd = dla.DLABackbone(P), dla.RangeRpnHead(P)
    feats = bb.get_rpn_feature(data, coord)
    cls, reg = hd.get_fpn_output(feats)
    torch.cuda.synchronize()
    ref = dla_ref.Ref(P, bf16=True)
    res = {}
    with torch.no_grad():
        feats_r = ref.backbone(data, coord)
        cls_r, reg_r = ref.head(feats_r)
        for l in range(3):
            c = feats_r[l].shape[1]
            res["feat_%d" % l] = _rms_rel(ops.from_nhwc_padded(feats[l], c), feats_r[l])
            res["cls_%d" % l] = _rms_rel(cls[l], cls_r[l])
            res["reg_%d" % l] = _rms_rel(reg[l], reg_r[l])
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "parity_cfg4_b8.json"), "w") as f:
        json.dump(res, f, indent=1)
    assert max(res.values()) < 5e-2, res
