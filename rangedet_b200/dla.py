"""DLA aggregation backbone + RPN head forward (inference form) on the sm_100a kernels.

Host-side mirror of rangedet/symbol/backbone/dla_backbone.py:17-175 (DLABackboneBuilder / DLABackbone)
and rangedet/symbol/head/builder.py:198-266 (RangeRpnHead.get_fpn_output): same stages, same unit
and parameter names (``res1_unit1_conv1_weight``, ``res1_unit2aggregation_conv1_weight``,
``agg2_deconv_weight``, ``rpn_cls_conv_0_lvl_0_weight`` ...), so a checkpoint keyed by the reference's
names maps one to one.  Instead of emitting MXNet symbols, each layer runs the tcgen05 implicit-GEMM
kernel (ops.conv2d_nhwc / ops.deconv2d_nhwc) with BatchNorm (moving statistics, eps 1e-5+1e-10,
mxnext/complicate.py:14,32-43) + ReLU + residual folded into its epilogue, and the Meta-Kernel unit
runs the fused TMA/tcgen05 Meta-Kernel.  Activations stay in HBM as zero-haloed NHWC bf16.

The Meta-Kernel unit (meta_kernel_conv) runs as: NHWC bf16 -> NCHW fp32 conversion of the 64-channel
input (torch, the op boundary of the Meta-Kernel) -> ONE fused kernel (Meta-Kernel + BN(576) + ReLU,
NHWC bf16 out) -> tcgen05 1x1 aggregation conv.  `fuse_meta=False` keeps the reference op boundary
((B,576,H,W) fp32) with torch glue, for comparison.  The training graph (batch statistics, backward kernels, loss,
optimiser) is rangedet_b200/train.py.
"""
import torch

from . import ops

EPS = 1e-5 + 1e-10
NUM_BLOCK = {'res1': 2, 'res2a': 3, 'res2': 3, 'res3a': 5, 'res3': 5, 'agg1': 2, 'agg2': 2, 'agg2a': 1, 'agg3': 2}
META_UNITS = ('res1_unit2',)


def _pad64(c):
    return ((c + 63) // 64) * 64


class _Layer:
    """One conv / deconv with folded BN: packed bf16 weight + fp32 scale/shift (padded to 64/128 channels)."""

    def __init__(self, P, wname, bnname=None, bias=None, deconv=False, device="cuda", dtype=torch.bfloat16):
        w = P[wname + "_weight"].to(device)
        if deconv:
            ci, co = w.shape[0], w.shape[1]
        else:
            co, ci = w.shape[0], w.shape[1]
        self.cin, self.cout = ci, co
        self.cin_p, self.cout_p = _pad64(ci), (64 if co <= 64 else 128)
        self.w = (ops.pack_deconv_weight if deconv else ops.pack_conv_weight)(w, self.cin_p, self.cout_p, dtype)
        scale = torch.ones(self.cout_p, device=device)
        shift = torch.zeros(self.cout_p, device=device)
        if bnname is not None:
            s = P[bnname + "_gamma"].to(device) / torch.sqrt(P[bnname + "_moving_var"].to(device) + EPS)
            scale[:co] = s
            shift[:co] = P[bnname + "_beta"].to(device) - P[bnname + "_moving_mean"].to(device) * s
        if bias is not None:
            shift[:co] = P[bias].to(device)
        self.scale, self.shift = scale, shift
        self.deconv = deconv


class _BufferPool:
    """Output activations are reused across forward calls: the kernels only ever write the interior,
    so a haloed buffer zero-initialised once keeps a valid zero halo forever."""

    def __init__(self, dtype=torch.bfloat16):
        self.bufs, self.dtype = {}, dtype

    def get(self, key, shape, device):
        t = self.bufs.get(key)
        if t is None or tuple(t.shape) != tuple(shape):
            t = torch.zeros(shape, device=device, dtype=self.dtype)
            self.bufs[key] = t
        return t


class DLABackbone(object):
    """DLABackbone(pBackbone).get_rpn_feature(data) of the reference, over torch tensors."""

    def __init__(self, params, device="cuda", meta_impl=ops.IMPL_DEFAULT, fuse_meta=True, act_dtype=torch.bfloat16):
        self.P, self.device, self.meta_impl, self.fuse_meta = params, device, meta_impl, fuse_meta
        self.act_dtype = act_dtype      # storage of activations / operands: bf16 (cfg-4) or fp16 (config:35)
        self.L = {}
        self.pool = _BufferPool(act_dtype)

    def layer(self, wname, bnname=None, deconv=False):
        key = wname
        if key not in self.L:
            self.L[key] = _Layer(self.P, wname, bnname, deconv=deconv, device=self.device, dtype=self.act_dtype)
        return self.L[key]

    def conv_bn(self, x, wname, bnname, stride_w=1, relu=True, residual=None):
        l = self.layer(wname, bnname)
        N, Hp, Wp, _ = x.shape
        out = self.pool.get(wname, (N, Hp, (Wp - 2) // stride_w + 2, l.cout_p), x.device)
        return ops.conv2d_nhwc(x, l.w, l.scale, l.shift, relu=relu, residual_pad=residual, stride_w=stride_w, out=out)

    def meta_kernel_conv(self, x, coord, name):  # dla_backbone.py:58-103
        P, dev = self.P, self.device
        feat = ops.from_nhwc_padded(x)  # (B,64,H,W) fp32 -- op boundary of the Meta-Kernel
        bn = name + "point_wise_mlp_bn1"
        s = P[bn + "_gamma"].to(dev) / torch.sqrt(P[bn + "_moving_var"].to(dev) + EPS)
        b = P[bn + "_beta"].to(dev) - P[bn + "_moving_mean"].to(dev) * s
        args = (P[name + "_2656_mlp0_weight"].to(dev).reshape(32, 3), P[name + "_2656_mlp0_bias"].to(dev),
                P[name + "_2656_mlp1_weight"].to(dev).reshape(-1, 32), P[name + "_2656_mlp1_bias"].to(dev))
        B, C, H, W = feat.shape
        wname = name + "aggregation_conv1"
        if self.fuse_meta and C == 64 and W % 4 == 0:
            # Meta-Kernel + BN(576) + ReLU in one kernel, NHWC bf16 out with tap-major channels (k*64+c);
            # the 1x1 aggregation conv's input channels are permuted to match
            m = ops.meta_kernel_forward_nhwc(feat, coord, *args, s, b, relu=True,
                                             out=self.pool.get(name + "_meta", (B, H + 2, W + 2, 9 * C), dev))
            if wname + "#tapmajor" not in self.L:
                Pt = {wname + "_weight": ops.tap_major_weight(P[wname + "_weight"].to(dev), C)}
                Pt.update({k: v for k, v in P.items() if k.startswith(name + "aggregation_bn1")})
                self.L[wname + "#tapmajor"] = _Layer(Pt, wname, name + "aggregation_bn1", device=dev, dtype=self.act_dtype)
            l = self.L[wname + "#tapmajor"]
            return ops.conv2d_nhwc(m, l.w, l.scale, l.shift, relu=True, out=self.pool.get(wname, (B, H + 2, W + 2, l.cout_p), dev))
        m = ops.meta_kernel_forward(feat, coord, *args, impl=self.meta_impl)
        m = torch.relu_(m.mul_(s[None, :, None, None]).add_(b[None, :, None, None]))
        return self.conv_bn(ops.to_nhwc_padded(m, dtype=self.act_dtype), wname, name + "aggregation_bn1")

    def basicblock(self, x, coord, name, stride_w, proj):  # dla_backbone.py:17-56
        if name in META_UNITS:
            r1 = self.meta_kernel_conv(x, coord, name)
        else:
            r1 = self.conv_bn(x, name + "_conv1", name + "_bn1")
        sc = self.conv_bn(x, name + "_sc", name + "_sc_bn", stride_w=stride_w, relu=False) if proj else x
        return self.conv_bn(r1, name + "_conv2", name + "_bn2", stride_w=stride_w, relu=True, residual=sc)

    def res_stage(self, x, coord, name, stride_w):  # :105-114
        x = self.basicblock(x, coord, name + "_unit1", stride_w, True)
        for i in range(2, NUM_BLOCK[name.replace("_res", "")] + 1):
            x = self.basicblock(x, coord, "%s_unit%d" % (name, i), 1, False)
        return x

    def agg_stage(self, name, const, up):  # :116-127
        l = self.layer(name + "_deconv", name + "_deconv_bn", deconv=True)
        y = ops.deconv2d_nhwc(up, l.w, l.scale, l.shift, relu=True, residual_pad=const,
                              out=self.pool.get(name + "_deconv", const.shape, up.device))
        return self.res_stage(y, None, name + "_res", 1)

    def get_rpn_feature(self, data, coord):
        """data (B,8,H,W) fp32, coord (B,3,H,W) fp32 -> [agg3+data (72 of 128 ch), agg2a (64), agg2 (128)]
        as haloed NHWC bf16 (backbone_factory :129-161, fpn_strides (1,2,4), add_data_sc)."""
        x = ops.to_nhwc_padded(data, 64, dtype=self.act_dtype)
        res1 = self.res_stage(x, coord, "res1", 1)
        res2a = self.res_stage(res1, None, "res2a", 2)
        res2 = self.res_stage(res2a, None, "res2", 2)
        res3a = self.res_stage(res2, None, "res3a", 2)
        res3 = self.res_stage(res3a, None, "res3", 2)
        agg2 = self.agg_stage("agg2", res2, res3)
        agg1 = self.agg_stage("agg1", res1, res2)
        agg2a = self.agg_stage("agg2a", res2a, agg2)
        agg3 = self.agg_stage("agg3", agg1, agg2a)
        c = data.shape[1]
        cat = self.pool.get("data_concat", agg3.shape[:3] + (128,), agg3.device)  # concat(data, agg3): 72 of 128 ch
        ops.copy_channels(x, 0, cat, 0, c)
        ops.copy_channels(agg3, 0, cat, c, 64)
        return [cat, agg2a, agg2]


class RangeRpnHead(object):
    """get_fpn_output (builder.py:198-266): per level, un-shared cls / reg towers + 1x1 heads."""

    def __init__(self, params, device="cuda", act_dtype=torch.bfloat16):
        self.P, self.device, self.act_dtype = params, device, act_dtype
        self.L = {}
        self.pool = _BufferPool(act_dtype)

    def _conv(self, x, l, name, relu):
        N, Hp, Wp, _ = x.shape
        return ops.conv2d_nhwc(x, l.w, l.scale, l.shift, relu=relu, out=self.pool.get(name, (N, Hp, Wp, l.cout_p), x.device))

    def _l(self, wname, bnname=None, bias=None):
        if wname not in self.L:
            self.L[wname] = _Layer(self.P, wname, bnname, bias=bias, device=self.device, dtype=self.act_dtype)
        return self.L[wname]

    def get_fpn_output(self, feats):
        cls_logit, bbox_delta = [], []
        for lvl, f in enumerate(feats):
            c = r = f
            for i in range(4):
                n = "rpn_cls_conv_%d_lvl_%d" % (i, lvl)
                l = self._l(n, n + "_bn")
                c = self._conv(c, l, n, True)
                n = "rpn_reg_conv_%d_lvl_%d" % (i, lvl)
                l = self._l(n, n + "_bn")
                r = self._conv(r, l, n, True)
            l = self._l("rpn_cls_logit_lvl_%d" % lvl, None, "rpn_cls_logit_lvl_%d_bias" % lvl)
            cls_logit.append(ops.from_nhwc_padded(self._conv(c, l, "rpn_cls_logit_lvl_%d" % lvl, False), 1))
            l = self._l("rpn_reg_delta_lvl_%d" % lvl, None, "rpn_reg_delta_lvl_%d_bias" % lvl)
            bbox_delta.append(ops.from_nhwc_padded(self._conv(r, l, "rpn_reg_delta_lvl_%d" % lvl, False), 8))
        return cls_logit, bbox_delta


class GraphedForward(object):
    """Whole backbone + head forward captured once in a CUDA graph and replayed (the forward is ~95
    kernel launches of 0.05-0.5 ms: launching them from Python is host-bound).  Inputs are copied into
    static buffers; outputs are the static tensors of the captured run."""

    def __init__(self, params, batch, H, W, device="cuda", act_dtype=torch.bfloat16):
        self.backbone, self.head = DLABackbone(params, device, act_dtype=act_dtype), RangeRpnHead(params, device, act_dtype)
        self.data = torch.zeros((batch, 8, H, W), device=device)
        self.coord = torch.zeros((batch, 3, H, W), device=device)
        for _ in range(2):  # warm-up: builds layers, buffers, function attributes outside the capture
            self.out = self.head.get_fpn_output(self.backbone.get_rpn_feature(self.data, self.coord))
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
            self.out = self.head.get_fpn_output(self.backbone.get_rpn_feature(self.data, self.coord))

    def __call__(self, data, coord):
        self.data.copy_(data, non_blocking=True)
        self.coord.copy_(coord, non_blocking=True)
        self.graph.replay()
        return self.out
