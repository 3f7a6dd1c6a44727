"""CPU/GPU-agnostic torch restatement of the DLA backbone + RPN head forward (inference form).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  "Parity unpinned": every op below is an MXNet
library operator in the reference (mx.sym.Convolution / Deconvolution / BatchNorm / Activation via
mxnext/simple.py:123-158,545-580 and mxnext/complicate.py:32-43); this file follows the graph
builders op for op and name for name:

  rangedet/symbol/backbone/dla_backbone.py
    :17-56    basicblock   conv3x3-BN-ReLU (or the Meta-Kernel unit :58-103), conv3x3(stride)-BN,
                           projection shortcut conv1x1(stride)-BN when `proj`, add, ReLU
    :105-114  res_stage    unit1 (proj=True, stride s) + units 2..n
    :116-127  agg_stage    deconv-BN-ReLU, + skip, res_stage
    :129-161  backbone_factory   res1, res2a, res2, res3a, res3, agg2, agg1, agg2a, agg3, concat(data)
  rangedet/symbol/head/builder.py:198-266   get_fpn_output (4 x conv3x3-BN-ReLU per branch, 1x1 heads)

BatchNorm is evaluated with its moving statistics (use_global_stats at test time), eps = 1e-5 + 1e-10
(mxnext/complicate.py:14).  `bf16=True` rounds activations and weights to bf16 at the points where the
B200 pipeline stores them, so that the comparison isolates kernel errors from format errors.
"""
import torch
import torch.nn.functional as F

from . import meta_kernel_ref

EPS = 1e-5 + 1e-10
NUM_BLOCK = {'res1': 2, 'res2a': 3, 'res2': 3, 'res3a': 5, 'res3': 5, 'agg1': 2, 'agg2': 2, 'agg2a': 1, 'agg3': 2}
NUM_FILTER = {'res1': 64, 'res2a': 64, 'res2': 128, 'res3a': 128, 'res3': 128, 'agg1': 64, 'agg2': 128, 'agg2a': 64,
              'agg3': 64}
META_UNITS = ('res1_unit2',)


def make_params(seed=0, in_channels=8, device="cpu", head_channels=128, levels_in=(72, 64, 128)):
    """Random parameters under the reference's names (config/rangedet/*.py:89-135)."""
    g = torch.Generator().manual_seed(seed)
    P = {}

    def conv(name, co, ci, kh, kw, bias=False):
        P[name + "_weight"] = torch.randn(co, ci, kh, kw, generator=g) * (2.0 / (ci * kh * kw)) ** 0.5
        if bias:
            P[name + "_bias"] = torch.randn(co, generator=g) * 0.1

    def bn(name, c):
        P[name + "_gamma"] = torch.rand(c, generator=g) * 0.5 + 0.75
        P[name + "_beta"] = torch.randn(c, generator=g) * 0.1
        P[name + "_moving_mean"] = torch.randn(c, generator=g) * 0.1
        P[name + "_moving_var"] = torch.rand(c, generator=g) * 0.5 + 0.75

    def block(name, ci, co, proj):
        if name in META_UNITS:
            conv(name + "_2656_mlp0", 32, 3, 1, 1, bias=True)      # "<name>_<W>_mlp{i}" (meta_kernel.py:138,198)
            conv(name + "_2656_mlp1", co, 32, 1, 1, bias=True)
            bn(name + "point_wise_mlp_bn1", 9 * co)
            conv(name + "aggregation_conv1", co, 9 * co, 1, 1)
            bn(name + "aggregation_bn1", co)
        else:
            conv(name + "_conv1", co, ci, 3, 3)
            bn(name + "_bn1", co)
        conv(name + "_conv2", co, co, 3, 3)
        bn(name + "_bn2", co)
        if proj:
            conv(name + "_sc", co, ci, 1, 1)
            bn(name + "_sc_bn", co)

    def stage(name, ci):
        co = NUM_FILTER[name.replace("_res", "")]
        for i in range(1, NUM_BLOCK[name.replace("_res", "")] + 1):
            block("%s_unit%d" % (name, i), ci if i == 1 else co, co, i == 1)
        return co

    c = stage("res1", in_channels)
    c = stage("res2a", c)
    c = stage("res2", c)
    c = stage("res3a", c)
    c = stage("res3", c)
    for name, cin, kw in (("agg2", 128, 8), ("agg1", 128, 8), ("agg2a", 128, 4), ("agg3", 64, 4)):
        co = NUM_FILTER[name]
        P[name + "_deconv_weight"] = torch.randn(cin, co, 3, kw, generator=g) * (2.0 / (cin * 6)) ** 0.5
        bn(name + "_deconv_bn", co)
        stage(name + "_res", co)
    for lvl, cin in enumerate(levels_in):
        for br in ("cls", "reg"):
            ci = cin
            for i in range(4):
                conv("rpn_%s_conv_%d_lvl_%d" % (br, i, lvl), head_channels, ci, 3, 3)
                bn("rpn_%s_conv_%d_lvl_%d_bn" % (br, i, lvl), head_channels)
                ci = head_channels
        conv("rpn_cls_logit_lvl_%d" % lvl, 1, head_channels, 1, 1, bias=True)
        conv("rpn_reg_delta_lvl_%d" % lvl, 8, head_channels, 1, 1, bias=True)
    return {k: v.to(device) for k, v in P.items()}


def bn_fold(P, name):
    scale = P[name + "_gamma"] / torch.sqrt(P[name + "_moving_var"] + EPS)
    return scale, P[name + "_beta"] - P[name + "_moving_mean"] * scale


class Ref:
    def __init__(self, P, bf16=True, store=None):
        self.P = P
        self.store = store if store is not None else (torch.bfloat16 if bf16 else None)
        self.bf16 = self.store is not None

    def r(self, x):  # storage format of the B200 pipeline (bf16, or fp16 as the reference trains: config:35)
        return x.to(self.store).to(x.dtype) if self.bf16 else x

    def conv_bn(self, x, wname, bnname, stride=(1, 1), relu=True, residual=None, res_after_relu=False):
        w = self.r(self.P[wname + "_weight"])
        y = F.conv2d(x, w, stride=stride, padding=w.shape[-1] // 2)
        s, b = bn_fold(self.P, bnname)
        y = y * s[None, :, None, None] + b[None, :, None, None]
        if residual is not None and not res_after_relu:
            y = y + residual
        if relu:
            y = y.relu()
        return self.r(y)

    def basicblock(self, x, coord, name, stride, proj):
        if name in META_UNITS:  # dla_backbone.py:58-103
            P = self.P
            m = meta_kernel_ref.meta_baseline_bias(x, coord, P[name + "_2656_mlp0_weight"].reshape(32, 3),
                                                   P[name + "_2656_mlp0_bias"], P[name + "_2656_mlp1_weight"].reshape(-1, 32),
                                                   P[name + "_2656_mlp1_bias"])
            s, b = bn_fold(P, name + "point_wise_mlp_bn1")
            m = self.r((m * s[None, :, None, None] + b[None, :, None, None]).relu())
            r1 = self.conv_bn(m, name + "aggregation_conv1", name + "aggregation_bn1")
        else:
            r1 = self.conv_bn(x, name + "_conv1", name + "_bn1")
        sc = self.conv_bn(x, name + "_sc", name + "_sc_bn", stride=stride, relu=False) if proj else x
        return self.conv_bn(r1, name + "_conv2", name + "_bn2", stride=stride, relu=True, residual=sc)

    def res_stage(self, x, coord, name, stride):
        x = self.basicblock(x, coord, name + "_unit1", stride, True)
        for i in range(2, NUM_BLOCK[name.replace("_res", "")] + 1):
            x = self.basicblock(x, coord, "%s_unit%d" % (name, i), (1, 1), False)
        return x

    def agg_stage(self, name, const, up, sw, pad):
        w = self.r(self.P[name + "_deconv_weight"])
        y = F.conv_transpose2d(up, w, stride=(1, sw), padding=(1, pad))
        s, b = bn_fold(self.P, name + "_deconv_bn")
        y = self.r((y * s[None, :, None, None] + b[None, :, None, None]).relu() + const)
        return self.res_stage(y, None, name + "_res", (1, 1))

    def backbone(self, data, coord):
        data = self.r(data)
        res1 = self.res_stage(data, coord, "res1", (1, 1))
        res2a = self.res_stage(res1, None, "res2a", (1, 2))
        res2 = self.res_stage(res2a, None, "res2", (1, 2))
        res3a = self.res_stage(res2, None, "res3a", (1, 2))
        res3 = self.res_stage(res3a, None, "res3", (1, 2))
        agg2 = self.agg_stage("agg2", res2, res3, 4, 2)
        agg1 = self.agg_stage("agg1", res1, res2, 4, 2)
        agg2a = self.agg_stage("agg2a", res2a, agg2, 2, 1)
        agg3 = self.agg_stage("agg3", agg1, agg2a, 2, 1)
        return [torch.cat([data, agg3], 1), agg2a, agg2]  # add_data_sc, fpn_strides (1,2,4)

    def head(self, feats):
        cls, reg = [], []
        for lvl, f in enumerate(feats):
            c = r = f
            for i in range(4):
                c = self.conv_bn(c, "rpn_cls_conv_%d_lvl_%d" % (i, lvl), "rpn_cls_conv_%d_lvl_%d_bn" % (i, lvl))
                r = self.conv_bn(r, "rpn_reg_conv_%d_lvl_%d" % (i, lvl), "rpn_reg_conv_%d_lvl_%d_bn" % (i, lvl))
            P = self.P
            cls.append(F.conv2d(c, self.r(P["rpn_cls_logit_lvl_%d_weight" % lvl]), P["rpn_cls_logit_lvl_%d_bias" % lvl]))
            reg.append(F.conv2d(r, self.r(P["rpn_reg_delta_lvl_%d_weight" % lvl]), P["rpn_reg_delta_lvl_%d_bias" % lvl]))
        return cls, reg
