"""Random-init parameters of the RangeDet graph under the reference's names and shapes.

Names / shapes follow the symbols built by rangedet/symbol/backbone/dla_backbone.py:17-161 (stages res1,
res2a, res2, res3a, res3, agg2, agg1, agg2a, agg3; num_block / num_filter from config/rangedet/
rangedet_veh_wo_aug_4_18e.py:89-108), the Meta-Kernel unit ``res1_unit2`` (meta_kernel.py:138,198:
``<unit>_<W>_mlp{i}_weight``) and rangedet/symbol/head/builder.py:198-266.  Initialiser as in
tools/train.py:198: Xavier, gaussian, fan-in, magnitude 2 for weights; gamma 1, beta 0, moving_mean 0,
moving_var 1; biases 0.
"""
import torch

NUM_BLOCK = {'res1': 2, 'res2a': 3, 'res2': 3, 'res3a': 5, 'res3': 5, 'agg1': 2, 'agg2': 2, 'agg2a': 1, 'agg3': 2}
NUM_FILTER = {'res1': 64, 'res2a': 64, 'res2': 128, 'res3a': 128, 'res3': 128, 'agg1': 64, 'agg2': 128, 'agg2a': 64,
              'agg3': 64}
META_UNITS = ('res1_unit2',)


def make_params(seed=0, in_channels=8, device="cuda", head_channels=128, levels_in=(72, 64, 128)):
    g = torch.Generator().manual_seed(seed)
    P = {}

    def weight(name, shape, fan_in):
        P[name + "_weight"] = torch.randn(*shape, generator=g) * (2.0 / fan_in) ** 0.5

    def conv(name, co, ci, kh, kw, bias=False):
        weight(name, (co, ci, kh, kw), ci * kh * kw)
        if bias:
            P[name + "_bias"] = torch.zeros(co)

    def bn(name, c):
        P[name + "_gamma"] = torch.ones(c)
        P[name + "_beta"] = torch.zeros(c)
        P[name + "_moving_mean"] = torch.zeros(c)
        P[name + "_moving_var"] = torch.ones(c)

    def block(name, ci, co, proj):
        if name in META_UNITS:
            conv(name + "_2656_mlp0", 32, 3, 1, 1, bias=True)
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
        key = name.replace("_res", "")
        co = NUM_FILTER[key]
        for i in range(1, NUM_BLOCK[key] + 1):
            block("%s_unit%d" % (name, i), ci if i == 1 else co, co, i == 1)
        return co

    c = in_channels
    for s in ("res1", "res2a", "res2", "res3a", "res3"):
        c = stage(s, c)
    for name, cin, kw in (("agg2", 128, 8), ("agg1", 128, 8), ("agg2a", 128, 4), ("agg3", 64, 4)):
        co = NUM_FILTER[name]
        weight(name + "_deconv", (cin, co, 3, kw), cin * 3 * kw)  # MXNet Deconvolution weight: (Cin, Cout, kh, kw)
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


def num_parameters(P):
    return sum(v.numel() for k, v in P.items() if not k.endswith(("_moving_mean", "_moving_var")))
