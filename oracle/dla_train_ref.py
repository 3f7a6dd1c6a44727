"""Torch-autograd restatement of the TRAINING graph of the DLA backbone + Meta-Kernel unit + RPN head.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  "Parity unpinned": in the reference every op below
is an MXNet library operator and the backward pass is MXNet's autograd (mxnet==2.0.0, requirements.txt:2,
absent here); this file follows the graph builders op for op, name for name, in training mode:

  rangedet/symbol/backbone/dla_backbone.py:17-161   basicblock / meta_kernel_conv / res_stage / agg_stage /
                                                     backbone_factory
  rangedet/symbol/head/builder.py:198-266           get_fpn_output
  mxnext/complicate.py:14,32-43                     BatchNorm: batch statistics (use_global_stats False),
                                                     biased variance, eps 1e-5+1e-10, momentum 0.9

`bf16=True` (or `store=torch.float16`) rounds values to the storage type wherever the B200 pipeline stores them (conv outputs z,
activations y, activation gradients, weight operands), with a straight-through gradient that is itself
rounded -- so the comparison isolates kernel errors from the storage format.  Gradients come from
torch.autograd.
"""
import torch
import torch.nn.functional as F

from . import meta_kernel_ref
from .dla_ref import META_UNITS, NUM_BLOCK

EPS = 1e-5 + 1e-10


class _Round(torch.autograd.Function):
    """Round to the storage type in both directions (straight-through: the gradient is what the kernels store)."""

    @staticmethod
    def forward(ctx, x, dtype):
        ctx.dtype = dtype
        return x.to(dtype).to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        return g.to(ctx.dtype).to(g.dtype), None


class TrainRef:
    """jitter > 0 multiplies every value by (1 + jitter*N(0,1)) just before it is rounded to bf16: the
    same graph under rounding-level perturbations.  A ReLU network's gradient is discontinuous where a
    pre-activation crosses zero, so two correct implementations whose conv outputs differ in the last
    fp32 bits disagree on a few masks per layer and hence visibly in deep gradients; the difference
    between the jittered and the plain reference measures that sensitivity (the tests' noise floor)."""

    def __init__(self, P, bf16=True, use_meta=True, jitter=0.0, seed=1234, store=None):
        """store: storage type to emulate (torch.bfloat16 / torch.float16 -- the reference itself trains in fp16,
        config:35); default bf16 when `bf16` is set, none (plain fp32 graph) otherwise."""
        self.P = {k: v.detach().clone().requires_grad_(not k.endswith(("_moving_mean", "_moving_var"))) for k, v in P.items()}
        self.store = store if store is not None else (torch.bfloat16 if bf16 else None)
        bf16 = self.store is not None
        self.bf16, self.use_meta, self.jitter = bf16, use_meta, jitter
        self.gen = None
        if jitter:
            dev = next(iter(P.values())).device
            self.gen = torch.Generator(device=dev).manual_seed(seed)

    def r(self, x, scale=1.0):
        if not self.bf16:
            return x
        if self.jitter:
            x = x * (1.0 + self.jitter * scale * torch.randn(x.shape, device=x.device, generator=self.gen))
        return _Round.apply(x, self.store)

    def bn(self, z, name):
        return F.batch_norm(z, None, None, self.P[name + "_gamma"], self.P[name + "_beta"], training=True, eps=EPS)

    def conv_bn(self, x, wname, bnname, stride=(1, 1), relu=True, residual=None):
        w = self.r(self.P[wname + "_weight"])
        z = self.r(F.conv2d(x, w, stride=stride, padding=w.shape[-1] // 2))
        y = self.bn(z, bnname)
        if residual is not None:
            y = y + residual
        if relu:
            y = y.relu()
        return self.r(y)

    def basicblock(self, x, coord, name, stride, proj):
        if self.use_meta and name in META_UNITS:  # dla_backbone.py:58-103
            P = self.P
            m = meta_kernel_ref.meta_baseline_bias(x, coord, P[name + "_2656_mlp0_weight"].reshape(32, 3),
                                                   P[name + "_2656_mlp0_bias"], P[name + "_2656_mlp1_weight"].reshape(-1, 32),
                                                   P[name + "_2656_mlp1_bias"])
            m = self.r(m, 10.0)  # the tcgen05 Meta-Kernel carries ~2^-16 relative error (split-bf16 products)
            m = self.r(self.bn(m, name + "point_wise_mlp_bn1").relu())
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
        z = self.r(F.conv_transpose2d(up, w, stride=(1, sw), padding=(1, pad)))
        y = self.r(self.bn(z, name + "_deconv_bn").relu() + const)
        return self.res_stage(y, None, name + "_res", (1, 1))

    def forward(self, data, coord):
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
        feats = [torch.cat([data, agg3], 1), agg2a, agg2]
        cls, reg = [], []
        P = self.P
        for lvl, f in enumerate(feats):
            c = r = f
            for i in range(4):
                c = self.conv_bn(c, "rpn_cls_conv_%d_lvl_%d" % (i, lvl), "rpn_cls_conv_%d_lvl_%d_bn" % (i, lvl))
                r = self.conv_bn(r, "rpn_reg_conv_%d_lvl_%d" % (i, lvl), "rpn_reg_conv_%d_lvl_%d_bn" % (i, lvl))
            cls.append(F.conv2d(c, self.r(P["rpn_cls_logit_lvl_%d_weight" % lvl]), P["rpn_cls_logit_lvl_%d_bias" % lvl]))
            reg.append(F.conv2d(r, self.r(P["rpn_reg_delta_lvl_%d_weight" % lvl]), P["rpn_reg_delta_lvl_%d_bias" % lvl]))
        return cls, reg

    def forward_backward(self, data, coord, d_cls, d_reg):
        """-> (cls, reg, {name: grad}) for the linear loss sum(cls*d_cls) + sum(reg*d_reg)."""
        cls, reg = self.forward(data, coord)
        rnd = (lambda g: g.to(self.store).to(g.dtype)) if self.bf16 else (lambda g: g)
        loss = sum((c * rnd(g)).sum() for c, g in zip(cls, d_cls)) + sum((r * rnd(g)).sum() for r, g in zip(reg, d_reg))
        loss.backward()
        grads = {k: v.grad for k, v in self.P.items() if v.requires_grad and v.grad is not None}
        return [c.detach() for c in cls], [r.detach() for r in reg], grads
