"""Training-mode forward + backward of the DLA backbone + Meta-Kernel unit + RPN head on the sm_100a kernels.

Host-side mirror, for the training graph, of rangedet/symbol/backbone/dla_backbone.py:17-175 and
rangedet/symbol/head/builder.py:198-266 -- same stages, units and parameter names as the reference's
symbols (``res1_unit1_conv1_weight``, ``agg2_deconv_bn_gamma``, ``rpn_cls_conv_0_lvl_0_weight`` ...), so
the gradient dictionary returned by ``backward`` is keyed like the reference's ``arg_params``.  What MXNet's
autograd + cuDNN do there is done here by a small tape over the C-ABI kernels:

  layer forward   conv (tcgen05, raw bf16 output) -> batch statistics -> y = relu(z*a+b + res) (+ res)
  layer backward  g = dy*mask, BN reductions, dz          (bn_train.cu, HBM-bound passes)
                  dW = sum_pixels dz (x) x                  (conv_wgrad.cu, tcgen05, MN-major operands)
                  dx = conv^T(dz, W) (+ other consumers)   (conv_tc.cu: the data gradient of a conv is a
                                                            conv with transposed weights; accumulation
                                                            into an existing gradient rides its epilogue)
  transposed convs (agg stages): dz is written in the phase-grouped view [N][H+2][W_in+2][S*C], in which
                  both gradients are plain 3x3 stride-1 contractions.
  Meta-Kernel unit (dla_backbone.py:58-103): fused Meta-Kernel forward writing raw NHWC bf16 (tap-major
                  channels) -> BN(576)+ReLU -> 1x1 aggregation conv; backward through rd_meta_kernel_bwd
                  at the reference's op boundary ((B,576,H,W) fp32).

BatchNorm uses batch statistics per GPU (mxnext/complicate.py:32-43; config: normalizer = local BN) and
updates the moving statistics like MXNet (momentum 0.9, biased variance).  Activations and activation
gradients are stored as zero-haloed NHWC bf16; parameters and parameter gradients are fp32.
"""
import torch

from . import ops

NUM_BLOCK = {'res1': 2, 'res2a': 3, 'res2': 3, 'res3a': 5, 'res3': 5, 'agg1': 2, 'agg2': 2, 'agg2a': 1, 'agg3': 2}
META_UNITS = ('res1_unit2',)


def _pad64(c):
    return ((c + 63) // 64) * 64


def _cout_pad(c):
    return 64 if c <= 64 else 128


class _Pool:
    """Named activation / gradient buffers reused across steps.  Kernels only ever write the interior, so
    a haloed buffer zero-initialised once keeps a valid zero halo."""

    def __init__(self, device):
        self.device, self.bufs = device, {}

    def get(self, key, shape):
        t = self.bufs.get(key)
        if t is None or tuple(t.shape) != tuple(shape):
            t = torch.zeros(tuple(shape), device=self.device, dtype=torch.bfloat16)
            self.bufs[key] = t
        return t


class TrainGraph(object):
    """forward(data, coord) -> (cls_logit[3], bbox_delta[3]) fp32 NCHW; backward(d_cls, d_reg) -> {name: grad}."""

    def __init__(self, params, device="cuda", use_meta=True):
        self.P = params  # fp32 master parameters (reference names), updated in place by the optimiser
        self.device = device
        self.use_meta = use_meta
        self.pool = _Pool(device)
        self.packed = {}
        self.tape = []
        self.grads = {}
        self.pgrads = {}
        self.nograd = set()
        self._n = 0
        self.debug = None  # set to a dict to capture intermediates (diagnostics)

    # ---- parameter packing (bf16 operand copies; call refresh() after an optimiser step) -------------
    def refresh(self):
        self.packed = {}

    def _w(self, name, kind, ci_p, co_p, S=1):
        key = (name, kind)
        t = self.packed.get(key)
        if t is not None:
            return t
        w = self.P[name + "_weight"]
        if kind == "fwd":          # conv, cross-correlation: [tap][co][ci]
            t = ops.pack_conv_weight(w, ci_p, co_p)
        elif kind == "fwd_tapmajor":  # 1x1 aggregation conv over tap-major Meta-Kernel channels
            t = ops.pack_conv_weight(ops.tap_major_weight(w, w.shape[1] // 9), ci_p, co_p)
        elif kind == "dgrad":      # stride 1: flipped taps, transposed channels: [tap][ci][co]
            t = ops.pack_conv_weight(w.flip(2, 3).transpose(0, 1), co_p, ci_p)
        elif kind == "dgrad_tapmajor":
            t = ops.pack_conv_weight(ops.tap_major_weight(w, w.shape[1] // 9).transpose(0, 1), co_p, ci_p)
        elif kind == "dgrad_s2":   # W-stride 2: transposed conv (3,3)/(1,2), taps not flipped; 1x1 -> centre tap
            if w.shape[2] == 1:
                w3 = torch.zeros((w.shape[0], w.shape[1], 3, 3), device=w.device, dtype=w.dtype)
                w3[:, :, 1, 1] = w[:, :, 0, 0]
                w = w3
            t = ops.pack_deconv_weight(w, co_p, ci_p)  # (Cin_t = co, Cout_t = ci)
        elif kind == "deconv_fwd":
            t = ops.pack_deconv_weight(w, ci_p, co_p)
        elif kind == "deconv_dgrad":  # 3x3 stride-1 conv over the phase-grouped dz: [tap][ci][(ph,co)]
            ci, co, _, KW = w.shape
            pad = KW // 4
            g = torch.zeros((3, 3, ci_p, S, co_p), device=w.device, dtype=torch.float32)
            for tx in range(3):
                for ph in range(S):
                    kx = ph + pad - (1 - tx) * S
                    if 0 <= kx < KW:
                        g[:, tx, :ci, ph, :co] = w[:, :, :, kx].permute(2, 0, 1)
            t = g.reshape(9, ci_p, S * co_p).to(torch.bfloat16).contiguous()
        else:
            raise KeyError(kind)
        self.packed[key] = t
        return t

    # ---- tape helpers ---------------------------------------------------------------------------------
    def _buf(self, tag, shape):
        self._n += 1
        return self.pool.get("%s#%d" % (tag, self._n), shape)

    def _acc(self, t, g):
        """grad(t) += g (g is a fresh haloed NHWC bf16 tensor owned by the tape)."""
        if id(t) in self.nograd:
            return
        old = self.grads.get(id(t))
        self.grads[id(t)] = g if old is None else ops.add_nhwc(old, g, out=self._buf("acc", g.shape))

    def begin(self):
        """Start a new step: empty tape, no gradients."""
        self.tape, self.grads, self.pgrads, self.nograd, self._n = [], {}, {}, set(), 0

    def seed_grad(self, t, g):
        """Set the gradient of activation t (haloed NHWC bf16) before run_tape()."""
        self.grads[id(t)] = g

    def grad_of(self, t):
        return self.grads[id(t)]

    def run_tape(self):
        for fn in reversed(self.tape):
            fn()
        self.tape = []

    def _pg(self, name, g):
        self.pgrads[name] = g if name not in self.pgrads else self.pgrads[name] + g

    # ---- layers -----------------------------------------------------------------------------------------
    def conv_bn(self, x, wname, bnname, stride_w=1, relu=True, res_before=None, kinds=("fwd", "dgrad")):
        P = self.P
        w = P[wname + "_weight"]
        co, ci, k = w.shape[0], w.shape[1], w.shape[2]
        ci_p, co_p = x.shape[3], _cout_pad(co)
        N, Hp, Wp, _ = x.shape
        W_out = (Wp - 2) // stride_w
        z = ops.conv2d_nhwc(x, self._w(wname, kinds[0], ci_p, co_p), relu=False, stride_w=stride_w,
                            out=self._buf("z", (N, Hp, W_out + 2, co_p)))
        coef = ops.bn_train_stats(z, P[bnname + "_gamma"], P[bnname + "_beta"], P[bnname + "_moving_mean"],
                                  P[bnname + "_moving_var"])
        y = ops.bn_act_fwd(z, coef, relu=relu, res_before=res_before, out=self._buf("y", z.shape))

        def bwd():
            dy = self.grads.pop(id(y))
            dz, dgamma, dbeta, g = ops.bn_act_bwd(dy, z, coef, 1 if relu else 0, y_mask=y, dz_out=self._buf("dz", z.shape),
                                                  want_g=res_before is not None and relu,
                                                  g_out=self._buf("g", z.shape) if (res_before is not None and relu) else None)
            if res_before is not None:
                self._acc(res_before, g if relu else dy)
            self._pg(bnname + "_gamma", dgamma[:co])
            self._pg(bnname + "_beta", dbeta[:co])
            G = ops.conv2d_wgrad(dz, x, k, stride_w)  # [tap][co_p][ci_p]
            if kinds[0] == "fwd_tapmajor":
                C = ci // 9
                gw = G[0, :co, :ci].reshape(co, 9, C).transpose(1, 2).reshape(co, ci, 1, 1)
            else:
                gw = G[:, :co, :ci].reshape(k, k, co, ci).permute(2, 3, 0, 1)
            self._pg(wname + "_weight", gw.contiguous())
            if id(x) in self.nograd:
                return
            old = self.grads.pop(id(x), None)
            if ci_p > 128:
                # wide data gradient (576 channels of the Meta-Kernel unit): 128/64-channel output slices
                assert stride_w == 1 and old is None
                wt = self._w(wname, kinds[1], ci_p, co_p)
                dx = self._buf("dx", x.shape)
                c0 = 0
                while c0 < ci_p:
                    cs = 128 if ci_p - c0 >= 128 else 64
                    ops.conv2d_nhwc_slice(dz, wt[:, c0:c0 + cs].contiguous(), dx, c0)
                    c0 += cs
            elif stride_w == 1:
                dx = ops.conv2d_nhwc(dz, self._w(wname, kinds[1], ci_p, co_p), relu=False, residual_pad=old,
                                     out=self._buf("dx", x.shape))
            else:
                dx = ops.deconv2d_nhwc(dz, self._w(wname, "dgrad_s2", ci_p, co_p), relu=False, residual_pad=old,
                                       out=self._buf("dx", x.shape))
            self.grads[id(x)] = dx
            if self.debug is not None:
                self.debug[wname] = dict(dy=dy, dz=dz, x=x, z=z, y=y, G=G, dx=dx)

        self.tape.append(bwd)
        return y

    def deconv_bn(self, up, const, name):
        """agg_stage head (dla_backbone.py:116-124): const + relu(bn(deconv(up)))."""
        P = self.P
        wname, bnname = name + "_deconv", name + "_deconv_bn"
        w = P[wname + "_weight"]
        ci, co, _, KW = w.shape
        S = KW // 2
        ci_p, co_p = up.shape[3], _cout_pad(co)
        N, Hp, Wp, _ = up.shape
        W_in = Wp - 2
        z = ops.deconv2d_nhwc(up, self._w(wname, "deconv_fwd", ci_p, co_p), relu=False,
                              out=self._buf("z", (N, Hp, W_in * S + 2, co_p)))
        coef = ops.bn_train_stats(z, P[bnname + "_gamma"], P[bnname + "_beta"], P[bnname + "_moving_mean"],
                                  P[bnname + "_moving_var"])
        y = ops.bn_act_fwd(z, coef, relu=True, res_after=const, out=self._buf("y", z.shape))

        def bwd():
            dy = self.grads.pop(id(y))
            dzg, dgamma, dbeta, _ = ops.bn_act_bwd(dy, z, coef, 2, dz_halo_w=S,
                                                   dz_out=self._buf("dzg", (N, Hp, W_in * S + 2 * S, co_p)))
            self._acc(const, dy)
            self._pg(bnname + "_gamma", dgamma[:co])
            self._pg(bnname + "_beta", dbeta[:co])
            dzg = dzg.view(N, Hp, W_in + 2, S * co_p)  # phase-grouped: pixel group j holds output pixels j*S .. j*S+S-1
            G = ops.conv2d_wgrad(up, dzg, 3, 1).reshape(3, 3, ci_p, S, co_p)  # [ty][tx][ci][ph][co]
            gw = torch.zeros_like(w)
            pad = KW // 4
            for kx in range(KW):
                tx, ph = divmod(kx - pad + S, S)
                gw[:, :, :, kx] = G[:, tx, :ci, ph, :co].permute(1, 2, 0)
            self._pg(wname + "_weight", gw)
            old = self.grads.pop(id(up), None)
            self.grads[id(up)] = ops.conv2d_nhwc(dzg, self._w(wname, "deconv_dgrad", ci_p, co_p, S), relu=False,
                                                 residual_pad=old, out=self._buf("dx", up.shape))

        self.tape.append(bwd)
        return y

    def head_out(self, x, wname, co):
        """1x1 conv + bias, no norm (builder.py:247-262) -> fp32 NCHW."""
        P = self.P
        ci_p = x.shape[3]
        bias = torch.zeros(64, device=self.device)
        bias[:co] = P[wname + "_bias"]
        zp = ops.conv2d_nhwc(x, self._w(wname, "fwd", ci_p, 64), None, bias, relu=False, out=self._buf("z", x.shape[:3] + (64,)))
        out = ops.nhwc_to_nchw(zp, co)

        def bwd(d_out):
            dz = ops.nchw_to_nhwc(d_out.contiguous(), self._buf("dz", zp.shape))
            self._pg(wname + "_bias", ops.channel_sums(dz)[:co])
            G = ops.conv2d_wgrad(dz, x, 1, 1)
            self._pg(wname + "_weight", G[0, :co, :P[wname + "_weight"].shape[1]].reshape(P[wname + "_weight"].shape).contiguous())
            old = self.grads.pop(id(x), None)
            self.grads[id(x)] = ops.conv2d_nhwc(dz, self._w(wname, "dgrad", ci_p, 64), relu=False, residual_pad=old,
                                                out=self._buf("dx", x.shape))

        return out, bwd

    def meta_kernel_conv(self, x, coord, name):
        """dla_backbone.py:58-103 in training mode."""
        a = self.meta_kernel_front(x, coord, name)
        return self.conv_bn(a, name + "aggregation_conv1", name + "aggregation_bn1",
                            kinds=("fwd_tapmajor", "dgrad_tapmajor"))

    def meta_kernel_front(self, x, coord, name):
        """Meta-Kernel -> BN(576) -> ReLU (dla_backbone.py:79-94), NHWC bf16 with tap-major channels."""
        P = self.P
        C = 64
        B, Hp, Wp, _ = x.shape
        H, W = Hp - 2, Wp - 2
        feat = ops.nhwc_to_nchw(x)  # (B,64,H,W) fp32: the op boundary of the Meta-Kernel
        mlp = (P[name + "_2656_mlp0_weight"].reshape(32, 3), P[name + "_2656_mlp0_bias"],
               P[name + "_2656_mlp1_weight"].reshape(-1, 32), P[name + "_2656_mlp1_bias"])
        one, zero = torch.ones(9 * C, device=self.device), torch.zeros(9 * C, device=self.device)
        # raw Meta-Kernel output, NHWC bf16, tap-major channels k*C+c (no (B,576,H,W) fp32 intermediate)
        m = ops.meta_kernel_forward_nhwc(feat, coord, *mlp, one, zero, relu=False, out=self._buf("meta", (B, Hp, Wp, 9 * C)))
        bn = name + "point_wise_mlp_bn1"
        tm = lambda v: v.reshape(C, 9).t().reshape(-1).contiguous()      # reference order c*9+k -> tap-major
        untm = lambda v: v.reshape(9, C).t().reshape(-1).contiguous()
        gamma_t, beta_t = tm(P[bn + "_gamma"]), tm(P[bn + "_beta"])
        mm_t, mv_t = tm(P[bn + "_moving_mean"]), tm(P[bn + "_moving_var"])
        coef = ops.bn_train_stats(m, gamma_t, beta_t, mm_t, mv_t)
        P[bn + "_moving_mean"].copy_(untm(mm_t))
        P[bn + "_moving_var"].copy_(untm(mv_t))
        a = ops.bn_act_fwd(m, coef, relu=True, out=self._buf("meta_act", m.shape))

        def bwd():
            da = self.grads.pop(id(a))
            dm, dgamma, dbeta, _ = ops.bn_act_bwd(da, m, coef, 1, y_mask=a, dz_out=self._buf("dmeta", m.shape))
            self._pg(bn + "_gamma", untm(dgamma))
            self._pg(bn + "_beta", untm(dbeta))
            # back to the reference op boundary: grad_out (B, 576 = c*9+k, H, W) fp32
            go = ops.nhwc_to_nchw(dm, tap_major=True)
            gd, gw0, gb0, gw1, gb1 = ops.meta_kernel_backward(go, feat, coord, *mlp)
            if self.debug is not None:
                self.debug.update(meta_da=da, meta_dm=dm, meta_go=go, meta_gd=gd, meta_m=m, meta_a=a)
            self._pg(name + "_2656_mlp0_weight", gw0.reshape(P[name + "_2656_mlp0_weight"].shape))
            self._pg(name + "_2656_mlp0_bias", gb0)
            self._pg(name + "_2656_mlp1_weight", gw1.reshape(P[name + "_2656_mlp1_weight"].shape))
            self._pg(name + "_2656_mlp1_bias", gb1)
            self._acc(x, ops.nchw_to_nhwc(gd, self._buf("dx", x.shape)))

        self.tape.append(bwd)
        return a

    def basicblock(self, x, coord, name, stride_w, proj):  # dla_backbone.py:17-56
        if self.use_meta and name in META_UNITS:
            r1 = self.meta_kernel_conv(x, coord, name)
        else:
            r1 = self.conv_bn(x, name + "_conv1", name + "_bn1")
        sc = self.conv_bn(x, name + "_sc", name + "_sc_bn", stride_w=stride_w, relu=False) if proj else x
        return self.conv_bn(r1, name + "_conv2", name + "_bn2", stride_w=stride_w, relu=True, res_before=sc)

    def res_stage(self, x, coord, name, stride_w):
        x = self.basicblock(x, coord, name + "_unit1", stride_w, True)
        for i in range(2, NUM_BLOCK[name.replace("_res", "")] + 1):
            x = self.basicblock(x, coord, "%s_unit%d" % (name, i), 1, False)
        return x

    def agg_stage(self, name, const, up):
        return self.res_stage(self.deconv_bn(up, const, name), None, name + "_res", 1)

    # ---- whole graph --------------------------------------------------------------------------------------
    def forward(self, data, coord):
        self.begin()
        c = data.shape[1]
        x = ops.nchw_to_nhwc(data.contiguous(), self._buf("data", (data.shape[0], data.shape[2] + 2, data.shape[3] + 2, 64)))
        self.nograd.add(id(x))
        res1 = self.res_stage(x, coord, "res1", 1)
        res2a = self.res_stage(res1, None, "res2a", 2)
        res2 = self.res_stage(res2a, None, "res2", 2)
        res3a = self.res_stage(res2, None, "res3a", 2)
        res3 = self.res_stage(res3a, None, "res3", 2)
        agg2 = self.agg_stage("agg2", res2, res3)
        agg1 = self.agg_stage("agg1", res1, res2)
        agg2a = self.agg_stage("agg2a", res2a, agg2)
        agg3 = self.agg_stage("agg3", agg1, agg2a)
        cat = self._buf("data_concat", agg3.shape[:3] + (128,))  # concat(data, agg3): 72 of 128 channels
        cat[..., :c] = x[..., :c]
        cat[..., c:c + 64] = agg3

        def cat_bwd():
            dcat = self.grads.pop(id(cat))
            self._acc(agg3, dcat[..., c:c + 64].contiguous())

        self.tape.append(cat_bwd)
        self.head_bwd = []
        cls_logit, bbox_delta = [], []
        for lvl, f in enumerate([cat, agg2a, agg2]):
            t_c = t_r = f
            for i in range(4):
                n = "rpn_cls_conv_%d_lvl_%d" % (i, lvl)
                t_c = self.conv_bn(t_c, n, n + "_bn")
                n = "rpn_reg_conv_%d_lvl_%d" % (i, lvl)
                t_r = self.conv_bn(t_r, n, n + "_bn")
            o, b = self.head_out(t_c, "rpn_cls_logit_lvl_%d" % lvl, 1)
            cls_logit.append(o)
            self.head_bwd.append(("cls", lvl, b, len(self.tape)))
            o, b = self.head_out(t_r, "rpn_reg_delta_lvl_%d" % lvl, 8)
            bbox_delta.append(o)
            self.head_bwd.append(("reg", lvl, b, len(self.tape)))
        return cls_logit, bbox_delta

    def backward(self, d_cls, d_reg):
        """d_cls[l] (B,1,H,W_l), d_reg[l] (B,8,H,W_l) fp32: gradients of the loss w.r.t. the head outputs.
        Returns {parameter name: fp32 gradient in the reference's shape}."""
        for kind, lvl, b, _ in self.head_bwd:
            b(d_cls[lvl] if kind == "cls" else d_reg[lvl])
        self.run_tape()
        return self.pgrads


def sgd_momentum_step(params, grads, momenta, lr, momentum=0.9, wd=1e-4, clip_gradient=None, rescale_grad=1.0):
    """MXNet SGD (tools/train.py:312-319): g = clip(rescale*g) + wd*w; m = momentum*m - lr*g; w += m.
    Fused multi-tensor update through torch._foreach (plumbing: 9.1 M parameters, 0.1 % of the step)."""
    names = [n for n in grads if n in params]
    ws = [params[n] for n in names]
    gs = [grads[n].to(params[n].dtype).reshape(params[n].shape) * rescale_grad for n in names]
    if clip_gradient is not None:
        gs = [g.clamp_(-clip_gradient, clip_gradient) for g in gs]
    ms = []
    for n in names:
        if n not in momenta:
            momenta[n] = torch.zeros_like(params[n])
        ms.append(momenta[n])
    torch._foreach_add_(gs, ws, alpha=wd)
    torch._foreach_mul_(ms, momentum)
    torch._foreach_add_(ms, gs, alpha=-lr)
    torch._foreach_add_(ws, ms)


class GraphedTrainStep(object):
    """Forward, backward and the SGD update captured ONCE in CUDA graphs and replayed: a training step is
    ~900 launches of 10-500 us kernels, launching them one by one from Python is host-bound.

        step = GraphedTrainStep(params, batch, H, W, lr=...)
        cls, reg = step.forward(data, coord)          # graph 1
        ... loss on cls / reg -> d_cls, d_reg ...
        step.backward_update(d_cls, d_reg)            # graph 2 (+ data-parallel all-reduce) + graph 3

    All activations / gradients live in the TrainGraph buffer pool, so replays touch static addresses; the
    bf16 operand copies of the weights are re-packed inside graph 1 from the fp32 masters that graph 3
    updates in place.  With `world_size > 1` the parameter gradients are averaged across ranks between
    backward and update by one flat NCCL all-reduce (the reference: hvd.DistributedOptimizer,
    tools/train.py:364-368)."""

    def __init__(self, params, batch, H, W, lr, momentum=0.9, wd=1e-4, clip_gradient=None, device="cuda", use_meta=True,
                 allreduce=None):
        self.tg = TrainGraph(params, device, use_meta)
        self.P = params
        self.allreduce = allreduce
        self.hyper = dict(lr=lr, momentum=momentum, wd=wd, clip_gradient=clip_gradient)
        self.data = torch.zeros((batch, 8, H, W), device=device)
        self.coord = torch.zeros((batch, 3, H, W), device=device)
        self.d_cls = [torch.zeros((batch, 1, H, W >> l), device=device) for l in range(3)]
        self.d_reg = [torch.zeros((batch, 8, H, W >> l), device=device) for l in range(3)]
        self.mom = {}
        names = sorted(k for k in params if not k.endswith(("_moving_mean", "_moving_var")))
        self.names = names
        sizes = [params[k].numel() for k in names]
        self.flat = torch.zeros(sum(sizes), device=device)
        self.gviews, o = {}, 0
        for k, n in zip(names, sizes):
            self.gviews[k] = self.flat[o:o + n].view(params[k].shape)
            o += n
        # warm-up outside the capture (buffer pool, workspaces, function attributes); lr 0 leaves the weights alone
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(2):
                self._fwd()
                self._bwd()
                self._update(lr=0.0)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self.pool = torch.cuda.graph_pool_handle()
        self.g_fwd, self.g_bwd, self.g_upd = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.g_fwd, pool=self.pool):
            self._fwd()
        with torch.cuda.graph(self.g_bwd, pool=self.pool):
            self._bwd()
        with torch.cuda.graph(self.g_upd, pool=self.pool):
            self._update(lr=lr)

    def _fwd(self):
        self.tg.refresh()
        self.out = self.tg.forward(self.data, self.coord)

    def _bwd(self):
        grads = self.tg.backward(self.d_cls, self.d_reg)
        torch._foreach_copy_([self.gviews[k] for k in self.names], [grads[k].reshape(self.P[k].shape) for k in self.names])

    def _update(self, lr):
        h = self.hyper
        sgd_momentum_step(self.P, self.gviews, self.mom, lr=lr, momentum=h["momentum"], wd=h["wd"],
                          clip_gradient=h["clip_gradient"])

    def forward(self, data, coord):
        self.data.copy_(data, non_blocking=True)
        self.coord.copy_(coord, non_blocking=True)
        self.g_fwd.replay()
        return self.out

    def backward_update(self, d_cls, d_reg):
        for dst, src in zip(self.d_cls + self.d_reg, list(d_cls) + list(d_reg)):
            dst.copy_(src, non_blocking=True)
        self.g_bwd.replay()
        if self.allreduce is not None:
            self.allreduce(self.flat)
        self.g_upd.replay()
