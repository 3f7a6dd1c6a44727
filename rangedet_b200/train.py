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


def pack_operand(w, kind, ci_p, co_p, S=1, dtype=torch.bfloat16):
    """bf16 operand copy of a conv / deconv weight in the layout one of the kernels consumes.  A pure index
    permutation + zero padding, so it is also applied to tensors of element INDICES (dtype float64) to build the
    gather map of TrainGraph.enable_flat()."""
    if kind == "fwd":          # conv, cross-correlation: [tap][co][ci]
        return ops.pack_conv_weight(w, ci_p, co_p, dtype)
    if kind == "fwd_tapmajor":  # 1x1 aggregation conv over tap-major Meta-Kernel channels
        return ops.pack_conv_weight(ops.tap_major_weight(w, w.shape[1] // 9), ci_p, co_p, dtype)
    if kind == "dgrad":        # stride 1: flipped taps, transposed channels: [tap][ci][co]
        return ops.pack_conv_weight(w.flip(2, 3).transpose(0, 1), co_p, ci_p, dtype)
    if kind.startswith("dgrad@"):   # "dgrad@lo:hi": data gradient of input channels [lo, hi) only
        lo, hi = (int(v) for v in kind[6:].split(":"))
        return ops.pack_conv_weight(w[:, lo:hi].flip(2, 3).transpose(0, 1), co_p, ci_p, dtype)
    if kind == "dgrad_tapmajor":
        return ops.pack_conv_weight(ops.tap_major_weight(w, w.shape[1] // 9).transpose(0, 1), co_p, ci_p, dtype)
    if kind == "dgrad_s2":     # W-stride 2: transposed conv (3,3)/(1,2), taps not flipped; 1x1 -> centre tap
        if w.shape[2] == 1:
            w3 = torch.zeros((w.shape[0], w.shape[1], 3, 3), device=w.device, dtype=w.dtype)
            w3[:, :, 1, 1] = w[:, :, 0, 0]
            w = w3
        return ops.pack_deconv_weight(w, co_p, ci_p, dtype)  # (Cin_t = co, Cout_t = ci)
    if kind == "deconv_fwd":
        return ops.pack_deconv_weight(w, ci_p, co_p, dtype)
    if kind == "deconv_dgrad":  # 3x3 stride-1 conv over the phase-grouped dz: [tap][ci][(ph,co)]
        ci, co, _, KW = w.shape
        pad = KW // 4
        g = torch.zeros((3, 3, ci_p, S, co_p), device=w.device, dtype=w.dtype)
        for tx in range(3):
            for ph in range(S):
                kx = ph + pad - (1 - tx) * S
                if 0 <= kx < KW:
                    g[:, tx, :ci, ph, :co] = w[:, :, :, kx].permute(2, 0, 1)
        return g.reshape(9, ci_p, S * co_p).to(dtype).contiguous()
    raise KeyError(kind)


def _index_like(t, base):
    """float64 tensor shaped like t holding base+1, base+2, ... (0 stays free to mean 'zero padding')."""
    return (torch.arange(t.numel(), device=t.device, dtype=torch.float64) + (base + 1)).reshape(t.shape)


def _to_map(ix):
    return (ix.reshape(-1).to(torch.int64) - 1).to(torch.int32)


def _align(n, a=64):
    return (n + a - 1) // a * a


class _Pool:
    """Named activation / gradient buffers reused across steps.  Kernels only ever write the interior, so
    a haloed buffer zero-initialised once keeps a valid zero halo."""

    def __init__(self, device, dtype=torch.bfloat16):
        self.device, self.dtype, self.bufs = device, dtype, {}

    def get(self, key, shape):
        t = self.bufs.get(key)
        if t is None or tuple(t.shape) != tuple(shape):
            t = torch.zeros(tuple(shape), device=self.device, dtype=self.dtype)
            self.bufs[key] = t
        return t


class TrainGraph(object):
    """forward(data, coord) -> (cls_logit[3], bbox_delta[3]) fp32 NCHW; backward(d_cls, d_reg) -> {name: grad}."""

    def __init__(self, params, device="cuda", use_meta=True, act_dtype=torch.bfloat16):
        """act_dtype: storage type of activations, activation gradients and weight operands -- torch.float16 (what the
        reference trains in, config:35, with loss scale 128) or torch.bfloat16; accumulation, statistics, parameter
        gradients and master parameters are fp32 either way."""
        if act_dtype not in ops.ACT_DTYPES:
            raise TypeError("act_dtype must be torch.bfloat16 or torch.float16, got %r" % (act_dtype,))
        self.P = params  # fp32 master parameters (reference names), updated in place by the optimiser
        self.device = device
        self.use_meta = use_meta
        self.act_dtype = act_dtype
        self.fuse_stats = True     # BatchNorm batch statistics in the conv epilogue (False: separate rd_bn_train_stats pass)
        # BatchNorm-backward sums in the epilogue of the data-gradient conv above: True / False / "auto".  Per layer width at
        # B=2, 128 channels (scripts/bwdsums_ab.py, conv + BN backward, separate -> fused, profiles/r02_bwdsums_per_shape.jsonl):
        # W=2656 151 -> 130 us, W=1328 87 -> 77, W=664 55 -> 45, W=332 43 -> 46, W=166 43 -> 45 (single-tile launches: the extra
        # z tile is pure latency there; "auto" = layers of at least 50 000 pixels).  Whole step: 17.16 (off) / 16.67 (all) /
        # 16.70 ms (auto), profiles/r02_ab_bwdsums.jsonl -- inside the step the BN passes share HBM with the weight-gradient
        # stream, so even the narrow layers come out ahead: True.
        self.fuse_bwd_sums = True
        self.bn_below, self.bsums = {}, {}
        # False: the head outputs stay in the NHWC 16-bit tensors the head convs write (head_nhwc) and forward() returns
        # (None, None) -- for a loss that reads / writes that layout (rpn_loss_levels_nhwc); head_outputs() converts on demand
        self.fp32_heads = True
        self.head_nhwc = []
        self._stats_bufs = {}
        self.pool = _Pool(device, act_dtype)
        self.packed = {}
        self.tape = []
        self.grads = {}
        self.pgrads = {}
        self.nograd = set()
        self._n = 0
        self.debug = None  # set to a dict to capture intermediates (diagnostics)
        # flat mode (enable_flat): one gather launch packs every operand, one collects every gradient
        self.packed32 = {}
        self.pack_rec, self.pack32_rec = {}, {}
        self.flat_pack = False
        self.flat_grads = False
        self.arena = None          # fp32 scratch all parameter-gradient kernels write into (flat mode)
        self.g_sizes, self._gn = [], 0
        self.pg_rec = {}
        # optional second stream for the weight-gradient kernels: dW of a layer only feeds the optimiser, so it can
        # overlap the HBM-bound BatchNorm backward / data gradient of the layers below it (fork-join, capturable)
        self.side = None
        self._side_used = False

    # ---- parameter packing (bf16 operand copies; call refresh() after an optimiser step) -------------
    def refresh(self):
        if self.flat_pack:
            ops.gather_to_bf16(self.flatP, self.pmap, self.packed_flat)
            if self.pmap32.numel():
                ops.gather_f32(self.flatP, self.pmap32, self.packed32_flat)
            return
        self.packed = {}
        self.packed32 = {}

    def enable_flat(self, flatP, offsets, flat_g, run_step):
        """Switch to flat mode.  `flatP`: flat fp32 buffer holding every trainable parameter, `offsets[name]` its
        element offset (self.P[name] must already be views of it); `flat_g`: flat gradient buffer with the same
        layout; `run_step()`: runs one eager forward + backward.  Must follow at least one eager step (so every
        operand key and gradient source is recorded).  Afterwards
          refresh()   = ONE gather launch re-packing all bf16 operands (+ one for the small fp32 ones),
          backward()  = ... + ONE gather launch collecting all parameter gradients into flat_g."""
        dev = flatP.device
        self.flatP, self.flat_g, self.offsets = flatP, flat_g, offsets
        # -- operands: push element indices through the same packing functions
        keys = sorted(self.pack_rec)
        sizes = [_align(self.packed[k].numel(), 128) for k in keys]
        op_dtype = self.packed[keys[0]].dtype if keys else self.act_dtype     # the kernels' operand type (act_dtype)
        self.packed_flat = torch.zeros(sum(sizes), device=dev, dtype=op_dtype)
        self.pmap = torch.full((sum(sizes),), -1, device=dev, dtype=torch.int32)
        o, views = 0, {}
        for k, n in zip(keys, sizes):
            w = self.P[k[0] + "_weight"]
            ix = pack_operand(_index_like(w, offsets[k[0] + "_weight"]), k[1], *self.pack_rec[k], dtype=torch.float64)
            assert ix.shape == self.packed[k].shape
            self.pmap[o:o + ix.numel()] = _to_map(ix)
            views[k] = self.packed_flat[o:o + ix.numel()].view(ix.shape)
            o += n
        keys32 = sorted(self.pack32_rec)
        ix32 = []
        for k in keys32:
            names, fn = self.pack32_rec[k]
            ix32.append(fn(*[_index_like(self.P[n], offsets[n]) for n in names]))
        sizes32 = [_align(t.numel(), 64) for t in ix32]
        self.packed32_flat = torch.zeros(sum(sizes32), device=dev, dtype=torch.float32)
        self.pmap32 = torch.full((sum(sizes32),), -1, device=dev, dtype=torch.int32)
        o, views32 = 0, {}
        for k, ix, n in zip(keys32, ix32, sizes32):
            self.pmap32[o:o + ix.numel()] = _to_map(ix)
            views32[k] = self.packed32_flat[o:o + ix.numel()].view(ix.shape)
            o += n
        self.packed, self.packed32 = views, views32
        self.flat_pack = True
        # -- gradients: every producing kernel writes into one arena; record where each gradient element lives
        self.g_offsets, o = [], 0
        for n in self.g_sizes:
            self.g_offsets.append(o)
            o += _align(n, 64)
        self.arena = torch.zeros(max(o, 64), device=dev, dtype=torch.float32)
        self.arena_sizes = list(self.g_sizes)
        run_step()   # eager, arena-backed: pg_rec now refers to arena views
        gmap = torch.full((flat_g.numel(),), -1, device=dev, dtype=torch.int32)
        seen = set()
        for name, (src, fn) in self.pg_rec.items():
            if name not in offsets:
                continue
            base = (src.data_ptr() - self.arena.data_ptr()) // 4
            assert 0 <= base and base + src.numel() <= self.arena.numel(), name
            ix = _index_like(src, base)
            ix = fn(ix) if fn is not None else ix
            n = self.P[name].numel()
            assert ix.numel() == n, (name, tuple(ix.shape), tuple(self.P[name].shape))
            gmap[offsets[name]:offsets[name] + n] = _to_map(ix)
            seen.add(name)
        # parameters the graph does not use (e.g. the Meta-Kernel MLP when use_meta=False) keep gmap = -1: zero gradient
        self.no_grad_params = sorted(n for n in offsets if n not in seen)
        self.gmap = gmap
        self.flat_grads = True

    def _w(self, name, kind, ci_p, co_p, S=1):
        key = (name, kind)
        t = self.packed.get(key)
        if t is not None:
            return t
        if self.flat_pack:
            raise KeyError("operand %s was not recorded before enable_flat()" % (key,))
        self.pack_rec[key] = (ci_p, co_p, S)
        t = pack_operand(self.P[name + "_weight"], kind, ci_p, co_p, S, dtype=self.act_dtype)
        self.packed[key] = t
        return t

    def _p32(self, key, names, fn):
        """Small fp32 operand derived from parameters by a pure index permutation / zero padding `fn`
        (padded head bias, tap-major BN(576) vectors).  Flat mode: a view of the buffer refresh() gathers."""
        t = self.packed32.get(key)
        if t is not None:
            return t
        if self.flat_pack:
            raise KeyError("operand %s was not recorded before enable_flat()" % (key,))
        self.pack32_rec[key] = (names, fn)
        return fn(*[self.P[n] for n in names])

    # ---- tape helpers ---------------------------------------------------------------------------------
    def _stats_ws(self, C):
        """Per-call fp32 workspace of the fused statistics (reused across steps in tape order, like the activations)."""
        self._n += 1
        key = "stats#%d" % self._n
        t = self._stats_bufs.get(key)
        if t is None or t.numel() < 1184 * 2 * C:
            t = self._stats_bufs[key] = torch.zeros(1184 * 2 * C, device=self.device, dtype=torch.float32)
        return t

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
        self.bn_below, self.bsums = {}, {}
        self._gn = 0
        if self.arena is None:
            self.g_sizes = []

    def _gbuf(self, shape):
        """fp32 output buffer of a parameter-gradient kernel: fresh memory (eager mode) or the next slot of the
        gradient arena (flat mode; slots are handed out in the fixed order of the tape)."""
        n = 1
        for d in shape:
            n *= int(d)
        if self.arena is None:
            self.g_sizes.append(n)
            return torch.empty(tuple(shape), device=self.device, dtype=torch.float32)
        i = self._gn
        self._gn += 1
        assert self.arena_sizes[i] == n, "gradient arena slot %d: %d elements expected, %d requested" % (i, self.arena_sizes[i], n)
        o = self.g_offsets[i]
        return self.arena[o:o + n].view(tuple(shape))

    def seed_grad(self, t, g):
        """Set the gradient of activation t (haloed NHWC bf16) before run_tape()."""
        self.grads[id(t)] = g

    def grad_of(self, t):
        return self.grads[id(t)]

    def run_tape(self):
        for fn in reversed(self.tape):
            fn()
        self.tape = []

    def _wgrad(self, a_pad, b_pad, ksize, stride_w, out):
        if self.side is None:
            return ops.conv2d_wgrad(a_pad, b_pad, ksize, stride_w, out=out)
        ev = torch.cuda.Event()
        ev.record()
        self.side.wait_event(ev)
        with torch.cuda.stream(self.side):
            G = ops.conv2d_wgrad(a_pad, b_pad, ksize, stride_w, out=out)
        self._side_used = True
        return G

    def _join_side(self):
        if self.side is not None and self._side_used:
            torch.cuda.current_stream().wait_stream(self.side)
            self._side_used = False

    def _pg(self, name, src, fn=None, copy=False):
        """Gradient of parameter `name` = fn(src): `src` is what a kernel wrote (into a _gbuf() buffer, or anywhere
        with copy=True), `fn` a pure view / permutation into the parameter's layout.  Eager mode materialises
        fn(src) in the dictionary backward() returns; flat mode leaves it to the single gather launch."""
        if copy:
            slot = self._gbuf(src.shape)
            slot.copy_(src)
            src = slot
        assert name not in self.pgrads, "parameter %s has two gradient sources" % name
        self.pg_rec[name] = (src, fn)
        if not self.flat_grads:
            g = fn(src) if fn is not None else src
            self.pgrads[name] = g.contiguous()
        else:
            self.pgrads[name] = None

    # ---- layers -----------------------------------------------------------------------------------------
    def conv_bn(self, x, wname, bnname, stride_w=1, relu=True, res_before=None, kinds=("fwd", "dgrad"), dx_channels=None,
                sole_consumer=False):
        """dx_channels = (lo, hi, t): x is a channel concatenation whose channels [lo, hi) are the tensor t and whose other
        channels need no gradient -- the backward computes the data gradient of that slice only, straight into grad(t).
        sole_consumer: nothing else reads x, so this layer's data gradient IS grad(x); if x is the output of a conv_bn
        without residual, its BatchNorm-backward sums are accumulated by this layer's data-gradient kernel."""
        P = self.P
        w = P[wname + "_weight"]
        co, ci, k = w.shape[0], w.shape[1], w.shape[2]
        ci_p, co_p = x.shape[3], _cout_pad(co)
        N, Hp, Wp, _ = x.shape
        W_out = (Wp - 2) // stride_w
        if self.fuse_stats:
            # batch statistics ride the conv epilogue (one full read of z less per layer); each layer owns its partials
            z, part, nslots = ops.conv2d_nhwc_stats(x, self._w(wname, kinds[0], ci_p, co_p), stride_w=stride_w,
                                                    out=self._buf("z", (N, Hp, W_out + 2, co_p)), ws=self._stats_ws(co_p))
            coef = ops.bn_train_finalize(part, nslots, N, Hp - 2, W_out, co_p, P[bnname + "_gamma"], P[bnname + "_beta"],
                                         P[bnname + "_moving_mean"], P[bnname + "_moving_var"])
        else:
            z = ops.conv2d_nhwc(x, self._w(wname, kinds[0], ci_p, co_p), relu=False, stride_w=stride_w,
                                out=self._buf("z", (N, Hp, W_out + 2, co_p)))
            coef = ops.bn_train_stats(z, P[bnname + "_gamma"], P[bnname + "_beta"], P[bnname + "_moving_mean"],
                                      P[bnname + "_moving_var"])
        y = ops.bn_act_fwd(z, coef, relu=relu, res_before=res_before, out=self._buf("y", z.shape))
        # ReLU mask: without a residual the pre-activation is z*a+b, recomputed from the z the kernels read
        # anyway (mask_mode 2) -- saves one full read of y in each of the two backward passes
        mm = 0 if not relu else (2 if res_before is None else 1)
        if res_before is None:
            self.bn_below[id(y)] = (z, coef, mm)
        want = self.fuse_bwd_sums
        if want == "auto":
            want = N * (Hp - 2) * (Wp - 2) >= 50000
        below = self.bn_below.get(id(x)) if (sole_consumer and want and dx_channels is None and stride_w == 1 and k == 3
                                             and ops.conv_bwdstats_supported(co_p, ci_p)) else None

        def bwd():
            dy = self.grads.pop(id(y))
            dgb = self._gbuf((2, co_p))
            dz, dgamma, dbeta, g = ops.bn_act_bwd(dy, z, coef, mm, y_mask=y if mm == 1 else None, dz_out=self._buf("dz", z.shape),
                                                  want_g=res_before is not None and relu,
                                                  g_out=self._buf("g", z.shape) if (res_before is not None and relu) else None,
                                                  dgb_out=dgb, sums=self.bsums.pop(id(y), None))
            if res_before is not None:
                self._acc(res_before, g if relu else dy)
            self._pg(bnname + "_gamma", dgb, lambda t: t[0, :co])
            self._pg(bnname + "_beta", dgb, lambda t: t[1, :co])
            G = self._wgrad(dz, x, k, stride_w, self._gbuf((k * k, co_p, ci_p)))  # [tap][co_p][ci_p]
            if kinds[0] == "fwd_tapmajor":
                C = ci // 9
                self._pg(wname + "_weight", G, lambda t: t[0, :co, :ci].reshape(co, 9, C).transpose(1, 2).reshape(co, ci, 1, 1))
            else:
                self._pg(wname + "_weight", G, lambda t: t[:, :co, :ci].reshape(k, k, co, ci).permute(2, 3, 0, 1))
            if dx_channels is not None:
                lo, hi, t = dx_channels
                assert stride_w == 1 and k == 3 and hi - lo == t.shape[3]
                self.grads[id(t)] = ops.conv2d_nhwc(dz, self._w(wname, "dgrad@%d:%d" % (lo, hi), hi - lo, co_p), relu=False,
                                                    residual_pad=self.grads.pop(id(t), None), out=self._buf("dx", t.shape))
                return
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
            elif below is not None and old is None:
                # grad(x) is complete with this kernel: it also accumulates the sums of the BatchNorm that produced x
                dx, part, nslots = ops.conv2d_nhwc_bwdstats(dz, self._w(wname, kinds[1], ci_p, co_p), below[0], below[1], below[2],
                                                            out=self._buf("dx", x.shape), ws=self._stats_ws(ci_p))
                self.bsums[id(x)] = (part, nslots)
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
            dgb = self._gbuf((2, co_p))
            dzg, dgamma, dbeta, _ = ops.bn_act_bwd(dy, z, coef, 2, dz_halo_w=S,
                                                   dz_out=self._buf("dzg", (N, Hp, W_in * S + 2 * S, co_p)), dgb_out=dgb)
            self._acc(const, dy)
            self._pg(bnname + "_gamma", dgb, lambda t: t[0, :co])
            self._pg(bnname + "_beta", dgb, lambda t: t[1, :co])
            dzg = dzg.view(N, Hp, W_in + 2, S * co_p)  # phase-grouped: pixel group j holds output pixels j*S .. j*S+S-1
            G = self._wgrad(up, dzg, 3, 1, self._gbuf((9, ci_p, S * co_p)))  # [ty][tx][ci][ph][co]
            pad = KW // 4

            def gw_of(t):
                t = t.reshape(3, 3, ci_p, S, co_p)
                cols = []
                for kx in range(KW):
                    tx, ph = divmod(kx - pad + S, S)
                    cols.append(t[:, tx, :ci, ph, :co].permute(1, 2, 0))     # (ci, co, 3)
                return torch.stack(cols, 3)                                     # (ci, co, 3, KW)

            self._pg(wname + "_weight", G, gw_of)
            old = self.grads.pop(id(up), None)
            self.grads[id(up)] = ops.conv2d_nhwc(dzg, self._w(wname, "deconv_dgrad", ci_p, co_p, S), relu=False,
                                                 residual_pad=old, out=self._buf("dx", up.shape))

        self.tape.append(bwd)
        return y

    def head_out(self, x, wname, co):
        """1x1 conv + bias, no norm (builder.py:247-262) -> fp32 NCHW."""
        P = self.P
        ci_p = x.shape[3]
        def pad64(b):
            t = torch.zeros(64, device=b.device, dtype=b.dtype)
            t[:co] = b
            return t

        bias = self._p32((wname, "bias64"), [wname + "_bias"], pad64)
        zp = ops.conv2d_nhwc(x, self._w(wname, "fwd", ci_p, 64), None, bias, relu=False, out=self._buf("z", x.shape[:3] + (64,)))
        out = ops.nhwc_to_nchw(zp, co) if self.fp32_heads else None
        dzb = self._buf("dz", zp.shape)     # zero-initialised once; only channels < co are ever written
        self.head_nhwc.append((zp, dzb, co))

        def bwd(d_out):
            # d_out None: the gradient is already in dzb (written there by the NHWC loss)
            dz = dzb if d_out is None else ops.nchw_to_nhwc(d_out.contiguous(), dzb)
            self._pg(wname + "_bias", ops.channel_sums(dz, out=self._gbuf((dz.shape[3],))), lambda t: t[:co])
            G = self._wgrad(dz, x, 1, 1, self._gbuf((1, 64, ci_p)))
            wshape = tuple(P[wname + "_weight"].shape)
            self._pg(wname + "_weight", G, lambda t: t[0, :co, :wshape[1]].reshape(wshape))
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
        gamma_t, beta_t = self._p32((bn, "gamma_tm"), [bn + "_gamma"], tm), self._p32((bn, "beta_tm"), [bn + "_beta"], tm)
        mm_t, mv_t = tm(P[bn + "_moving_mean"]), tm(P[bn + "_moving_var"])
        coef = ops.bn_train_stats(m, gamma_t, beta_t, mm_t, mv_t)
        P[bn + "_moving_mean"].copy_(untm(mm_t))
        P[bn + "_moving_var"].copy_(untm(mv_t))
        a = ops.bn_act_fwd(m, coef, relu=True, out=self._buf("meta_act", m.shape))

        def bwd():
            da = self.grads.pop(id(a))
            dgb = self._gbuf((2, 9 * C))
            dm, dgamma, dbeta, _ = ops.bn_act_bwd(da, m, coef, 2, dz_out=self._buf("dmeta", m.shape), dgb_out=dgb)
            self._pg(bn + "_gamma", dgb, lambda t: untm(t[0]))
            self._pg(bn + "_beta", dgb, lambda t: untm(t[1]))
            # the Meta-Kernel backward reads the NHWC tap-major gradient as it is (the reference op boundary would be
            # grad_out (B, 576 = c*9+k, H, W) fp32: a 2x larger tensor and a layout pass)
            gd, gw0, gb0, gw1, gb1 = ops.meta_kernel_backward_nhwc(dm, feat, coord, *mlp)
            if self.debug is not None:
                self.debug.update(meta_da=da, meta_dm=dm, meta_gd=gd, meta_m=m, meta_a=a)
            self._pg(name + "_2656_mlp0_weight", gw0, lambda t: t.reshape(32, 3, 1, 1), copy=True)
            self._pg(name + "_2656_mlp0_bias", gb0, copy=True)
            self._pg(name + "_2656_mlp1_weight", gw1, lambda t: t.reshape(-1, 32, 1, 1), copy=True)
            self._pg(name + "_2656_mlp1_bias", gb1, copy=True)
            self._acc(x, ops.nchw_to_nhwc(gd, self._buf("dx", x.shape)))

        self.tape.append(bwd)
        return a

    def basicblock(self, x, coord, name, stride_w, proj):  # dla_backbone.py:17-56
        if self.use_meta and name in META_UNITS:
            r1 = self.meta_kernel_conv(x, coord, name)
        else:
            r1 = self.conv_bn(x, name + "_conv1", name + "_bn1")
        sc = self.conv_bn(x, name + "_sc", name + "_sc_bn", stride_w=stride_w, relu=False) if proj else x
        return self.conv_bn(r1, name + "_conv2", name + "_bn2", stride_w=stride_w, relu=True, res_before=sc, sole_consumer=True)

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
        self.low_tape_start = len(self.tape)      # res2a, res2: bucket 2; res1 (before this mark): bucket 3, the last of the backward
        res2a = self.res_stage(res1, None, "res2a", 2)
        res2 = self.res_stage(res2a, None, "res2", 2)
        self.mid_tape_start = len(self.tape)      # res3a, res3 and the aggregation stages: bucket 1 of the backward
        res3a = self.res_stage(res2, None, "res3a", 2)
        res3 = self.res_stage(res3a, None, "res3", 2)
        agg2 = self.agg_stage("agg2", res2, res3)
        agg1 = self.agg_stage("agg1", res1, res2)
        agg2a = self.agg_stage("agg2a", res2a, agg2)
        agg3 = self.agg_stage("agg3", agg1, agg2a)
        cat = self._buf("data_concat", agg3.shape[:3] + (128,))  # concat(data, agg3): 72 of 128 channels
        ops.copy_channels(x, 0, cat, 0, c)          # c = 8 input channels
        ops.copy_channels(agg3, 0, cat, c, 64)
        self.head_tape_start = len(self.tape)     # tape entries from here on belong to the RPN head towers
        self.head_bwd = []
        self.head_nhwc = []                       # [(z_pad, dz_pad, channels)] in the order cls lvl 0, reg lvl 0, cls lvl 1, ...
        cls_logit, bbox_delta = [], []
        for lvl, f in enumerate([cat, agg2a, agg2]):
            t_c = t_r = f
            # level 0 reads concat(data, agg3): only the agg3 channels carry a gradient (builder.py:198-266)
            dxc = (c, c + 64, agg3) if lvl == 0 else None
            for i in range(4):
                n = "rpn_cls_conv_%d_lvl_%d" % (i, lvl)
                t_c = self.conv_bn(t_c, n, n + "_bn", dx_channels=dxc if i == 0 else None, sole_consumer=i > 0)
                n = "rpn_reg_conv_%d_lvl_%d" % (i, lvl)
                t_r = self.conv_bn(t_r, n, n + "_bn", dx_channels=dxc if i == 0 else None, sole_consumer=i > 0)
            o, b = self.head_out(t_c, "rpn_cls_logit_lvl_%d" % lvl, 1)
            cls_logit.append(o)
            self.head_bwd.append(("cls", lvl, b, len(self.tape)))
            o, b = self.head_out(t_r, "rpn_reg_delta_lvl_%d" % lvl, 8)
            bbox_delta.append(o)
            self.head_bwd.append(("reg", lvl, b, len(self.tape)))
        return cls_logit, bbox_delta

    def head_outputs(self):
        """-> (cls_logit[3], bbox_delta[3]) fp32 NCHW from the NHWC head tensors of the last forward() (what forward()
        returns itself unless fp32_heads is off)."""
        cls = [ops.nhwc_to_nchw(z, co) for z, _, co in self.head_nhwc[0::2]]
        reg = [ops.nhwc_to_nchw(z, co) for z, _, co in self.head_nhwc[1::2]]
        return cls, reg

    def backward(self, d_cls, d_reg):
        """d_cls[l] (B,1,H,W_l), d_reg[l] (B,8,H,W_l) fp32: gradients of the loss w.r.t. the head outputs (None, None: they
        are already in the NHWC gradient tensors of head_nhwc).  Returns {parameter name: fp32 gradient in the reference's shape}."""
        for kind, lvl, b, _ in self.head_bwd:
            b(None if d_cls is None else (d_cls[lvl] if kind == "cls" else d_reg[lvl]))
        self.run_tape()
        self._join_side()
        if self.flat_grads:   # one launch: every parameter gradient -> the flat (all-reduce) buffer
            ops.gather_f32(self.arena, self.gmap, self.flat_g)
        return self.pgrads

    # ---- backward in buckets (flat mode): gradient exchange overlapping the rest of the backward ----------------
    N_BUCKETS = 4

    @staticmethod
    def bucket_of(name):
        """Backward runs head -> aggregation stages -> res3 -> res3a -> res2 -> res2a -> res1.  Bucket 0: the RPN head
        (`rpn_*`, 13.1 MB of gradients); bucket 1: aggregation stages + res3 / res3a (large parameters, early and cheap
        backward, 18.9 MB); bucket 2: res2 / res2a (4.3 MB); bucket 3: res1 incl. the Meta-Kernel unit -- 0.4 MB of
        parameters behind the longest stretch of the backward (the 2656-wide layers), so the res2 / res2a exchange hides
        behind it and only a latency-sized exchange is left exposed after the backward."""
        if name.startswith("rpn_"):
            return 0
        if name.startswith(("agg", "res3_", "res3a_")):
            return 1
        if name.startswith(("res2_", "res2a_")):
            return 2
        return 3

    def bucket_ranges(self):
        """-> [[(lo, hi), ...] per bucket]: contiguous element ranges of the name-sorted flat buffers."""
        names = sorted(self.offsets)
        out = [[] for _ in range(self.N_BUCKETS)]
        for n in names:
            k, lo = self.bucket_of(n), self.offsets[n]
            hi = lo + self.P[n].numel()
            if out[k] and out[k][-1][1] == lo:
                out[k][-1] = (out[k][-1][0], hi)
            else:
                out[k].append((lo, hi))
        return out

    def backward_bucket(self, k, d_cls=None, d_reg=None):
        """Bucket k of backward() (call 0 .. N_BUCKETS-1 in order): its slice of the tape, then its parameter gradients
        gathered into their ranges of flat_g.  Flat mode only."""
        s0, s1, s2 = self.low_tape_start, self.mid_tape_start, self.head_tape_start
        if k == 0:
            for kind, lvl, b, _ in self.head_bwd:
                b(None if d_cls is None else (d_cls[lvl] if kind == "cls" else d_reg[lvl]))
            seg = self.tape[s2:]
        elif k == 1:
            seg = self.tape[s1:s2]
        elif k == 2:
            seg = self.tape[s0:s1]
        else:
            seg = self.tape[:s0]
        for fn in reversed(seg):
            fn()
        self._join_side()
        for lo, hi in self.bucket_ranges()[k]:
            ops.gather_f32(self.arena, self.gmap[lo:hi], self.flat_g[lo:hi])
        if k == self.N_BUCKETS - 1:
            self.tape = []
        return self.pgrads


STRIDES = (1, 2, 4)
LOSS_HYPER = dict(iou_type="bev", alpha=1.0, gamma=2.0, smooth_l1_scalar=3.0, scale_loss_shift=128.0,
                  cls_loss_weight=10.0, reg_loss_weight=8.0)   # config/rangedet/rangedet_veh_wo_aug_4_18e.py:36,122-129


def rpn_loss_levels(cls_logit, bbox_delta, targets, out=None, hyper=LOSS_HYPER, gt_name="gt_bbox_veh_for_iou_pred"):
    """RangeRpnHead.get_fpn_loss (builder.py:268-348) on the head outputs: per level one fused launch pair
    (ops.rpn_loss).  `targets`: dict of CUDA tensors with the graph-input names of builder.py:20-37.
    Returns [per-level dict(iou_target, cls_loss, reg_loss, d_cls, d_reg)]."""
    res = []
    for lvl, s in enumerate(STRIDES):
        res.append(ops.rpn_loss(cls_logit[lvl], bbox_delta[lvl], targets["pc_vehicle_frame_s%d" % s],
                                targets[gt_name], targets["range_image_mask_s%d" % s],
                                targets["rpn_reg_target_s%d" % s], targets["rpn_reg_weight_s%d" % s],
                                targets["reg_normalize_weight_s%d" % s], out=None if out is None else out[lvl], **hyper))
    return res


def rpn_loss_levels_nhwc(head_nhwc, targets, out=None, hyper=LOSS_HYPER, gt_name="gt_bbox_veh_for_iou_pred"):
    """rpn_loss_levels on TrainGraph.head_nhwc: the loss reads the logits / deltas from the NHWC 16-bit tensors the head
    convolutions wrote and writes their gradients into the NHWC tensors the head backward reads -- no fp32 planar copies
    in either direction (same values: those copies only ever held widened 16-bit numbers)."""
    res = []
    for lvl, s in enumerate(STRIDES):
        (zc, dzc, _), (zr, dzr, _) = head_nhwc[2 * lvl], head_nhwc[2 * lvl + 1]
        o = None if out is None else {k: v for k, v in out[lvl].items() if k in ("iou_target", "cls_loss", "reg_loss")}
        res.append(ops.rpn_loss_nhwc(zc, zr, targets["pc_vehicle_frame_s%d" % s], targets[gt_name],
                                     targets["range_image_mask_s%d" % s], targets["rpn_reg_target_s%d" % s],
                                     targets["rpn_reg_weight_s%d" % s], targets["reg_normalize_weight_s%d" % s], dzc, dzr,
                                     out=o, **hyper))
    return res


def wd_mult(name):
    """MXNet Optimizer.set_wd_mult: weight decay only on *_weight and *_gamma."""
    return 1.0 if name.endswith(("_weight", "_gamma")) else 0.0


def sgd_momentum_step(params, grads, momenta, lr, momentum=0.9, wd=1e-4, clip_gradient=None, rescale_grad=1.0):
    """MXNet SGD (tools/train.py:312-319), per-tensor torch restatement used by the eager path and as the
    check of rd_sgd_mom_update: g = clip(rescale*g) + wd*wd_mult*w; m = momentum*m - lr*g; w += m."""
    for n in grads:
        if n not in params:
            continue
        g = grads[n].to(params[n].dtype).reshape(params[n].shape) * rescale_grad
        if clip_gradient is not None:
            g = g.clamp(-clip_gradient, clip_gradient)
        g = g + wd * wd_mult(n) * params[n]
        if n not in momenta:
            momenta[n] = torch.zeros_like(params[n])
        momenta[n].mul_(momentum).add_(g, alpha=-lr)
        params[n].add_(momenta[n])


class GraphedTrainStep(object):
    """The training iteration captured ONCE in CUDA graphs and replayed: forward | loss + backward | update.

        step = GraphedTrainStep(params, batch, H, W, lr=...)
        step.set_targets(targets)                      # synthetic roidb record (names of builder.py:20-37)
        step.train_step(data, coord)                   # forward, RPN loss, backward, all-reduce, SGD
      or, with an external loss:
        cls, reg = step.forward(data, coord); step.backward_update(d_cls, d_reg)

    The step is ~950 launches of 10-500 us kernels; launching them one by one from Python is host-bound, and the
    per-parameter plumbing (operand packing, gradient extraction, optimiser) used to be ~2 800 more.  Here the
    trainable parameters live in ONE flat fp32 buffer (`params[name]` become views of it), TrainGraph runs in flat
    mode (one gather launch packs all bf16 operands, one collects all gradients) and the MXNet SGD-momentum
    update is one launch (rd_sgd_mom_update) whose learning rate is read from device memory (set_lr).  With
    `allreduce` given, the flat gradient buffer is summed across ranks between backward and update and averaged through
    rescale_grad / world_size (the reference: hvd.DistributedOptimizer, tools/train.py:364-368).  `allreduce(view)` must
    SUM the given contiguous view of the flat buffer over ranks; it may return a handle with `.wait()` (e.g.
    dist.all_reduce(..., async_op=True)): with world_size > 1 the captured backward is then split into four buckets (head |
    aggregation stages + res3 / res3a | res2 / res2a | res1, TrainGraph.bucket_of) and each bucket's exchange overlaps the
    backward kernels of the buckets after it; only the last, smallest one (5 MB) is exposed."""

    def __init__(self, params, batch, H, W, lr, momentum=0.9, wd=1e-5, clip_gradient=35.0, rescale_grad=1.0 / 128.0,
                 device="cuda", use_meta=True, allreduce=None, world_size=1, with_loss=True, overlap_wgrad=True,
                 loss_hyper=None, gt_name="gt_bbox_veh_for_iou_pred", capture=True, act_dtype=torch.bfloat16,
                 overlap_allreduce=True, fuse_stats=True, fuse_bwd_sums=True, nhwc_loss=True):
        self.P = params
        self.allreduce = allreduce
        self.with_loss = with_loss
        self.loss_hyper = dict(LOSS_HYPER if loss_hyper is None else loss_hyper)
        self.gt_name = gt_name
        names = sorted(k for k in params if not k.endswith(("_moving_mean", "_moving_var")))
        self.names = names
        sizes = [params[k].numel() for k in names]
        total = sum(sizes)
        self.flatP = torch.empty(total, device=device)
        self.flat = torch.zeros(total, device=device)       # gradients (all-reduced)
        self.flat_m = torch.zeros(total, device=device)     # momentum
        self.flat_wd = torch.empty(total, device=device)
        self.offsets, o = {}, 0
        for k, n in zip(names, sizes):
            self.offsets[k] = o
            self.flatP[o:o + n].copy_(params[k].reshape(-1))
            params[k] = self.flatP[o:o + n].view(params[k].shape)   # masters now live in the flat buffer
            self.flat_wd[o:o + n] = wd * wd_mult(k)
            o += n
        self.gviews = {k: self.flat[self.offsets[k]:self.offsets[k] + n].view(params[k].shape) for k, n in zip(names, sizes)}
        # `allreduce(flat)` SUMS the flat gradient over ranks; the 1/world_size of the average rides rescale_grad
        self.hyper = torch.tensor([lr, momentum, rescale_grad / world_size, clip_gradient if clip_gradient else 0.0],
                                  device=device)
        self.tg = TrainGraph(params, device, use_meta, act_dtype)
        self.tg.fuse_stats = bool(fuse_stats)     # False: separate rd_bn_train_stats pass after every conv (A/B timing)
        self.tg.fuse_bwd_sums = fuse_bwd_sums     # True / False / "auto" (TrainGraph.__init__)
        # fused loss reading / writing the head tensors in place (NHWC 16-bit): no fp32 planar head outputs or gradients in the
        # step; `out` converts on demand
        self.nhwc_loss = bool(nhwc_loss) and with_loss
        self.tg.fp32_heads = not self.nhwc_loss
        self.act_dtype = act_dtype
        self.capture = capture    # False: same buffers and flat plumbing, kernels launched eagerly (debugging)
        if overlap_wgrad and capture:
            self.tg.side = torch.cuda.Stream(device=device)
        self.data = torch.zeros((batch, 8, H, W), device=device)
        self.coord = torch.zeros((batch, 3, H, W), device=device)
        self.d_cls = [torch.zeros((batch, 1, H, W // s), device=device) for s in STRIDES]
        self.d_reg = [torch.zeros((batch, 8, H, W // s), device=device) for s in STRIDES]
        self.targets, self.loss_out = None, None
        self.launches = None   # kernels per graph replay (captured mode)
        self.split_bwd = False
        if with_loss:
            z = lambda *shape: torch.zeros(shape, device=device)
            self.targets = {gt_name: z(batch, 200, 8 if self.loss_hyper["iou_type"] == "bev" else 7)}
            if self.loss_hyper["iou_type"] == "bev":
                self.targets[gt_name][:, :, 3:7] = 1e-3     # all-padding GT (input.py:264-265)
            else:
                self.targets[gt_name][:, :, 3:6] = 1e-3
            self.loss_out = []
            for s, dc, dr in zip(STRIDES, self.d_cls, self.d_reg):
                self.targets["rpn_reg_target_s%d" % s] = z(batch, 8, H, W // s)
                self.targets["rpn_reg_weight_s%d" % s] = z(batch, 8, H, W // s)
                self.targets["reg_normalize_weight_s%d" % s] = z(batch, 8, H, W // s)
                self.targets["range_image_mask_s%d" % s] = z(batch, 1, H, W // s)
                self.targets["pc_vehicle_frame_s%d" % s] = z(batch, H * W // s, 3)
                self.loss_out.append(dict(iou_target=z(batch, 1, H, W // s), cls_loss=z(batch, 1, H, W // s),
                                          reg_loss=z(batch, 8, H, W // s), d_cls=dc, d_reg=dr))
        # eager passes outside the capture: (1) sizes every buffer / records every operand and gradient source,
        # (2) inside enable_flat: builds the gather maps, (3) flat mode warm-up.  lr = 0 leaves the weights alone.
        lr_saved = self.hyper.clone()
        self.hyper[0] = 0.0
        # the warm-up passes run on all-zero inputs: keep them out of the BatchNorm moving statistics (a checkpoint's
        # values must survive bind(); MXNet's executor has no such passes)
        aux_saved = {k: v.clone() for k, v in params.items() if k.endswith(("_moving_mean", "_moving_var"))}

        def restore_aux():
            for k, v in aux_saved.items():
                params[k].copy_(v)

        def warm():
            self._fwd()
            self._bwd()
            self.tg.enable_flat(self.flatP, self.offsets, self.flat, lambda: (self._fwd(), self._bwd()))
            self._fwd()
            self._bwd()
            self._update()

        if not capture:
            warm()
            self.hyper.copy_(lr_saved)
            self.flat_m.zero_()
            restore_aux()
            return
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            warm()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self.hyper.copy_(lr_saved)
        self.flat_m.zero_()
        self.pool = torch.cuda.graph_pool_handle()
        self.g_fwd, self.g_bwd, self.g_upd = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        # thread_local: another thread's CUDA calls (the NCCL watchdog of a data-parallel job polls events) must not
        # invalidate the capture
        mode = dict(capture_error_mode="thread_local")
        # kernels per replay: the C-ABI counts every launch it issues (rd_launch_count), also under capture
        from . import _lib
        c0 = _lib.launch_count()
        with torch.cuda.graph(self.g_fwd, pool=self.pool, **mode):
            self._fwd()
        c1 = _lib.launch_count()
        # With a gradient exchange the backward is captured as one graph per bucket, so that a bucket's all-reduce runs
        # while the next bucket's backward computes
        self.split_bwd = self.allreduce is not None and world_size > 1 and overlap_allreduce
        if self.split_bwd:
            self.g_bwd_parts = [torch.cuda.CUDAGraph() for _ in range(TrainGraph.N_BUCKETS)]
            for k, g in enumerate(self.g_bwd_parts):
                with torch.cuda.graph(g, pool=self.pool, **mode):
                    self._bwd_bucket(k)
            self.bucket_ranges = self.tg.bucket_ranges()
        else:
            with torch.cuda.graph(self.g_bwd, pool=self.pool, **mode):
                self._bwd()
        c2 = _lib.launch_count()
        with torch.cuda.graph(self.g_upd, pool=self.pool, **mode):
            self._update()
        self.launches = dict(fwd=c1 - c0, bwd=c2 - c1, upd=_lib.launch_count() - c2)
        restore_aux()   # capture executes nothing: this undoes the eager warm-up passes

    @property
    def out(self):
        """(cls_logit[3], bbox_delta[3]) fp32 NCHW of the last forward; with the NHWC loss they are not part of the step and
        are converted here, on demand."""
        return self.tg.head_outputs() if self.nhwc_loss else self._out

    def _loss(self):
        if self.nhwc_loss:
            rpn_loss_levels_nhwc(self.tg.head_nhwc, self.targets, out=self.loss_out, hyper=self.loss_hyper, gt_name=self.gt_name)
        elif self.with_loss:
            rpn_loss_levels(self._out[0], self._out[1], self.targets, out=self.loss_out, hyper=self.loss_hyper, gt_name=self.gt_name)

    def _fwd(self):
        self.tg.refresh()
        self._out = self.tg.forward(self.data, self.coord)

    def _bwd(self):
        self._loss()
        grads = self.tg.backward(*((None, None) if self.nhwc_loss else (self.d_cls, self.d_reg)))
        if not self.tg.flat_grads:
            ks = [k for k in self.names if k in grads]
            torch._foreach_copy_([self.gviews[k] for k in ks], [grads[k].reshape(self.P[k].shape) for k in ks])

    def _bwd_bucket(self, k):
        if k == 0:
            self._loss()
        self.tg.backward_bucket(k, *((None, None) if self.nhwc_loss else (self.d_cls, self.d_reg)))

    def _update(self):
        ops.sgd_mom_update(self.flatP, self.flat, self.flat_m, self.flat_wd, self.hyper)

    def broadcast_parameters(self, src=0):
        """tools/train.py:219-229: all ranks start from rank `src`'s arg + aux parameters (one flat broadcast of the
        master buffer + one of the moving statistics).  Call after construction, before the first step."""
        from . import dist as rd_dist
        rd_dist.broadcast_params_(self.P, src=src, flat=self.flatP)

    def average_aux(self):
        """utils/detection_module.py:1164-1170: epoch-end average of the BatchNorm moving statistics over ranks."""
        from . import dist as rd_dist
        rd_dist.average_aux_(self.P)

    def set_lr(self, lr):
        """Learning-rate schedule: the captured update reads lr from device memory."""
        self.hyper[0:1].fill_(float(lr))

    def set_targets(self, targets):
        """Copy one synthetic-roidb / loader record (numpy or torch, names of builder.py:20-37) into the static
        buffers the captured loss kernels read."""
        for k, dst in self.targets.items():
            src = targets[k]
            src = torch.from_numpy(src) if not isinstance(src, torch.Tensor) else src
            dst.copy_(src.reshape(dst.shape), non_blocking=True)

    def forward(self, data, coord):
        self.data.copy_(data, non_blocking=True)
        self.coord.copy_(coord, non_blocking=True)
        if self.capture:
            self.g_fwd.replay()
        else:
            self._fwd()
        return self.out

    def backward_update(self, d_cls=None, d_reg=None):
        if not self.with_loss:
            for dst, src in zip(self.d_cls + self.d_reg, list(d_cls) + list(d_reg)):
                dst.copy_(src, non_blocking=True)
        if self.split_bwd:
            # bucket k's backward graph, then its (asynchronous) exchange, which overlaps the graphs of the buckets after it;
            # the update waits for all of them
            handles = []
            for g, ranges in zip(self.g_bwd_parts, self.bucket_ranges):
                g.replay()
                for lo, hi in ranges:
                    handles.append(self.allreduce(self.flat[lo:hi]))
            for h in handles:
                if h is not None and hasattr(h, "wait"):
                    h.wait()
            self.g_upd.replay()
            return
        if self.capture:
            self.g_bwd.replay()
        else:
            self._bwd()
        if self.allreduce is not None:
            h = self.allreduce(self.flat)
            if h is not None and hasattr(h, "wait"):
                h.wait()
        if self.capture:
            self.g_upd.replay()
        else:
            self._update()

    def train_step(self, data, coord):
        """One iteration of tools/train.py's fit loop on the device: forward, loss, backward, all-reduce, SGD."""
        self.forward(data, coord)
        self.backward_update()
        return self.loss_out


class HostFedTrainStep(object):
    """The body of tools/train.py's fit loop as the caller sees it: one HOST loader record in, loss values out.

        fed = HostFedTrainStep(step)                  # step: GraphedTrainStep (e.g. from TrainSymbol.bind())
        fed.feed(record)                              # pinned host tensors named like builder.py:20-37 + input_data, coord_s1
        losses = fed.step()                           # forward, loss, backward, all-reduce, SGD on that record
        fed.feed(next_record); ...                    # upload of record i+1 overlaps the kernels of record i
        values = fed.losses()                         # waits; (6,) numpy: [cls_s1, cls_s2, cls_s4, reg_s1, reg_s2, reg_s4] sums

    The reference's loader hands MXNet a DataBatch of host arrays per iteration (utils/detection_input.py:160-177) and its
    metrics read the six loss outputs back (ScalarLoss, rangedet/core/detection_metric.py:200-211).  Here the record is
    uploaded on a copy stream into one of `slots` staging sets (PCIe: ~41 MB per frame), the captured graphs read their
    static buffers (filled device-to-device from the staging set), and the six per-level loss sums come back through
    one pinned 24-byte buffer.  Nothing synchronises the host except `losses()`."""

    def __init__(self, step, slots=2):
        if not step.with_loss:
            raise ValueError("HostFedTrainStep needs a GraphedTrainStep built with the fused loss (with_loss=True)")
        self.step_, self.slots, self.n_fed, self.n_run = step, slots, 0, 0
        dev = step.data.device
        self.names = ["input_data", "coord_s1"] + sorted(step.targets)
        self.static = dict(step.targets, input_data=step.data, coord_s1=step.coord)
        self.staging = [{k: torch.empty_like(self.static[k]) for k in self.names} for _ in range(slots)]
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.ready = [torch.cuda.Event() for _ in range(slots)]       # record i is on the device
        self.consumed = [torch.cuda.Event() for _ in range(slots)]    # its staging set has been read
        self.loss_dev = torch.zeros(6, device=dev)
        self.loss_host = torch.zeros(6).pin_memory()
        self.loss_done = torch.cuda.Event()
        self.h2d_bytes = sum(self.static[k].numel() * self.static[k].element_size() for k in self.names)
        self.d2h_bytes = self.loss_host.numel() * 4

    def feed(self, record):
        if self.n_fed - self.n_run >= self.slots:
            raise RuntimeError("HostFedTrainStep: %d records already in flight" % self.slots)
        s = self.n_fed % self.slots
        self.n_fed += 1
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.consumed[s])
            for k in self.names:
                src = record[k]
                src = torch.from_numpy(src) if not isinstance(src, torch.Tensor) else src
                self.staging[s][k].copy_(src.reshape(self.staging[s][k].shape), non_blocking=True)
            self.ready[s].record(self.copy_stream)

    def step(self):
        if self.n_run >= self.n_fed:
            raise RuntimeError("HostFedTrainStep.step() without a fed record")
        s = self.n_run % self.slots
        self.n_run += 1
        st, cur = self.step_, torch.cuda.current_stream()
        cur.wait_event(self.ready[s])
        stg = self.staging[s]
        torch._foreach_copy_([self.static[k] for k in self.names], [stg[k] for k in self.names])
        self.consumed[s].record(cur)
        if st.capture:
            st.g_fwd.replay()
        else:
            st._fwd()
        st.backward_update()
        lo = st.loss_out
        torch.stack([l["cls_loss"].sum() for l in lo] + [l["reg_loss"].sum() for l in lo], out=self.loss_dev)
        self.loss_host.copy_(self.loss_dev, non_blocking=True)
        self.loss_done.record(cur)
        return self.loss_host

    def losses(self):
        self.loss_done.synchronize()
        return self.loss_host.numpy().copy()
