"""Host-side operator surface of the RangeDet hot path over torch CUDA tensors.

Each function mirrors one reference operator (same argument meaning, shapes and error behaviour)
and calls the sm_100a kernels through the C-ABI (include/rangedet_b200.h).  torch is used only for
device memory and streams.  No function here has a CPU path.

  reference operator                                         here
  ---------------------------------------------------------  -------------------------------
  MetaKernel.meta_baseline_bias (meta_kernel.py:166-240)     meta_kernel / MetaKernelFunction
  mx.sym.contrib.Decode3DBbox   (decode_3d_bbox.cc:15-65)    decode_3d_bbox
  mx.nd.contrib.RotatedIOU      (rotated_iou.cc:12-60)       rotated_iou
  Custom op 'batch_rotated_iou' (batch_rotated_iou.py)       batch_rotated_iou
  processing_cxx.wnms_4c        (pybinding.cpp:8)            rangedet_b200.processing_cxx.wnms_4c
  RangeRpnHead.get_fpn_loss, one level (builder.py:300-422)  rpn_loss
  processing_cxx.assign3D_v2 / get_point_num (assigner.h)    rangedet_b200.processing_cxx.* / *_device
  GenerateTarget.get_rpn_reg_target (input.py:452-506)       rpn_reg_target
"""
import ctypes

import torch

from . import _lib

IMPL_DEFAULT, IMPL_FP32, IMPL_TMA_WS = 0, 1, 3   # (2, the first tcgen05 version, was removed: superseded by 3)


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def _chk(t, name, ndim=None, last=None):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor (rangedet_b200 has no CPU fallback)" % name)
    if t.dtype != torch.float32:
        raise TypeError("%s must be float32, got %s" % (name, t.dtype))
    if ndim is not None and t.dim() != ndim:
        raise ValueError("%s must have %d dims, got shape %s" % (name, ndim, tuple(t.shape)))
    if last is not None and t.shape[-1] not in (last if isinstance(last, tuple) else (last,)):
        raise ValueError("%s last dim must be %s, got shape %s" % (name, last, tuple(t.shape)))
    return t.contiguous()


# ------------------------------------------------------------------------------------------------
# Meta-Kernel
# ------------------------------------------------------------------------------------------------
def meta_kernel_forward(data, coord, w0, b0, w1, b1, impl=IMPL_DEFAULT):
    """(B,C,H,W),(B,3,H,W),(32,3),(32),(C,32),(C) -> (B, 9C, H, W)."""
    data = _chk(data, "data", 4)
    coord = _chk(coord, "coord", 4)
    B, C, H, W = data.shape
    if tuple(coord.shape) != (B, 3, H, W):
        raise ValueError("coord must be (B,3,H,W)=%s, got %s" % ((B, 3, H, W), tuple(coord.shape)))
    w0 = _chk(w0.reshape(w0.shape[0], -1), "w0", 2)
    w1 = _chk(w1.reshape(w1.shape[0], -1), "w1", 2)
    b0, b1 = _chk(b0, "b0", 1), _chk(b1, "b1", 1)
    if tuple(w0.shape) != (32, 3) or b0.shape[0] != 32 or tuple(w1.shape) != (C, 32) or b1.shape[0] != C:
        raise ValueError("MLP parameter shapes must be w0 (32,3) b0 (32) w1 (C,32) b1 (C)")
    out = torch.empty((B, 9 * C, H, W), device=data.device, dtype=torch.float32)
    with torch.cuda.device(data.device):
        st = _lib.lib().rd_meta_kernel_fwd(_p(data), _p(coord), _p(w0), _p(b0), _p(w1), _p(b1), _p(out),
                                           B, C, H, W, int(impl), _stream())
    _lib.check(st, "meta_kernel_forward")
    return out


def meta_kernel_forward_nhwc(data, coord, w0, b0, w1, b1, scale, shift, relu=True, out=None, dtype=torch.bfloat16):
    """Meta-Kernel forward fused with the following per-channel scale/shift (+ReLU), written as haloed NHWC
    bf16 / fp16 (B,H+2,W+2,9C) with tap-major channels k*C+c (storage type = that of `out`, else `dtype`);
    `scale`/`shift` are given in the REFERENCE order (c*9+k, i.e. the BatchNorm(576) parameters of
    dla_backbone.py:93) and permuted here."""
    data = _chk(data, "data", 4)
    coord = _chk(coord, "coord", 4)
    B, C, H, W = data.shape
    w0 = _chk(w0.reshape(w0.shape[0], -1), "w0", 2)
    w1 = _chk(w1.reshape(w1.shape[0], -1), "w1", 2)
    b0, b1 = _chk(b0, "b0", 1), _chk(b1, "b1", 1)
    sc = scale.float().reshape(C, 9).t().contiguous().reshape(-1)   # (c*9+k) -> (k*C+c)
    sh = shift.float().reshape(C, 9).t().contiguous().reshape(-1)
    if out is None:
        out = torch.zeros((B, H + 2, W + 2, 9 * C), device=data.device, dtype=dtype)
    with torch.cuda.device(data.device):
        st = _lib.act_fn("rd_meta_kernel_fwd_nhwc_bf16", out.dtype)(_p(data), _p(coord), _p(w0), _p(b0), _p(w1), _p(b1), _p(sc), _p(sh),
                                                     int(bool(relu)), _p(out), B, C, H, W, _stream())
    _lib.check(st, "meta_kernel_forward_nhwc")
    return out


def tap_major_weight(w_oihw, C=64):
    """Permute the input channels of the aggregation conv weight (Cout, 9C, 1, 1) from c*9+k to k*C+c."""
    co = w_oihw.shape[0]
    return w_oihw.reshape(co, C, 9, *w_oihw.shape[2:]).transpose(1, 2).reshape(co, 9 * C, *w_oihw.shape[2:]).contiguous()


def meta_kernel_backward(grad_out, data, coord, w0, b0, w1, b1, impl=IMPL_DEFAULT):
    """-> (grad_data, grad_w0, grad_b0, grad_w1, grad_b1)."""
    data = _chk(data, "data", 4)
    coord = _chk(coord, "coord", 4)
    B, C, H, W = data.shape
    grad_out = _chk(grad_out, "grad_out", 4)
    if tuple(grad_out.shape) != (B, 9 * C, H, W):
        raise ValueError("grad_out must be (B,9C,H,W)")
    w0 = _chk(w0.reshape(w0.shape[0], -1), "w0", 2)
    w1 = _chk(w1.reshape(w1.shape[0], -1), "w1", 2)
    b0, b1 = _chk(b0, "b0", 1), _chk(b1, "b1", 1)
    L = _lib.lib()
    dev = data.device
    gd = torch.empty_like(data)
    gw0, gb0 = torch.empty((32, 3), device=dev), torch.empty((32,), device=dev)
    gw1, gb1 = torch.empty((C, 32), device=dev), torch.empty((C,), device=dev)
    nbytes = int(L.rd_meta_kernel_bwd_workspace_bytes(B, C, H, W))
    ws = torch.empty((max(nbytes, 4) + 3) // 4, device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        st = L.rd_meta_kernel_bwd(_p(grad_out), _p(data), _p(coord), _p(w0), _p(b0), _p(w1), _p(b1),
                                  _p(gd), _p(gw0), _p(gb0), _p(gw1), _p(gb1), _p(ws),
                                  ctypes.c_size_t(ws.numel() * 4), B, C, H, W, int(impl), _stream())
    _lib.check(st, "meta_kernel_backward")
    return gd, gw0, gb0, gw1, gb1


def meta_kernel_backward_nhwc(grad_out_pad, data, coord, w0, b0, w1, b1, need_data_grad=True):
    """meta_kernel_backward fed the gradient in the layout meta_kernel_forward_nhwc writes: grad_out_pad (B,H+2,W+2,9C)
    zero-haloed bf16 / fp16, tap-major channels k*C+c.  Both kernels read the 2-byte tensor directly (no (B,9C,H,W) fp32
    intermediate); results are bit-identical to meta_kernel_backward on the converted tensor.
    -> (grad_data or None, grad_w0, grad_b0, grad_w1, grad_b1)."""
    data = _chk(data, "data", 4)
    coord = _chk(coord, "coord", 4)
    B, C, H, W = data.shape
    _chk_nhwc(grad_out_pad, "grad_out_pad")
    if tuple(grad_out_pad.shape) != (B, H + 2, W + 2, 9 * C):
        raise ValueError("grad_out_pad must be (B,H+2,W+2,9C)")
    w0 = _chk(w0.reshape(w0.shape[0], -1), "w0", 2)
    w1 = _chk(w1.reshape(w1.shape[0], -1), "w1", 2)
    b0, b1 = _chk(b0, "b0", 1), _chk(b1, "b1", 1)
    L = _lib.lib()
    dev = data.device
    gd = torch.empty_like(data) if need_data_grad else None
    gw0, gb0 = torch.empty((32, 3), device=dev), torch.empty((32,), device=dev)
    gw1, gb1 = torch.empty((C, 32), device=dev), torch.empty((C,), device=dev)
    nbytes = int(L.rd_meta_kernel_bwd_workspace_bytes(B, C, H, W))
    ws = torch.empty((max(nbytes, 4) + 3) // 4, device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        st = _lib.act_fn("rd_meta_kernel_bwd_nhwc_bf16", grad_out_pad.dtype)(
            _p(grad_out_pad), _p(data), _p(coord), _p(w0), _p(b0), _p(w1), _p(b1), _p(gd) if gd is not None else None,
            _p(gw0), _p(gb0), _p(gw1), _p(gb1), _p(ws), ctypes.c_size_t(ws.numel() * 4), B, C, H, W, _stream())
    _lib.check(st, "meta_kernel_backward_nhwc")
    return gd, gw0, gb0, gw1, gb1


class MetaKernelFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, data, coord, w0, b0, w1, b1, impl):
        ctx.save_for_backward(data, coord, w0, b0, w1, b1)
        ctx.impl = impl
        ctx.shapes = (w0.shape, w1.shape)
        return meta_kernel_forward(data, coord, w0, b0, w1, b1, impl)

    @staticmethod
    def backward(ctx, grad_out):
        data, coord, w0, b0, w1, b1 = ctx.saved_tensors
        gd, gw0, gb0, gw1, gb1 = meta_kernel_backward(grad_out, data, coord, w0, b0, w1, b1, ctx.impl)
        return gd, None, gw0.reshape(ctx.shapes[0]), gb0, gw1.reshape(ctx.shapes[1]), gb1, None


def meta_kernel(data, coord, w0, b0, w1, b1, impl=IMPL_DEFAULT):
    """Differentiable Meta-Kernel (gradients to data and the four MLP parameters)."""
    return MetaKernelFunction.apply(data, coord, w0, b0, w1, b1, impl)


class MetaKernelHostPipeline(object):
    """Meta-Kernel forward + backward over HOST (pinned) tensors: the call a host-side framework makes.

    The op moves ~0.87 GB per frame across PCIe in EACH direction and computes for 0.3 ms, so the call is
    a copy pipeline: frames flow through `slots` device buffers on three streams -- host->device, compute
    (rd_meta_kernel_fwd + rd_meta_kernel_bwd per frame), device->host -- so that the uploads of frame
    i+1 overlap the downloads of frame i (PCIe is full duplex).  Parameter gradients are summed over the
    frames on the device, as a batched call would.  Streams keep running across calls: nothing here
    synchronises the host; `wait()` makes the caller's stream wait for everything issued so far.
    """

    def __init__(self, C, H, W, device, slots=2, impl=IMPL_DEFAULT):
        self.C, self.H, self.W, self.dev, self.impl, self.slots = C, H, W, torch.device(device), impl, slots
        dev = self.dev
        mk = lambda *shape: torch.empty(shape, device=dev, dtype=torch.float32)
        self.d_data = [mk(1, C, H, W) for _ in range(slots)]
        self.d_coord = [mk(1, 3, H, W) for _ in range(slots)]
        self.d_go = [mk(1, 9 * C, H, W) for _ in range(slots)]
        self.d_out = [mk(1, 9 * C, H, W) for _ in range(slots)]
        self.d_gd = [mk(1, C, H, W) for _ in range(slots)]
        self.gp = [mk(32, 3), mk(32), mk(C, 32), mk(C)]          # per-frame parameter gradients
        self.gp_sum = mk(32 * 3 + 32 + C * 32 + C)
        L = _lib.lib()
        self.ws = mk((int(L.rd_meta_kernel_bwd_workspace_bytes(1, C, H, W)) + 3) // 4 + 1)
        self.s_in, self.s_comp, self.s_out = (torch.cuda.Stream(dev) for _ in range(3))
        self.ev_in = [torch.cuda.Event() for _ in range(slots)]      # inputs of the slot are on the device
        self.ev_comp = [torch.cuda.Event() for _ in range(slots)]    # kernels of the slot are done
        self.ev_out = [torch.cuda.Event() for _ in range(slots)]     # results of the slot are on the host
        self.ev_gp = torch.cuda.Event()                              # summed parameter gradients are on the host
        self.n = 0

    def __call__(self, h_data, h_coord, h_grad_out, w0, b0, w1, b1, h_out, h_grad_data, h_grad_params):
        """All h_* are pinned CPU tensors: data (B,C,H,W), coord (B,3,H,W), grad_out / out (B,9C,H,W),
        grad_data (B,C,H,W), grad_params (96+32+32C+C,) = [gW0, gb0, gW1, gb1] flattened."""
        for t in (h_data, h_coord, h_grad_out, h_out, h_grad_data, h_grad_params):
            if t.device.type != "cpu" or not t.is_pinned():
                raise ValueError("MetaKernelHostPipeline needs pinned host tensors")
        L, C, H, W = _lib.lib(), self.C, self.H, self.W
        B = h_data.shape[0]
        cur = torch.cuda.current_stream(self.dev)
        self.s_in.wait_stream(cur)      # the caller may just have filled / be done with the host buffers
        self.s_comp.wait_stream(cur)    # parameters produced on the caller's stream
        for b in range(B):
            s = self.n % self.slots
            self.n += 1
            with torch.cuda.stream(self.s_in):
                self.s_in.wait_event(self.ev_comp[s])   # kernels that last read this slot's inputs are done
                self.d_data[s].copy_(h_data[b:b + 1], non_blocking=True)
                self.d_coord[s].copy_(h_coord[b:b + 1], non_blocking=True)
                self.d_go[s].copy_(h_grad_out[b:b + 1], non_blocking=True)
                self.ev_in[s].record(self.s_in)
            with torch.cuda.stream(self.s_comp):
                self.s_comp.wait_event(self.ev_in[s])
                self.s_comp.wait_event(self.ev_out[s])  # previous results of this slot have left the device
                st = L.rd_meta_kernel_fwd(_p(self.d_data[s]), _p(self.d_coord[s]), _p(w0), _p(b0), _p(w1), _p(b1),
                                          _p(self.d_out[s]), 1, C, H, W, int(self.impl), _stream())
                _lib.check(st, "meta_kernel_fwd (host pipeline)")
                st = L.rd_meta_kernel_bwd(_p(self.d_go[s]), _p(self.d_data[s]), _p(self.d_coord[s]), _p(w0), _p(b0), _p(w1),
                                          _p(b1), _p(self.d_gd[s]), _p(self.gp[0]), _p(self.gp[1]), _p(self.gp[2]),
                                          _p(self.gp[3]), _p(self.ws), ctypes.c_size_t(self.ws.numel() * 4), 1, C, H, W,
                                          int(self.impl), _stream())
                _lib.check(st, "meta_kernel_bwd (host pipeline)")
                flat = torch.cat([g.reshape(-1) for g in self.gp])
                if b == 0:
                    self.s_comp.wait_event(self.ev_gp)  # the previous call's sum has been downloaded
                    self.gp_sum.copy_(flat)
                else:
                    self.gp_sum.add_(flat)
                self.ev_comp[s].record(self.s_comp)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(self.ev_comp[s])
                h_out[b:b + 1].copy_(self.d_out[s], non_blocking=True)
                h_grad_data[b:b + 1].copy_(self.d_gd[s], non_blocking=True)
                if b == B - 1:
                    h_grad_params.copy_(self.gp_sum, non_blocking=True)
                    self.ev_gp.record(self.s_out)
                self.ev_out[s].record(self.s_out)

    def wait(self):
        """Make the current stream wait for every copy / kernel issued so far."""
        cur = torch.cuda.current_stream(self.dev)
        cur.wait_stream(self.s_out)
        cur.wait_stream(self.s_comp)
        cur.wait_stream(self.s_in)


# ------------------------------------------------------------------------------------------------
# Decode3DBbox / RotatedIOU / batch_rotated_iou
# ------------------------------------------------------------------------------------------------
def decode_3d_bbox(bbox_deltas, pc_laser_frame, is_bin=False):
    """_contrib_Decode3DBbox: (B,N,8|7) + (B,N,3) -> (B,N,10).  Shape rules as FInferShape
    (decode_3d_bbox.cc:30-60).  Zero gradient (decode_3d_bbox.cc:62): output is detached."""
    d = _chk(bbox_deltas, "bbox_deltas", 3, 7 if is_bin else 8)
    p = _chk(pc_laser_frame, "pc_laser_frame", 3, 3)
    if d.shape[:2] != p.shape[:2]:
        raise ValueError("bbox_deltas and pc_laser_frame must agree on (B,N)")
    out = torch.empty(d.shape[:2] + (10,), device=d.device, dtype=torch.float32)
    with torch.cuda.device(d.device):
        st = _lib.lib().rd_decode_3d_bbox(_p(d), _p(p), _p(out), d.shape[0] * d.shape[1], int(bool(is_bin)),
                                          _stream())
    _lib.check(st, "decode_3d_bbox")
    return out


def rotated_iou(boxes1, boxes2):
    """_contrib_RotatedIOU: (N1,T),(N2,T) -> (N1,N2), T in {5,7,8} (rotated_iou.cc:25-50)."""
    a = _chk(boxes1, "boxes1", 2, (5, 7, 8))
    b = _chk(boxes2, "boxes2", 2, (5, 7, 8))
    if a.shape[1] != b.shape[1]:
        raise ValueError("boxes1 and boxes2 must have the same box type")
    out = torch.empty((a.shape[0], b.shape[0]), device=a.device, dtype=torch.float32)
    with torch.cuda.device(a.device):
        st = _lib.lib().rd_rotated_iou(_p(a), _p(b), _p(out), a.shape[0], b.shape[0], a.shape[1], _stream())
    _lib.check(st, "rotated_iou")
    return out


def batch_rotated_iou(proposal, gt_bbox, iou_type="bev"):
    """Custom op 'batch_rotated_iou' (operator_py/batch_rotated_iou.py): (B,N,10) + (B,G,8|7) ->
    (B,N) max-over-GT IoU target; shape checks as BatchRotatedIOUProp.infer_shape (:84-104)."""
    if iou_type not in ("bev", "3d"):
        raise ValueError("Unknown iou type!")
    pr = _chk(proposal, "proposal", 3, 10)
    gt = _chk(gt_bbox, "gt_bbox", 3, 8 if iou_type == "bev" else 7)
    if pr.shape[0] != gt.shape[0]:
        raise ValueError("proposal and gt_bbox must agree on the batch size")
    out = torch.empty(pr.shape[:2], device=pr.device, dtype=torch.float32)
    with torch.cuda.device(pr.device):
        st = _lib.lib().rd_batch_rotated_iou_max(_p(pr), _p(gt), _p(out), pr.shape[0], pr.shape[1],
                                                 gt.shape[1], 0 if iou_type == "bev" else 1, _stream())
    _lib.check(st, "batch_rotated_iou")
    return out


def rpn_loss(cls_logit, reg_delta, pc, gt_bbox, mask, reg_target, reg_weight, reg_norm_weight, iou_type="bev",
             alpha=1.0, gamma=2.0, smooth_l1_scalar=3.0, scale_loss_shift=128.0, cls_loss_weight=10.0,
             reg_loss_weight=8.0, want_loss=True, out=None):
    """One pyramid level of RangeRpnHead.get_fpn_loss (rangedet/symbol/head/builder.py:300-348): get_iou_target
    (:155-197, Decode3DBbox + 'batch_rotated_iou', stop_gradient) -> get_vfl_loss (:350-379, loss.py:4-30) and
    get_normalize_reg_loss (:381-422), fused.  Hyper-parameters default to config/rangedet/
    rangedet_veh_wo_aug_4_18e.py:36,122-129.

    cls_logit (B,1,H,W)  reg_delta (B,8,H,W)  pc (B,H*W,3)  gt_bbox (B,G,8|7)  mask (B,1,H,W)
    reg_target / reg_weight / reg_norm_weight (B,8,H,W)
    -> dict(iou_target (B,1,H,W), cls_loss (B,1,H,W), reg_loss (B,8,H,W), d_cls, d_reg): the loss tensors are
    the graph outputs `rpn_cls_loss_s*` / `rpn_reg_loss_s*`; d_cls / d_reg are the gradients MakeLoss
    (grad_scale = scale_loss_shift * cls_loss_weight resp. scale_loss_shift) sends into the head outputs.
    `out`: optional dict of preallocated outputs (static addresses for CUDA-graph capture)."""
    if iou_type not in ("bev", "3d"):
        raise ValueError("Unknown iou type!")
    x = _chk(cls_logit, "cls_logit", 4)
    d = _chk(reg_delta, "reg_delta", 4)
    B, _, H, W = d.shape
    N = H * W
    if tuple(x.shape) != (B, 1, H, W) or d.shape[1] != 8:
        raise ValueError("cls_logit must be (B,1,H,W) and reg_delta (B,8,H,W)")
    pc = _chk(pc, "pc", 3, 3)
    gt = _chk(gt_bbox, "gt_bbox", 3, 8 if iou_type == "bev" else 7)
    m = _chk(mask, "mask")
    rt, rw, rn = _chk(reg_target, "reg_target", 4), _chk(reg_weight, "reg_weight", 4), _chk(reg_norm_weight, "reg_norm_weight", 4)
    if tuple(pc.shape[:2]) != (B, N) or gt.shape[0] != B or m.numel() != B * N or any(tuple(t.shape) != (B, 8, H, W) for t in (rt, rw, rn)):
        raise ValueError("rpn_loss: inconsistent shapes")
    dev = d.device
    o = dict(out) if out is not None else {}
    for k, shp in (("iou_target", (B, 1, H, W)), ("cls_loss", (B, 1, H, W)), ("reg_loss", (B, 8, H, W)),
                   ("d_cls", (B, 1, H, W)), ("d_reg", (B, 8, H, W))):
        if k not in o:
            o[k] = torch.empty(shp, device=dev) if (want_loss or k in ("d_cls", "d_reg")) else None
    L = _lib.lib()
    ws = _workspace(int(L.rd_rpn_loss_workspace_bytes()), dev, "rpn_loss")
    nul = ctypes.c_void_p(0)
    pp = lambda t: nul if t is None else _p(t)
    with torch.cuda.device(dev):
        st = L.rd_rpn_loss(_p(x), _p(d), _p(pc), _p(gt), _p(m), _p(rt), _p(rw), _p(rn), B, N, gt.shape[1],
                           0 if iou_type == "bev" else 1, float(alpha), float(gamma), float(smooth_l1_scalar),
                           float(scale_loss_shift * cls_loss_weight), float(reg_loss_weight), float(scale_loss_shift),
                           pp(o["iou_target"]), pp(o["cls_loss"]), pp(o["reg_loss"]), pp(o["d_cls"]), pp(o["d_reg"]),
                           _p(ws), ctypes.c_size_t(ws.numel()), _stream())
    _lib.check(st, "rpn_loss")
    return o


def rpn_loss_nhwc(cls_pad, reg_pad, pc, gt_bbox, mask, reg_target, reg_weight, reg_norm_weight, dcls_pad, dreg_pad,
                  iou_type="bev", alpha=1.0, gamma=2.0, smooth_l1_scalar=3.0, scale_loss_shift=128.0, cls_loss_weight=10.0,
                  reg_loss_weight=8.0, want_loss=True, out=None):
    """rpn_loss on the head tensors as the head convolutions hold them: cls_pad / reg_pad (B,H+2,W+2,cpad) zero-haloed
    NHWC bf16 / fp16 (logit = channel 0, deltas = channels 0..7); the gradients are written into the same channels of
    dcls_pad / dreg_pad (whose other channels and halo must be zero).  The other arguments and the loss outputs are
    those of rpn_loss; -> dict(iou_target, cls_loss, reg_loss)."""
    if iou_type not in ("bev", "3d"):
        raise ValueError("Unknown iou type!")
    _chk_nhwc(cls_pad, "cls_pad")
    for t, n in ((reg_pad, "reg_pad"), (dcls_pad, "dcls_pad"), (dreg_pad, "dreg_pad")):
        _chk_nhwc(t, n, cls_pad)
        if t.shape != cls_pad.shape:
            raise ValueError("rpn_loss_nhwc: %s must have the shape of cls_pad" % n)
    B, Hp, Wp, cpad = cls_pad.shape
    H, W = Hp - 2, Wp - 2
    N = H * W
    pc = _chk(pc, "pc", 3, 3)
    gt = _chk(gt_bbox, "gt_bbox", 3, 8 if iou_type == "bev" else 7)
    m = _chk(mask, "mask")
    rt, rw, rn = _chk(reg_target, "reg_target", 4), _chk(reg_weight, "reg_weight", 4), _chk(reg_norm_weight, "reg_norm_weight", 4)
    if tuple(pc.shape[:2]) != (B, N) or gt.shape[0] != B or m.numel() != B * N or any(tuple(t.shape) != (B, 8, H, W) for t in (rt, rw, rn)):
        raise ValueError("rpn_loss_nhwc: inconsistent shapes")
    dev = cls_pad.device
    o = dict(out) if out is not None else {}
    for k, shp in (("iou_target", (B, 1, H, W)), ("cls_loss", (B, 1, H, W)), ("reg_loss", (B, 8, H, W))):
        if k not in o:
            o[k] = torch.empty(shp, device=dev) if want_loss else None
    L = _lib.lib()
    ws = _workspace(int(L.rd_rpn_loss_workspace_bytes()), dev, "rpn_loss")
    nul = ctypes.c_void_p(0)
    pp = lambda t: nul if t is None else _p(t)
    with torch.cuda.device(dev):
        st = _lib.act_fn("rd_rpn_loss_nhwc_bf16", cls_pad.dtype)(
            _p(cls_pad), _p(reg_pad), H, W, cpad, _p(pc), _p(gt), _p(m), _p(rt), _p(rw), _p(rn), B, gt.shape[1],
            0 if iou_type == "bev" else 1, float(alpha), float(gamma), float(smooth_l1_scalar),
            float(scale_loss_shift * cls_loss_weight), float(reg_loss_weight), float(scale_loss_shift),
            pp(o["iou_target"]), pp(o["cls_loss"]), pp(o["reg_loss"]), _p(dcls_pad), _p(dreg_pad),
            _p(ws), ctypes.c_size_t(ws.numel()), _stream())
    _lib.check(st, "rpn_loss_nhwc")
    return o


def assign3d_v2_device(pc, bbox, bbox_center, bbox_radius, mask, is_in_nlz, max_x, min_x, max_y, min_y, max_z, min_z,
                       max_dist):
    """Device-resident assign3D_v2 (operator_cxx/src_cxx/assigner.h:11-87): pc (N,3), bbox (M,24), bbox_center (M,3),
    bbox_radius (M[,1]), mask (N[,1]), is_in_nlz (N[,1]) CUDA float32 -> (N,) int32 box index, -1 = none."""
    pc = _chk(pc, "pc", 2, 3)
    N = pc.shape[0]
    bbox = _chk(bbox.reshape(-1, 24), "bbox", 2, 24)
    M = bbox.shape[0]
    ctr = _chk(bbox_center.reshape(-1, 3), "bbox_center", 2, 3)
    rad = _chk(bbox_radius.reshape(-1), "bbox_radius", 1)
    mask = _chk(mask.reshape(-1), "mask", 1)
    nlz = _chk(is_in_nlz.reshape(-1), "is_in_nlz", 1)
    if ctr.shape[0] != M or rad.shape[0] != M or mask.shape[0] != N or nlz.shape[0] != N:
        raise ValueError("assign3D_v2: inconsistent shapes")
    out = torch.empty((N,), device=pc.device, dtype=torch.int32)
    with torch.cuda.device(pc.device):
        st = _lib.lib().rd_assign3d_v2(_p(pc), _p(bbox), _p(ctr), _p(rad), _p(mask), _p(nlz), float(max_x), float(min_x),
                                       float(max_y), float(min_y), float(max_z), float(min_z), float(max_dist), N, M,
                                       _p(out), _stream())
    _lib.check(st, "assign3D_v2")
    return out


def get_point_num_device(bbox_inds_each_pt, return_hist=False):
    """Device-resident get_point_num (assigner.h:89-109): (N,) float32 box index per point -> (N,) float32 number of
    points in that box (-1 where the index is negative).  return_hist: also the (500,) int32 per-box counts."""
    inds = _chk(bbox_inds_each_pt.reshape(-1), "bbox_inds_each_pt", 1)
    out = torch.empty_like(inds)
    L = _lib.lib()
    hist = torch.zeros(int(L.rd_get_point_num_workspace_bytes()) // 4, device=inds.device, dtype=torch.int32)
    with torch.cuda.device(inds.device):
        st = L.rd_get_point_num(_p(inds), inds.numel(), _p(out), _p(hist), ctypes.c_size_t(hist.numel() * 4), _stream())
    _lib.check(st, "get_point_num")
    return (out, hist) if return_hist else out


def rpn_reg_target(pc, gt_box7, bbox_ind, point_hist, reg_dim_weight):
    """GenerateTarget for one frame, num_classes == 1 (rangedet/core/input.py:345-372, 430-506): pc (N,3), gt_box7 (M,7)
    [x,y,z,l,w,h,yaw], bbox_ind (N) int32, point_hist (500) int32, reg_dim_weight (8) ->
    (rpn_reg_target, reg_normalize_weight, rpn_reg_weight), each (N,8) float32."""
    pc = _chk(pc, "pc", 2, 3)
    N = pc.shape[0]
    g7 = _chk(gt_box7.reshape(-1, 7), "gt_box7", 2, 7)
    dw = _chk(reg_dim_weight.reshape(-1), "reg_dim_weight", 1)
    if bbox_ind.dtype != torch.int32 or point_hist.dtype != torch.int32 or bbox_ind.numel() != N or dw.numel() != 8:
        raise ValueError("rpn_reg_target: bbox_ind (N) / point_hist must be int32 and reg_dim_weight have 8 entries")
    outs = [torch.empty((N, 8), device=pc.device) for _ in range(3)]
    with torch.cuda.device(pc.device):
        st = _lib.lib().rd_rpn_reg_target(_p(pc), _p(g7), _p(bbox_ind.contiguous()), _p(point_hist.contiguous()), _p(dw), N,
                                          g7.shape[0], _p(outs[0]), _p(outs[1]), _p(outs[2]), _stream())
    _lib.check(st, "rpn_reg_target")
    return tuple(outs)


def wnms_4c_device(dets, thresh, thresh_vote, is_3d=False, hash_scale=100):
    """Device-resident weighted NMS: dets (N,12) CUDA float32 -> (out_dets (K,12), keep_inds (K) int32)."""
    d = _chk(dets, "dets", 2, 12)
    n = d.shape[0]
    if n == 0:
        return (torch.empty((0, 12), device=d.device), torch.empty((0,), device=d.device, dtype=torch.int32))
    L = _lib.lib()
    out = torch.empty((n, 12), device=d.device, dtype=torch.float32)
    keep = torch.empty((n,), device=d.device, dtype=torch.int32)
    nbytes = int(L.rd_wnms_4c_workspace_bytes(n))
    ws = torch.empty(((nbytes + 7) // 8,), device=d.device, dtype=torch.int64)
    cnt = ctypes.c_int(0)
    with torch.cuda.device(d.device):
        st = L.rd_wnms_4c(_p(d), n, float(thresh), float(thresh_vote), int(bool(is_3d)), int(hash_scale),
                          _p(out), _p(keep), ctypes.byref(cnt), _p(ws), ctypes.c_size_t(ws.numel() * 8),
                          _stream())
    _lib.check(st, "wnms_4c")
    k = cnt.value
    return out[:k], keep[:k]


def get_sorted_foreground(cls_score, bbox_delta, pc, mask, num_fgs):
    """operator_py/get_sorted_foreground.py:11-40: (B,N), (B,N,8), (B,N,3), (B,N) ->
    (sorted_fg_score (B,K), sorted_fg_bbox_delta (B,K,8), sorted_fg_pc (B,K,3)), K = num_fgs.
    `num_fgs` may arrive as a string, as CustomOp kwargs do (:51)."""
    num_fgs = int(num_fgs)
    cls_score = _chk(cls_score, "cls_score", 2)
    B, N = cls_score.shape
    bbox_delta = _chk(bbox_delta, "bbox_delta", 3, 8)
    pc = _chk(pc, "pc", 3, 3)
    mask = _chk(mask, "mask", 2)
    if tuple(bbox_delta.shape[:2]) != (B, N) or tuple(pc.shape[:2]) != (B, N) or tuple(mask.shape) != (B, N):
        raise ValueError("get_sorted_foreground: inconsistent shapes")
    L = _lib.lib()
    dev = cls_score.device
    out_s = torch.empty((B, num_fgs), device=dev)
    out_d = torch.empty((B, num_fgs, 8), device=dev)
    out_p = torch.empty((B, num_fgs, 3), device=dev)
    nbytes = int(L.rd_get_sorted_foreground_workspace_bytes(B, N))
    ws = torch.empty(max(nbytes, 4), device=dev, dtype=torch.uint8)
    with torch.cuda.device(dev):
        st = L.rd_get_sorted_foreground(_p(cls_score), _p(bbox_delta), _p(pc), _p(mask), B, N, num_fgs, _p(out_s), _p(out_d),
                                        _p(out_p), _p(ws), ctypes.c_size_t(ws.numel()), _stream())
    _lib.check(st, "get_sorted_foreground")
    return out_s, out_d, out_p


def nms3d(boxes, iou_thres, max_keep, normal_iou=False):
    """_contrib_NMS3D (nms_3d.cc:22-66): boxes (B,N,10) sorted by score -> (keep_idx (B,max_keep) int32 with -1
    fill, boxes_after_nms (B,max_keep,10) with 0 fill)."""
    bx = _chk(boxes, "boxes", 3, 10)
    B, N, _ = bx.shape
    keep = torch.empty((B, max_keep), device=bx.device, dtype=torch.int32)
    out = torch.empty((B, max_keep, 10), device=bx.device, dtype=torch.float32)
    L = _lib.lib()
    nbytes = int(L.rd_nms3d_workspace_bytes(B, N))
    ws = torch.empty((max(nbytes, 16) + 15) // 16 * 4, device=bx.device, dtype=torch.float32)
    with torch.cuda.device(bx.device):
        st = L.rd_nms3d(_p(bx), B, N, float(iou_thres), int(max_keep), int(bool(normal_iou)), _p(keep), _p(out), _p(ws),
                        ctypes.c_size_t(ws.numel() * 4), _stream())
    _lib.check(st, "nms3d")
    return keep, out


def tc_probe_gemm(a, b, mn_major=False):
    """D = A . B^T through tcgen05 (bf16 operands): a (128,k), b (n,k) -> (128,n)."""
    a, b = _chk(a, "a", 2), _chk(b, "b", 2)
    d = torch.empty((128, b.shape[0]), device=a.device, dtype=torch.float32)
    with torch.cuda.device(a.device):
        st = _lib.lib().rd_tc_probe_gemm(_p(a), _p(b), _p(d), b.shape[0], a.shape[1], int(bool(mn_major)), _stream())
    _lib.check(st, "tc_probe_gemm")
    return d


def tma_probe(src, box_w, c0, c1, c2):
    """TMA self-test: returns (tile (64, box_w) loaded at signed coords, dst2 = zeros with the tile stored back)."""
    src = _chk(src, "src", 3)
    C, H, W = src.shape
    dst = torch.zeros((64, box_w), device=src.device)
    dst2 = torch.zeros_like(src)
    with torch.cuda.device(src.device):
        st = _lib.lib().rd_tma_probe(_p(src), _p(dst), _p(dst2), W, H, C, box_w, c0, c1, c2, _stream())
    _lib.check(st, "tma_probe")
    return dst, dst2


# ------------------------------------------------------------------------------------------------
# Convolutions (NHWC bf16, zero-haloed activations)
# ------------------------------------------------------------------------------------------------
ACT_DTYPES = (torch.bfloat16, torch.float16)   # storage types of activations / operands (csrc/act_type.cuh)
BN_EPS = 1e-5 + 1e-10   # mxnext/complicate.py:14
BN_MOMENTUM = 0.9       # mxnext/complicate.py:32-43


def to_nhwc_padded(x_nchw, channels=None, dtype=torch.bfloat16):
    """(N,C,H,W) float -> haloed NHWC bf16 / fp16 (N,H+2,W+2,C') with zero halo (and zero channel padding)."""
    N, C, H, W = x_nchw.shape
    Cp = channels or C
    out = torch.zeros((N, H + 2, W + 2, Cp), device=x_nchw.device, dtype=dtype)
    out[:, 1:H + 1, 1:W + 1, :C] = x_nchw.permute(0, 2, 3, 1).to(dtype)
    return out


def from_nhwc_padded(y_pad, channels=None):
    """haloed NHWC bf16 -> (N,C,H,W) float32 interior."""
    y = y_pad[:, 1:-1, 1:-1, :channels] if channels else y_pad[:, 1:-1, 1:-1, :]
    return y.permute(0, 3, 1, 2).float().contiguous()


def pack_conv_weight(w_oihw, cin=None, cout=None, dtype=torch.bfloat16):
    """(Cout,Cin,kh,kw) -> bf16 [kh*kw][Cout'][Cin'] (zero padded), the layout rd_conv2d_nhwc_bf16 expects.
    (`dtype` other than bf16: used to push element INDICES through the same permutation, see train.py.)"""
    co, ci, kh, kw = w_oihw.shape
    cin, cout = cin or ci, cout or co
    out = torch.zeros((kh * kw, cout, cin), device=w_oihw.device, dtype=dtype)
    out[:, :co, :ci] = w_oihw.permute(2, 3, 0, 1).reshape(kh * kw, co, ci).to(dtype)
    return out


def pack_deconv_weight(w_iohw, cin=None, cout=None, dtype=torch.bfloat16):
    """MXNet/torch transposed-conv weight (Cin,Cout,kh,kw) -> bf16 [kh*kw][Cout'][Cin'] (zero padded)."""
    ci, co, kh, kw = w_iohw.shape
    cin, cout = cin or ci, cout or co
    out = torch.zeros((kh * kw, cout, cin), device=w_iohw.device, dtype=dtype)
    out[:, :co, :ci] = w_iohw.permute(2, 3, 1, 0).reshape(kh * kw, co, ci).to(dtype)
    return out


def _conv_common(x_pad, w_packed, scale, shift, residual_pad, out, w_out, what):
    if x_pad.dtype not in ACT_DTYPES or w_packed.dtype != x_pad.dtype or not x_pad.is_cuda:
        raise TypeError("%s expects CUDA bf16 (or fp16) tensors of one storage type" % what)
    x_pad, w_packed = x_pad.contiguous(), w_packed.contiguous()
    N, Hp, Wp, Cin = x_pad.shape
    taps, Cout, Cin2 = w_packed.shape
    if Cin2 != Cin:
        raise ValueError("weight shape %s does not match input channels %d" % (tuple(w_packed.shape), Cin))
    if out is None:
        out = torch.zeros((N, Hp, w_out + 2, Cout), device=x_pad.device, dtype=x_pad.dtype)
    sc = scale.float().contiguous() if scale is not None else None
    sh = shift.float().contiguous() if shift is not None else None
    if residual_pad is not None:
        residual_pad = residual_pad.contiguous()
        if tuple(residual_pad.shape) != (N, Hp, w_out + 2, Cout) or residual_pad.dtype != x_pad.dtype:
            raise ValueError("residual must be haloed NHWC (storage type of x) at output resolution with Cout channels")
    return x_pad, w_packed, sc, sh, residual_pad, out


def conv2d_nhwc(x_pad, w_packed, scale=None, shift=None, relu=False, residual_pad=None, out=None, stride_w=1):
    """y = relu?(conv(x) * scale + shift + residual); haloed NHWC bf16 in / out; W-stride 1 or 2."""
    N, Hp, Wp, Cin = x_pad.shape
    taps = w_packed.shape[0]
    if taps not in (1, 9):
        raise ValueError("conv2d_nhwc supports 1x1 and 3x3 kernels")
    H, W = Hp - 2, Wp - 2
    x_pad, w_packed, sc, sh, residual_pad, out = _conv_common(x_pad, w_packed, scale, shift, residual_pad, out,
                                                              W // stride_w, "conv2d_nhwc")
    with torch.cuda.device(x_pad.device):
        st = _lib.act_fn("rd_conv2d_nhwc_bf16", x_pad.dtype)(_p(x_pad), _p(w_packed), _p(sc) if sc is not None else None,
                                            _p(sh) if sh is not None else None,
                                            _p(residual_pad) if residual_pad is not None else None, _p(out),
                                            N, H, W, Cin, w_packed.shape[1], 3 if taps == 9 else 1, int(stride_w),
                                            int(bool(relu)), _stream())
    _lib.check(st, "conv2d_nhwc")
    return out


def conv2d_nhwc_stats(x_pad, w_packed, out=None, stride_w=1, ws=None):
    """Training forward of a conv that is followed by BatchNorm: z = conv(x) (raw) AND the partial per-channel sums of
    the stored z (sum, sum of squares), accumulated by the conv epilogue.  Returns (z, partial, nslots) for
    bn_train_finalize.  `ws`: statistics workspace (a private one per layer call is required when several such convs
    are in flight before their finalize; default: a fresh buffer)."""
    N, Hp, Wp, Cin = x_pad.shape
    taps = w_packed.shape[0]
    if taps not in (1, 9):
        raise ValueError("conv2d_nhwc_stats supports 1x1 and 3x3 kernels")
    H, W = Hp - 2, Wp - 2
    x_pad, w_packed, _, _, _, out = _conv_common(x_pad, w_packed, None, None, None, out, W // stride_w, "conv2d_nhwc_stats")
    L = _lib.lib()
    Cout = w_packed.shape[1]
    nb = int(L.rd_bn_workspace_bytes(Cout))
    if ws is None:
        ws = torch.empty(nb // 4, device=x_pad.device, dtype=torch.float32)
    if ws.dtype != torch.float32 or ws.numel() * 4 < nb:
        raise ValueError("conv2d_nhwc_stats: statistics workspace needs %d bytes of float32" % nb)
    slots = ctypes.c_int(0)
    with torch.cuda.device(x_pad.device):
        st = _lib.act_fn("rd_conv2d_nhwc_bf16_stats", x_pad.dtype)(_p(x_pad), _p(w_packed), _p(out), N, H, W, Cin, Cout,
                                                                   3 if taps == 9 else 1, int(stride_w), _p(ws),
                                                                   ctypes.c_size_t(ws.numel() * 4), ctypes.byref(slots), _stream())
    _lib.check(st, "conv2d_nhwc_stats")
    return out, ws, slots.value


def conv_bwdstats_supported(cin, cout, ksize=3, stride_w=1):
    """Whether conv2d_nhwc_bwdstats handles this data-gradient convolution (the transposed-orientation kernel's shapes)."""
    return ksize == 3 and stride_w == 1 and cout == 128 and cin % 64 == 0 and 64 <= cin <= 1024


def conv2d_nhwc_bwdstats(x_pad, w_packed, bn_z_pad, bn_coef, bn_mask_mode, out=None, ws=None):
    """Data-gradient convolution dy = conv3x3(x) whose epilogue also accumulates the sums of the BatchNorm backward of the
    layer below (z = bn_z_pad, coefficients bn_coef, ReLU mask recomputed from z when bn_mask_mode == 2).
    Returns (dy, sums, nslots) for bn_act_bwd(..., sums=(sums, nslots))."""
    N, Hp, Wp, Cin = x_pad.shape
    if w_packed.shape[0] != 9:
        raise ValueError("conv2d_nhwc_bwdstats: 3x3 kernels only")
    H, W = Hp - 2, Wp - 2
    x_pad, w_packed, _, _, _, out = _conv_common(x_pad, w_packed, None, None, None, out, W, "conv2d_nhwc_bwdstats")
    _chk_nhwc(bn_z_pad, "bn_z_pad", x_pad)
    Cout = w_packed.shape[1]
    if tuple(bn_z_pad.shape) != (N, Hp, Wp, Cout):
        raise ValueError("conv2d_nhwc_bwdstats: bn_z_pad must have the output's shape")
    L = _lib.lib()
    nb = int(L.rd_bn_workspace_bytes(Cout))
    if ws is None:
        ws = torch.empty(nb // 4, device=x_pad.device, dtype=torch.float32)
    if ws.dtype != torch.float32 or ws.numel() * 4 < nb:
        raise ValueError("conv2d_nhwc_bwdstats: sums workspace needs %d bytes of float32" % nb)
    slots = ctypes.c_int(0)
    with torch.cuda.device(x_pad.device):
        st = _lib.act_fn("rd_conv2d_nhwc_bf16_bwdstats", x_pad.dtype)(_p(x_pad), _p(w_packed), _p(out), _p(bn_z_pad), _p(bn_coef),
                                                                      int(bn_mask_mode), N, H, W, Cin, Cout, _p(ws),
                                                                      ctypes.c_size_t(ws.numel() * 4), ctypes.byref(slots), _stream())
    _lib.check(st, "conv2d_nhwc_bwdstats")
    return out, ws, slots.value


def bn_train_finalize(partial, nslots, N, H, W, C, gamma=None, beta=None, moving_mean=None, moving_var=None, eps=None,
                      momentum=None, coef=None):
    """Partial sums (conv2d_nhwc_stats) -> coef (6,C) fp32 as bn_train_stats returns; moving statistics updated in place."""
    if coef is None:
        coef = torch.empty((6, C), device=partial.device, dtype=torch.float32)
    opt = lambda t: _p(t) if t is not None else None
    with torch.cuda.device(partial.device):
        st = _lib.lib().rd_bn_train_finalize(_p(partial), int(nslots), N, H, W, C, opt(gamma), opt(beta),
                                             BN_EPS if eps is None else eps, BN_MOMENTUM if momentum is None else momentum,
                                             opt(moving_mean), opt(moving_var), _p(coef), _stream())
    _lib.check(st, "bn_train_finalize")
    return coef


def conv2d_nhwc_slice(x_pad, w_packed, out, c_off, relu=False, stride_w=1):
    """conv(x) written into channels [c_off, c_off + Cout) of the haloed NHWC bf16 tensor `out`."""
    N, Hp, Wp, Cin = x_pad.shape
    taps, Cout, Cin2 = w_packed.shape
    if taps not in (1, 9) or Cin2 != Cin or x_pad.dtype not in ACT_DTYPES or out.dtype != x_pad.dtype or w_packed.dtype != x_pad.dtype:
        raise ValueError("conv2d_nhwc_slice: bad operands")
    H, W = Hp - 2, Wp - 2
    if tuple(out.shape[:3]) != (N, Hp, W // stride_w + 2) or not out.is_contiguous():
        raise ValueError("conv2d_nhwc_slice: output shape %s does not match" % (tuple(out.shape),))
    with torch.cuda.device(x_pad.device):
        st = _lib.act_fn("rd_conv2d_nhwc_bf16_slice", x_pad.dtype)(_p(x_pad.contiguous()), _p(w_packed.contiguous()), None, None, _p(out), N, H,
                                                  W, Cin, Cout, 3 if taps == 9 else 1, int(stride_w), int(bool(relu)),
                                                  out.shape[3], int(c_off), _stream())
    _lib.check(st, "conv2d_nhwc_slice")
    return out


def deconv2d_nhwc(x_pad, w_packed, scale=None, shift=None, relu=False, residual_pad=None, out=None):
    """y = relu?(deconv(x) * scale + shift) + residual for the two agg_stage shapes: weight packed from
    (Cin,Cout,3,8) -> stride (1,4) pad (1,2), or (Cin,Cout,3,4) -> stride (1,2) pad (1,1)."""
    N, Hp, Wp, Cin = x_pad.shape
    taps = w_packed.shape[0]
    if taps not in (24, 12, 9):
        raise ValueError("deconv2d_nhwc supports (3,8)/(1,4), (3,4)/(1,2) and (3,3)/(1,2) kernels")
    kw = taps // 3
    H, W = Hp - 2, Wp - 2
    x_pad, w_packed, sc, sh, residual_pad, out = _conv_common(x_pad, w_packed, scale, shift, residual_pad, out,
                                                              W * (4 if kw == 8 else 2), "deconv2d_nhwc")
    with torch.cuda.device(x_pad.device):
        st = _lib.act_fn("rd_deconv2d_nhwc_bf16", x_pad.dtype)(_p(x_pad), _p(w_packed), _p(sc) if sc is not None else None,
                                              _p(sh) if sh is not None else None,
                                              _p(residual_pad) if residual_pad is not None else None, _p(out),
                                              N, H, W, Cin, w_packed.shape[1], kw, int(bool(relu)), _stream())
    _lib.check(st, "deconv2d_nhwc")
    return out


# ------------------------------------------------------------------------------------------------
# Training path: weight gradient, training-mode BatchNorm + ReLU + residual (forward / backward)
# ------------------------------------------------------------------------------------------------
_WS = {}
_WS_RETIRED = []   # outgrown buffers stay allocated: a captured CUDA graph may have their address baked in


def _workspace(nbytes, device, tag="ws"):
    """Grow-only scratch buffer per (device, tag); the C-ABI never allocates.  A buffer that is outgrown is
    retired, not freed: graphs captured while it was current (train.GraphedTrainStep, dla.GraphedForward) keep
    replaying through its address, so returning it to the caching allocator would let them scribble over
    someone else's memory."""
    key = (str(device), tag)
    t = _WS.get(key)
    if t is None or t.numel() < nbytes:
        if t is not None:
            _WS_RETIRED.append(t)
        t = torch.empty(max(int(nbytes), 1 << 20), device=device, dtype=torch.uint8)
        _WS[key] = t
    return t


def _chk_nhwc(t, name, like=None):
    if t.dtype not in ACT_DTYPES or not t.is_cuda or t.dim() != 4 or not t.is_contiguous():
        raise TypeError("%s must be a contiguous CUDA bf16 / fp16 haloed NHWC tensor" % name)
    if like is not None and t.dtype != like.dtype:
        raise TypeError("%s is stored as %s but its companion tensor as %s" % (name, t.dtype, like.dtype))


def conv2d_wgrad(a_pad, b_pad, ksize, stride_w=1, out=None):
    """G[tap][a][b] = sum_pixels A[p][a] * B[p + tap][b] -> fp32 (ksize*ksize, CA, CB).
    a_pad (N,H+2,W+2,CA) and b_pad (N,H+2,W*stride_w+2,CB): haloed NHWC bf16.  With A = grad of the conv
    output and B = the conv input this is the gradient of the packed forward weight [tap][Cout][Cin]."""
    _chk_nhwc(a_pad, "a_pad")
    _chk_nhwc(b_pad, "b_pad", a_pad)
    N, Hp, Wp, CA = a_pad.shape
    H, W = Hp - 2, Wp - 2
    CB = b_pad.shape[3]
    if tuple(b_pad.shape[:3]) != (N, Hp, W * stride_w + 2):
        raise ValueError("conv2d_wgrad: b_pad shape %s does not match a_pad %s at W-stride %d"
                         % (tuple(b_pad.shape), tuple(a_pad.shape), stride_w))
    L = _lib.lib()
    nb = L.rd_conv2d_wgrad_workspace_bytes(N, H, W, CA, CB, ksize, stride_w)
    if nb == 0:
        raise RuntimeError("rangedet_b200.conv2d_wgrad: %s" % _lib.last_error())
    ws = _workspace(nb, a_pad.device, "wgrad")
    g = out if out is not None else torch.empty((ksize * ksize, CA, CB), device=a_pad.device, dtype=torch.float32)
    if g.dtype != torch.float32 or g.numel() != ksize * ksize * CA * CB or not g.is_contiguous():
        raise ValueError("conv2d_wgrad: out must be a contiguous float32 tensor of %d elements" % (ksize * ksize * CA * CB))
    g = g.view(ksize * ksize, CA, CB)
    with torch.cuda.device(a_pad.device):
        st = _lib.act_fn("rd_conv2d_wgrad_nhwc_bf16", a_pad.dtype)(_p(a_pad), _p(b_pad), _p(g), N, H, W, CA, CB, ksize, stride_w, _p(ws), nb,
                                         _stream())
    _lib.check(st, "conv2d_wgrad")
    return g


def bn_train_stats(z_pad, gamma=None, beta=None, moving_mean=None, moving_var=None, eps=BN_EPS,
                   momentum=BN_MOMENTUM):
    """Batch statistics of a haloed NHWC bf16 tensor -> coef (6,C) fp32: a | b | mean | invstd | var | sum;
    moving_mean / moving_var (fp32, in place) follow MXNet's update."""
    _chk_nhwc(z_pad, "z_pad")
    N, Hp, Wp, C = z_pad.shape
    coef = torch.empty((6, C), device=z_pad.device, dtype=torch.float32)
    L = _lib.lib()
    nb = L.rd_bn_workspace_bytes(C)
    ws = _workspace(nb, z_pad.device, "bn")
    opt = lambda t: _p(t) if t is not None else None
    with torch.cuda.device(z_pad.device):
        st = _lib.act_fn("rd_bn_train_stats_nhwc_bf16", z_pad.dtype)(_p(z_pad), N, Hp - 2, Wp - 2, C, opt(gamma), opt(beta), eps, momentum,
                                           opt(moving_mean), opt(moving_var), _p(coef), _p(ws), nb, _stream())
    _lib.check(st, "bn_train_stats")
    return coef


def bn_act_fwd(z_pad, coef, relu=True, res_before=None, res_after=None, out=None):
    """y = relu?(z*a + b + res_before) + res_after over the interior of haloed NHWC bf16 tensors."""
    _chk_nhwc(z_pad, "z_pad")
    N, Hp, Wp, C = z_pad.shape
    for r in (res_before, res_after):
        if r is not None:
            _chk_nhwc(r, "residual", z_pad)
            if r.shape != z_pad.shape:
                raise ValueError("bn_act_fwd: residual shape %s != %s" % (tuple(r.shape), tuple(z_pad.shape)))
    if out is None:
        out = torch.zeros_like(z_pad)
    opt = lambda t: _p(t) if t is not None else None
    with torch.cuda.device(z_pad.device):
        st = _lib.act_fn("rd_bn_act_fwd_nhwc_bf16", z_pad.dtype)(_p(z_pad), _p(coef), opt(res_before), opt(res_after), _p(out), N, Hp - 2,
                                                Wp - 2, C, int(bool(relu)), _stream())
    _lib.check(st, "bn_act_fwd")
    return out


def bn_act_bwd(dy_pad, z_pad, coef, mask_mode, y_mask=None, dz_halo_w=1, dz_out=None, want_g=False, g_out=None,
               dgb_out=None, sums=None):
    """Backward of bn_act_fwd w.r.t. z, gamma, beta (and the masked gradient g that flows into res_before).
    Returns (dz, dgamma, dbeta, g or None).  dz has a W halo of dz_halo_w pixels.  sums = (partial, nslots) from
    conv2d_nhwc_bwdstats: the reduction pass is skipped."""
    _chk_nhwc(dy_pad, "dy_pad")
    _chk_nhwc(z_pad, "z_pad", dy_pad)
    N, Hp, Wp, C = z_pad.shape
    H, W = Hp - 2, Wp - 2
    if dy_pad.shape != z_pad.shape:
        raise ValueError("bn_act_bwd: dy shape %s != z shape %s" % (tuple(dy_pad.shape), tuple(z_pad.shape)))
    if dz_out is None:
        dz_out = torch.zeros((N, Hp, W + 2 * dz_halo_w, C), device=z_pad.device, dtype=z_pad.dtype)
    if want_g and g_out is None:
        g_out = torch.zeros_like(z_pad)
    if dgb_out is None:
        dgb_out = torch.empty((2, C), device=z_pad.device, dtype=torch.float32)
    dgamma, dbeta = dgb_out.view(2, C)[0], dgb_out.view(2, C)[1]
    L = _lib.lib()
    nb = L.rd_bn_workspace_bytes(C) + 8 * C * 4
    ws = _workspace(nb, z_pad.device, "bn")
    opt = lambda t: _p(t) if t is not None else None
    with torch.cuda.device(z_pad.device):
        if sums is not None:
            st = _lib.act_fn("rd_bn_act_bwd_apply_nhwc_bf16", z_pad.dtype)(
                _p(dy_pad), opt(y_mask), _p(z_pad), _p(coef), int(mask_mode), _p(sums[0]), int(sums[1]), _p(dz_out), int(dz_halo_w),
                opt(g_out), _p(dgamma), _p(dbeta), N, H, W, C, _p(ws), nb, _stream())
        else:
            st = _lib.act_fn("rd_bn_act_bwd_nhwc_bf16", z_pad.dtype)(_p(dy_pad), opt(y_mask), _p(z_pad), _p(coef), int(mask_mode),
                                                                     _p(dz_out), int(dz_halo_w), opt(g_out), _p(dgamma), _p(dbeta),
                                                                     N, H, W, C, _p(ws), nb, _stream())
    _lib.check(st, "bn_act_bwd")
    return dz_out, dgamma, dbeta, g_out


def channel_sums(x_pad, out=None):
    """sums[c] = sum over interior pixels (fp32)."""
    _chk_nhwc(x_pad, "x_pad")
    N, Hp, Wp, C = x_pad.shape
    if out is None:
        out = torch.empty(C, device=x_pad.device, dtype=torch.float32)
    L = _lib.lib()
    nb = L.rd_bn_workspace_bytes(C) + 8 * C * 4
    ws = _workspace(nb, x_pad.device, "bn")
    with torch.cuda.device(x_pad.device):
        st = _lib.act_fn("rd_channel_sums_nhwc_bf16", x_pad.dtype)(_p(x_pad), N, Hp - 2, Wp - 2, C, _p(out), _p(ws), nb, _stream())
    _lib.check(st, "channel_sums")
    return out


def add_nhwc(x0_pad, x1_pad, out=None):
    """y = x0 + x1 over the interior (haloed NHWC bf16)."""
    _chk_nhwc(x0_pad, "x0_pad")
    _chk_nhwc(x1_pad, "x1_pad", x0_pad)
    if x0_pad.shape != x1_pad.shape:
        raise ValueError("add_nhwc: shapes differ")
    N, Hp, Wp, C = x0_pad.shape
    if out is None:
        out = torch.zeros_like(x0_pad)
    with torch.cuda.device(x0_pad.device):
        st = _lib.act_fn("rd_add_nhwc_bf16", x0_pad.dtype)(_p(x0_pad), _p(x1_pad), _p(out), N, Hp - 2, Wp - 2, C, _stream())
    _lib.check(st, "add_nhwc")
    return out


def copy_channels(src_pad, src_off, dst_pad, dst_off, nchan):
    """dst_pad[..., dst_off:dst_off+nchan] = src_pad[..., src_off:src_off+nchan] (haloed NHWC, same N/H/W, halo included):
    channel concatenation into an operand buffer.  Offsets and counts multiples of 8."""
    _chk_nhwc(src_pad, "src_pad")
    _chk_nhwc(dst_pad, "dst_pad", src_pad)
    if src_pad.shape[:3] != dst_pad.shape[:3]:
        raise ValueError("copy_channels: pixel grids differ")
    npix = src_pad.shape[0] * src_pad.shape[1] * src_pad.shape[2]
    with torch.cuda.device(src_pad.device):
        st = _lib.lib().rd_copy_channels_16b(_p(src_pad), src_pad.shape[3], src_off, _p(dst_pad), dst_pad.shape[3], dst_off,
                                             nchan, npix, _stream())
    _lib.check(st, "copy_channels")
    return dst_pad


def nhwc_to_nchw(src_pad, channels=None, tap_major=False, out=None):
    """haloed NHWC bf16 -> (N,C,H,W) fp32 (first `channels` channels); tap_major: NHWC k*(C/9)+c -> NCHW c*9+k."""
    _chk_nhwc(src_pad, "src_pad")
    N, Hp, Wp, Cs = src_pad.shape
    C = channels or Cs
    if out is None:
        out = torch.empty((N, C, Hp - 2, Wp - 2), device=src_pad.device, dtype=torch.float32)
    with torch.cuda.device(src_pad.device):
        st = _lib.act_fn("rd_nhwc_bf16_to_nchw_f32", src_pad.dtype)(_p(src_pad), _p(out), N, Hp - 2, Wp - 2, Cs, C, int(bool(tap_major)), _stream())
    _lib.check(st, "nhwc_to_nchw")
    return out


def nchw_to_nhwc(src, out, tap_major=False):
    """(N,C,H,W) fp32 -> interior of the haloed NHWC bf16 tensor `out` (N,H+2,W+2,Cd), channels [0,C)."""
    _chk_nhwc(out, "out")
    if src.dtype != torch.float32 or not src.is_cuda or not src.is_contiguous():
        raise TypeError("nchw_to_nhwc: src must be a contiguous CUDA float32 tensor")
    N, C, H, W = src.shape
    if tuple(out.shape[:3]) != (N, H + 2, W + 2) or out.shape[3] < C:
        raise ValueError("nchw_to_nhwc: output shape %s does not match %s" % (tuple(out.shape), tuple(src.shape)))
    with torch.cuda.device(src.device):
        st = _lib.act_fn("rd_nchw_f32_to_nhwc_bf16", out.dtype)(_p(src), _p(out), N, H, W, C, out.shape[3], int(bool(tap_major)), _stream())
    _lib.check(st, "nchw_to_nhwc")
    return out


# ------------------------------------------------------------------------------------------------
# Parameter plumbing: flat gathers and the SGD update (optim.cu)
# ------------------------------------------------------------------------------------------------
def gather_to_bf16(src, idx, out):
    """out[i] = bf16 / fp16 (src[idx[i]]) (0 where idx[i] < 0); src fp32 flat, idx int32, out bf16 or fp16, all CUDA."""
    if src.dtype != torch.float32 or idx.dtype != torch.int32 or out.dtype not in ACT_DTYPES or idx.numel() != out.numel():
        raise TypeError("gather_to_bf16: src float32, idx int32 and out bfloat16 / float16 of equal length expected")
    with torch.cuda.device(src.device):
        st = _lib.act_fn("rd_gather_f32_to_bf16", out.dtype)(_p(src), _p(idx), _p(out), idx.numel(), _stream())
    _lib.check(st, "gather_to_bf16")
    return out


def gather_f32(src, idx, out):
    """out[i] = src[idx[i]] (0 where idx[i] < 0); fp32 -> fp32."""
    if src.dtype != torch.float32 or idx.dtype != torch.int32 or out.dtype != torch.float32 or idx.numel() != out.numel():
        raise TypeError("gather_f32: src / out float32 and idx int32 of equal length expected")
    with torch.cuda.device(src.device):
        st = _lib.lib().rd_gather_f32(_p(src), _p(idx), _p(out), idx.numel(), _stream())
    _lib.check(st, "gather_f32")
    return out


def sgd_mom_update(weight, grad, mom, wd, hyper):
    """MXNet sgd_mom_update on flat fp32 buffers, in place (weight, mom).  hyper: CUDA float32 tensor
    [lr, momentum, rescale_grad, clip_gradient]; wd: per-element weight decay."""
    n = weight.numel()
    for t in (weight, grad, mom, wd):
        if t.dtype != torch.float32 or not t.is_cuda or not t.is_contiguous() or t.numel() != n:
            raise TypeError("sgd_mom_update: flat contiguous CUDA float32 buffers of equal length expected")
    if hyper.dtype != torch.float32 or hyper.numel() < 4 or not hyper.is_cuda:
        raise TypeError("sgd_mom_update: hyper must be a CUDA float32 tensor [lr, momentum, rescale, clip]")
    with torch.cuda.device(weight.device):
        st = _lib.lib().rd_sgd_mom_update(_p(weight), _p(grad), _p(mom), _p(wd), _p(hyper), n, _stream())
    _lib.check(st, "sgd_mom_update")
