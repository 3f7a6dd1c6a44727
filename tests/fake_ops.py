"""CPU emulation of the kernel API (rangedet_b200.ops) in plain torch -- TEST INFRASTRUCTURE ONLY.

Same function names, argument meaning, tensor layouts (zero-haloed NHWC bf16 activations, packed [tap][Cout][Cin]
weights, G[tap][CA][CB] weight gradients, (6,C) BatchNorm coefficient rows ...) as the C-ABI wrappers, so that the HOST
logic of rangedet_b200.train -- the tape, the data-gradient weight layouts, the phase-grouped transposed convolutions,
the tap-major Meta-Kernel unit, the flat-mode gather maps -- can run and be checked on a box without a GPU
(tests/test_train_cpu_emulation.py patches `train.ops` with this module).  Every function computes in fp32 from the
bf16-rounded operands and rounds its bf16 outputs once, like the kernels.  The product never imports this file.
"""
import torch
import torch.nn.functional as F

from rangedet_b200.ops import (ACT_DTYPES, BN_EPS, BN_MOMENTUM, IMPL_DEFAULT, conv_bwdstats_supported, from_nhwc_padded,  # noqa: F401
                               pack_conv_weight, pack_deconv_weight, tap_major_weight, to_nhwc_padded)  # (pure torch helpers of the real module)

bf16 = torch.bfloat16
COMPUTE = torch.float32      # arithmetic type of the emulated kernels


def set_exact(on):
    """Exact mode: activations and arithmetic in float64 (no storage rounding), so the tape can be compared with
    autograd to rounding-free precision; the default mimics the kernels (bf16 storage, fp32 arithmetic)."""
    global bf16, COMPUTE
    bf16, COMPUTE = (torch.float64, torch.float64) if on else (torch.bfloat16, torch.float32)


def _interior(t, halo_w=1):
    return t[:, 1:-1, halo_w:t.shape[2] - halo_w]


def _nchw(x_pad):
    return _interior(x_pad).permute(0, 3, 1, 2).to(COMPUTE)


def _store(out, y_nchw, c_off=0):
    """fp32 NCHW result -> interior of the haloed NHWC bf16 tensor `out` (channels c_off ...)."""
    C = y_nchw.shape[1]
    _interior(out)[..., c_off:c_off + C] = y_nchw.permute(0, 2, 3, 1).to(out.dtype)
    return out


def _epilogue(y, scale, shift, relu, residual_pad):
    if scale is not None:
        y = y * scale.to(COMPUTE).view(1, -1, 1, 1)
    if shift is not None:
        y = y + shift.to(COMPUTE).view(1, -1, 1, 1)
    if residual_pad is not None:
        y = y + _nchw(residual_pad)
    return y.relu() if relu else y


def _unpack_conv(w_packed):
    taps, co, ci = w_packed.shape
    k = 3 if taps == 9 else 1
    return w_packed.to(COMPUTE).reshape(k, k, co, ci).permute(2, 3, 0, 1), k


def conv2d_nhwc(x_pad, w_packed, scale=None, shift=None, relu=False, residual_pad=None, out=None, stride_w=1):
    w, k = _unpack_conv(w_packed)
    y = F.conv2d(_nchw(x_pad), w, stride=(1, stride_w), padding=k // 2)
    if out is None:
        out = torch.zeros((x_pad.shape[0], x_pad.shape[1], y.shape[3] + 2, w.shape[0]), dtype=bf16)
    return _store(out, _epilogue(y, scale, shift, relu, residual_pad))


def conv2d_nhwc_stats(x_pad, w_packed, out=None, stride_w=1, ws=None):
    """Emulation of the fused conv + batch-statistics call: the 'partials' handed to bn_train_finalize are z itself."""
    z = conv2d_nhwc(x_pad, w_packed, relu=False, out=out, stride_w=stride_w)
    return z, z, -1


def conv2d_nhwc_bwdstats(x_pad, w_packed, bn_z_pad, bn_coef, bn_mask_mode, out=None, ws=None):
    """Emulation of the data-gradient conv with fused BatchNorm-backward sums: the 'sums' are recomputed by bn_act_bwd."""
    return conv2d_nhwc(x_pad, w_packed, relu=False, out=out), None, -1


def bn_train_finalize(partial, nslots, N, H, W, C, gamma=None, beta=None, moving_mean=None, moving_var=None, eps=None,
                      momentum=None, coef=None):
    return bn_train_stats(partial, gamma, beta, moving_mean, moving_var)


def conv2d_nhwc_slice(x_pad, w_packed, out, c_off, relu=False, stride_w=1):
    w, k = _unpack_conv(w_packed)
    y = F.conv2d(_nchw(x_pad), w, stride=(1, stride_w), padding=k // 2)
    return _store(out, y.relu() if relu else y, c_off)


def deconv2d_nhwc(x_pad, w_packed, scale=None, shift=None, relu=False, residual_pad=None, out=None):
    """(3,8)/(1,4) pad (1,2); (3,4)/(1,2) pad (1,1); (3,3)/(1,2) pad (1,1) + output_padding 1 (width 2W);
    y = relu?(deconv * scale + shift) + residual."""
    taps, co, ci = w_packed.shape
    kw = taps // 3
    S, pad, op = {8: (4, 2, 0), 4: (2, 1, 0), 3: (2, 1, 1)}[kw]
    w = w_packed.to(COMPUTE).reshape(3, kw, co, ci).permute(3, 2, 0, 1)          # (Cin, Cout, kh, kw)
    y = F.conv_transpose2d(_nchw(x_pad), w, stride=(1, S), padding=(1, pad), output_padding=(0, op))
    y = _epilogue(y, scale, shift, relu, None)
    if residual_pad is not None:
        y = y + _nchw(residual_pad)
    if out is None:
        out = torch.zeros((x_pad.shape[0], x_pad.shape[1], y.shape[3] + 2, co), dtype=bf16)
    return _store(out, y)


def conv2d_wgrad(a_pad, b_pad, ksize, stride_w=1, out=None):
    """G[tap][a][b] = sum_p A[p][a] * B[p*stride + tap - centre][b] (zero halo)."""
    A, B = _nchw(a_pad), _nchw(b_pad)
    N, CA, H, W = A.shape
    Bp = F.pad(B, (1, 1, 1, 1))
    off = 0 if ksize == 3 else 1
    G = []
    for ky in range(ksize):
        for kx in range(ksize):
            Bs = Bp[:, :, ky + off:ky + off + H, kx + off:kx + off + (W - 1) * stride_w + 1:stride_w]
            G.append(torch.einsum("nahw,nbhw->ab", A, Bs))
    G = torch.stack(G)
    if out is not None:
        out.view(G.shape).copy_(G)
        return out.view(G.shape)
    return G


def bn_train_stats(z_pad, gamma=None, beta=None, moving_mean=None, moving_var=None, eps=BN_EPS, momentum=BN_MOMENTUM):
    z = _nchw(z_pad)
    C = z.shape[1]
    mean = z.mean((0, 2, 3))
    var = z.var((0, 2, 3), unbiased=False)
    invstd = 1.0 / torch.sqrt(var + eps)
    g = gamma.to(COMPUTE) if gamma is not None else torch.ones(C, dtype=COMPUTE)
    b = beta.to(COMPUTE) if beta is not None else torch.zeros(C, dtype=COMPUTE)
    a = g * invstd
    if moving_mean is not None:
        moving_mean.mul_(momentum).add_(mean * (1 - momentum))
    if moving_var is not None:
        moving_var.mul_(momentum).add_(var * (1 - momentum))
    return torch.stack([a, b - mean * a, mean, invstd, var, z.sum((0, 2, 3))])


def bn_act_fwd(z_pad, coef, relu=True, res_before=None, res_after=None, out=None):
    y = _nchw(z_pad) * coef[0].view(1, -1, 1, 1) + coef[1].view(1, -1, 1, 1)
    if res_before is not None:
        y = y + _nchw(res_before)
    if relu:
        y = y.relu()
    if res_after is not None:
        y = y + _nchw(res_after)
    return _store(torch.zeros_like(z_pad) if out is None else out, y)


def bn_act_bwd(dy_pad, z_pad, coef, mask_mode, y_mask=None, dz_halo_w=1, dz_out=None, want_g=False, g_out=None, dgb_out=None,
               sums=None):
    g, z = _nchw(dy_pad), _nchw(z_pad)
    a, b, mean, invstd = [coef[i].view(1, -1, 1, 1) for i in range(4)]
    if mask_mode == 1:
        g = g * (_nchw(y_mask) > 0)
    elif mask_mode == 2:
        g = g * ((z * a + b) > 0)
    M = z.shape[0] * z.shape[2] * z.shape[3]
    S1, S2 = g.sum((0, 2, 3)), (g * (z - mean)).sum((0, 2, 3))
    dz = a * (g - (S1 / M).view(1, -1, 1, 1) - (z - mean) * invstd * invstd * (S2 / M).view(1, -1, 1, 1))
    N, Hp, Wp, C = z_pad.shape
    if dz_out is None:
        dz_out = torch.zeros((N, Hp, Wp - 2 + 2 * dz_halo_w, C), dtype=bf16)
    _interior(dz_out, dz_halo_w)[...] = dz.permute(0, 2, 3, 1).to(dz_out.dtype)
    if dgb_out is None:
        dgb_out = torch.empty((2, C))
    dgb_out.view(2, C)[0] = coef[3] * S2
    dgb_out.view(2, C)[1] = S1
    if want_g:
        g_out = _store(torch.zeros_like(z_pad) if g_out is None else g_out, g)
    return dz_out, dgb_out.view(2, C)[0], dgb_out.view(2, C)[1], (g_out if want_g else None)


def channel_sums(x_pad, out=None):
    s = _nchw(x_pad).sum((0, 2, 3))
    if out is not None:
        out.copy_(s)
        return out
    return s


def add_nhwc(x0_pad, x1_pad, out=None):
    return _store(torch.zeros_like(x0_pad) if out is None else out, _nchw(x0_pad) + _nchw(x1_pad))


def copy_channels(src_pad, src_off, dst_pad, dst_off, nchan):
    dst_pad[..., dst_off:dst_off + nchan] = src_pad[..., src_off:src_off + nchan]
    return dst_pad


def _tap_major_to_ref(x, C9):
    """NCHW channels k*C+c -> c*9+k."""
    C = C9 // 9
    return x.reshape(x.shape[0], 9, C, *x.shape[2:]).transpose(1, 2).reshape(x.shape[0], C9, *x.shape[2:])


def _ref_to_tap_major(x, C9):
    C = C9 // 9
    return x.reshape(x.shape[0], C, 9, *x.shape[2:]).transpose(1, 2).reshape(x.shape[0], C9, *x.shape[2:])


def nhwc_to_nchw(src_pad, channels=None, tap_major=False, out=None):
    C = channels or src_pad.shape[3]
    x = _nchw(src_pad)[:, :C]
    x = (_tap_major_to_ref(x, C) if tap_major else x).contiguous()
    if out is not None:
        out.copy_(x)
        return out
    return x


def nchw_to_nhwc(src, out, tap_major=False):
    x = _ref_to_tap_major(src, src.shape[1]) if tap_major else src
    return _store(out, x.to(COMPUTE))


def meta_kernel_forward_nhwc(data, coord, w0, b0, w1, b1, scale, shift, relu=True, out=None, dtype=None):
    from oracle import meta_kernel_ref
    m = meta_kernel_ref.meta_baseline_bias(data, coord, w0.reshape(32, 3), b0, w1.reshape(-1, 32), b1)   # (B, c*9+k, H, W)
    m = m * scale.to(COMPUTE).view(1, -1, 1, 1) + shift.to(COMPUTE).view(1, -1, 1, 1)
    if relu:
        m = m.relu()
    B, C9, H, W = m.shape
    if out is None:
        out = torch.zeros((B, H + 2, W + 2, C9), dtype=bf16)
    return _store(out, _ref_to_tap_major(m, C9))


def meta_kernel_backward(grad_out, data, coord, w0, b0, w1, b1, impl=0):
    from oracle import meta_kernel_ref
    r = meta_kernel_ref.meta_baseline_bias_fwd_bwd(data, coord, w0.reshape(32, 3), b0, w1.reshape(-1, 32), b1, grad_out)
    return r[1], r[2], r[3], r[4], r[5]


def meta_kernel_backward_nhwc(grad_out_pad, data, coord, w0, b0, w1, b1, need_data_grad=True):
    go = nhwc_to_nchw(grad_out_pad, tap_major=True).to(data.dtype)
    r = meta_kernel_backward(go, data, coord, w0, b0, w1, b1)
    return (r[0] if need_data_grad else None,) + tuple(r[1:])


def gather_to_bf16(src, idx, out):
    out.copy_(torch.where(idx >= 0, src[idx.clamp(min=0).long()], torch.zeros((), dtype=src.dtype)).to(out.dtype))
    return out


def gather_f32(src, idx, out):
    out.copy_(torch.where(idx >= 0, src[idx.clamp(min=0).long()], torch.zeros((), dtype=src.dtype)))
    return out


def sgd_mom_update(weight, grad, mom, wd, hyper):
    lr, momentum, rescale, clip = [float(v) for v in hyper[:4]]
    g = grad * rescale
    if clip > 0:
        g = g.clamp(-clip, clip)
    g = g + wd * weight
    mom.mul_(momentum).add_(g, alpha=-lr)
    weight.add_(mom)


def rpn_loss(cls_logit, reg_delta, pc, gt_bbox, mask, reg_target, reg_weight, reg_norm_weight, iou_type="bev", alpha=1.0,
             gamma=2.0, smooth_l1_scalar=3.0, scale_loss_shift=128.0, cls_loss_weight=10.0, reg_loss_weight=8.0,
             want_loss=True, out=None):
    """The fused loss head through its torch restatement (oracle/loss_ref.py), writing into the caller's buffers."""
    from oracle import loss_ref
    r = loss_ref.rpn_loss_level(cls_logit.float(), reg_delta.float(), pc.numpy(), gt_bbox.numpy(), mask.float(), reg_target.float(),
                                reg_weight.float(), reg_norm_weight.float(), iou_type=iou_type, alpha=alpha, gamma=gamma,
                                smooth_l1_scalar=smooth_l1_scalar, scale_loss_shift=scale_loss_shift,
                                cls_loss_weight=cls_loss_weight, reg_loss_weight=reg_loss_weight)
    if out is None:
        return r
    for k, v in r.items():
        if out.get(k) is not None:
            out[k].copy_(v.reshape(out[k].shape))
    return out


def rpn_loss_nhwc(cls_pad, reg_pad, pc, gt_bbox, mask, reg_target, reg_weight, reg_norm_weight, dcls_pad, dreg_pad, out=None, **hyper):
    """Emulation of the loss on the NHWC head tensors: widen, run the planar emulation, round the gradients back in place."""
    r = rpn_loss(nhwc_to_nchw(cls_pad, 1), nhwc_to_nchw(reg_pad, 8), pc, gt_bbox, mask, reg_target, reg_weight, reg_norm_weight, **hyper)
    _store(dcls_pad, r["d_cls"].to(COMPUTE))
    _store(dreg_pad, r["d_reg"].to(COMPUTE))
    if out is not None:
        for k in ("iou_target", "cls_loss", "reg_loss"):
            if out.get(k) is not None:
                out[k].copy_(r[k].reshape(out[k].shape))
        return out
    return {k: r[k] for k in ("iou_target", "cls_loss", "reg_loss")}


def get_sorted_foreground(cls_score, bbox_delta, pc, mask, num_fgs):
    from oracle import sorted_fg_ref
    r = sorted_fg_ref.get_sorted_foreground(cls_score.float().numpy(), bbox_delta.float().numpy(), pc.float().numpy(),
                                            mask.float().numpy(), int(num_fgs))
    return tuple(torch.from_numpy(x) for x in r)


def decode_3d_bbox(bbox_deltas, pc_laser_frame, is_bin=False):
    from oracle import oracle
    return torch.from_numpy(oracle().decode_3d_bbox(bbox_deltas.float().numpy(), pc_laser_frame.float().numpy(), is_bin=is_bin))


def nms3d(boxes, iou_thres, max_keep, normal_iou=False):
    from oracle import oracle
    k, b = oracle().nms3d(boxes.float().numpy(), iou_thres, max_keep, normal_iou)
    return torch.from_numpy(k), torch.from_numpy(b)
