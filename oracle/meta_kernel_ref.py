"""CPU restatement of the reference Meta-Kernel graph in plain torch fp32.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  "Parity unpinned": the arithmetic of the
reference lives in MXNet operators (mx.sym.im2col / Convolution / broadcast_minus / elemwise mul),
a third-party dependency (mxnet==2.0.0, requirements.txt:2) that is neither under /root/reference
nor installable here; this file follows the reference graph op for op instead:

  rangedet/symbol/backbone/meta_kernel.py
    :16-38   sampler_im2col  -> F.unfold(x, 3, padding=1)  (same (N, C*kh*kw, L) channel-major
                                 layout as mx.sym.im2col; zero padding, not circular)
    :76-103  relative_coord  -> reshape (B,3,9,H,W) minus centre (B,3,1,H,W)
    :105-164 mlp             -> reshape (B,3,9H,W); conv1x1 3->32 +bias; ReLU; conv1x1 32->64 +bias;
                                 reshape (B,64,9,H,W)     (use_norm=False, no_bias=False :209-215)
    :217-239 data im2col (B,64,9,H,W) * weights -> reshape (B,576,H,W)

  rangedet/symbol/backbone/dla_backbone.py:91-97 (meta_kernel_conv tail):
    BN(576) -> ReLU -> conv1x1 576->filter (no bias) -> BN -> ReLU
"""
import torch
import torch.nn.functional as F


def meta_baseline_bias(data, coord, w0, b0, w1, b1, kernel_size=3):
    """data (B,C,H,W), coord (B,3,H,W); w0 (32,3) b0 (32) w1 (C,32) b1 (C) -> (B, C*9, H, W).

    out[b, c*9+k, h, w] = data[b,c,h+dy,w+dx] * MLP(coord[b,:,h+dy,w+dx] - coord[b,:,h,w])[c]
    with k = ky*3+kx, (dy,dx) = (ky-1,kx-1), zero padding on both im2col's.
    """
    B, C, H, W = data.shape
    cc = coord.shape[1]
    kk = kernel_size * kernel_size
    pad = (kernel_size - 1) // 2
    coord_sample = F.unfold(coord, kernel_size, padding=pad)            # (B, 3*9, HW)   :200-203
    rel = coord_sample.reshape(B, cc, kk, H, W) - coord.unsqueeze(2)     # :204-208
    x = rel.reshape(B, cc, kk * H, W)                                    # mlp reshape :126-134
    x = F.conv2d(x, w0.reshape(w0.shape[0], cc, 1, 1), b0)               # mlp0 :136-145
    x = F.relu(x)                                                        # :151-153
    x = F.conv2d(x, w1.reshape(w1.shape[0], w0.shape[0], 1, 1), b1)      # mlp1
    weights = x.reshape(B, w1.shape[0], kk, H, W)                        # :154-163
    data_sample = F.unfold(data, kernel_size, padding=pad).reshape(B, C, kk, H, W)  # :217-229
    out = data_sample * weights                                          # :231
    return out.reshape(B, C * kk, H, W)                                  # :232-239


def meta_baseline_bias_fwd_bwd(data, coord, w0, b0, w1, b1, grad_out):
    """Returns (out, grad_data, grad_w0, grad_b0, grad_w1, grad_b1) via autograd (coord is a
    data input of the graph: grad_req null, builder.py:20-37)."""
    data = data.detach().clone().requires_grad_(True)
    ps = [p.detach().clone().requires_grad_(True) for p in (w0, b0, w1, b1)]
    out = meta_baseline_bias(data, coord, *ps)
    out.backward(grad_out)
    return (out.detach(), data.grad) + tuple(p.grad for p in ps)


def batch_norm_train(x, gamma, beta, eps=1e-5 + 1e-10):
    """mx.sym.BatchNorm(fix_gamma=False, use_global_stats=False) in training mode
    (mxnext/complicate.py:32-43: eps=1e-5+1e-10, momentum 0.9): biased batch variance."""
    return F.batch_norm(x, None, None, gamma, beta, training=True, momentum=0.0, eps=eps)


def meta_kernel_conv_tail(meta_out, g0, be0, wagg, g1, be1):
    """dla_backbone.py:91-97: BN(576) -> ReLU -> conv1x1 (no bias) -> BN -> ReLU."""
    x = F.relu(batch_norm_train(meta_out, g0, be0))
    x = F.conv2d(x, wagg.reshape(wagg.shape[0], wagg.shape[1], 1, 1))
    return F.relu(batch_norm_train(x, g1, be1))
