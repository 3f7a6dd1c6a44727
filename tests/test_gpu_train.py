"""GPU parity tests of the training path of the convolution family (run on the B200 with `-m gpu`):
weight gradient (tcgen05, MN-major operands), data gradient of the W-strided convolutions, training-mode
BatchNorm + ReLU + residual forward / backward, and the whole backbone + head forward + backward
against a torch-fp32 autograd restatement (oracle/dla_train_ref.py).  All through the C-ABI.

Tolerances: the kernels take bf16 operands and accumulate in fp32, the torch reference computes in fp32
on the same bf16-rounded operands, so single kernels agree to fp32 summation-order noise (1e-4 of the
largest value) plus one bf16 rounding (2^-8 relative) where the output is bf16."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from rangedet_b200 import ops as o
    return o


def _bf(x):
    return x.to(torch.bfloat16).float()


def _maxrel(a, b):
    return float((a.double() - b.double()).abs().max() / max(float(b.double().abs().max()), 1e-30))


# (N, CA, CB, H, W, ksize, stride_w)
WGRAD_SHAPES = [(1, 64, 64, 3, 128, 3, 1), (2, 64, 64, 5, 300, 3, 1), (2, 128, 128, 4, 200, 3, 1),
                (1, 64, 128, 3, 130, 3, 1), (1, 128, 64, 2, 257, 3, 1), (2, 128, 64, 3, 150, 3, 2),
                (1, 64, 128, 4, 140, 1, 2), (1, 64, 576, 3, 300, 1, 1), (1, 128, 512, 3, 166, 3, 1),
                (1, 128, 256, 2, 100, 1, 1), (4, 128, 128, 64, 664, 3, 1)]


def _wgrad_ref(A, B, ksize, s):
    N, CA, H, W = A.shape
    Bp = F.pad(B, (1, 1, 1, 1))
    off = 0 if ksize == 3 else 1
    out = []
    for ky in range(ksize):
        for kx in range(ksize):
            Bs = Bp[:, :, ky + off:ky + off + H, kx + off:kx + off + (W - 1) * s + 1:s]
            out.append(torch.einsum("nahw,nbhw->ab", A.double(), Bs.double()))
    return torch.stack(out).float()


@pytest.mark.parametrize("noswz", [0, 1], ids=["sw128", "noswz"])
@pytest.mark.parametrize("shape", WGRAD_SHAPES)
def test_conv2d_wgrad_vs_torch(ops, shape, noswz):
    N, CA, CB, H, W, ks, s = shape
    if noswz and N * H * W > 100000:
        pytest.skip("diagnostic layout: small shapes only")
    g = torch.Generator(device="cuda").manual_seed(hash(shape) % 1000)
    A = _bf(torch.randn((N, CA, H, W), device="cuda", generator=g))
    B = _bf(torch.randn((N, CB, H, W * s), device="cuda", generator=g))
    os.environ["RD_WGRAD_NOSWZ"] = str(noswz)
    try:
        got = ops.conv2d_wgrad(ops.to_nhwc_padded(A), ops.to_nhwc_padded(B), ks, s)
        torch.cuda.synchronize()
    finally:
        os.environ.pop("RD_WGRAD_NOSWZ", None)
    want = _wgrad_ref(A, B, ks, s)
    assert got.shape == want.shape
    err = _maxrel(got, want)
    assert err < 1e-4, (shape, noswz, err)
    # deterministic: the split partials are added in fixed order
    got2 = ops.conv2d_wgrad(ops.to_nhwc_padded(A), ops.to_nhwc_padded(B), ks, s)
    assert torch.equal(got, got2)


def test_conv2d_wgrad_is_gradient_of_conv(ops):
    """<dz, conv(x; w)> is linear in w: G must equal torch autograd's weight gradient."""
    N, Ci, Co, H, W = 2, 64, 128, 4, 200
    g = torch.Generator(device="cuda").manual_seed(3)
    x = _bf(torch.randn((N, Ci, H, W), device="cuda", generator=g))
    dz = _bf(torch.randn((N, Co, H, W), device="cuda", generator=g))
    w = torch.zeros((Co, Ci, 3, 3), device="cuda", requires_grad=True)
    F.conv2d(x, w, padding=1).backward(dz)
    got = ops.conv2d_wgrad(ops.to_nhwc_padded(dz), ops.to_nhwc_padded(x), 3, 1)  # [tap][Co][Ci]
    want = w.grad.permute(2, 3, 0, 1).reshape(9, Co, Ci)
    assert _maxrel(got, want) < 1e-4


@pytest.mark.parametrize("shape", [(1, 64, 64, 3, 130), (2, 128, 128, 4, 83), (1, 128, 64, 2, 300), (1, 64, 128, 3, 64)])
def test_strided_conv_dgrad_vs_torch(ops, shape):
    """kw = 3 transposed convolution = data gradient of the 3x3 W-stride-2 convolutions (and, with a
    centre-only kernel, of the 1x1 W-stride-2 projections)."""
    N, Cz, Cx, H, W = shape  # dz has Cz channels at width W; dx has Cx channels at width 2W
    g = torch.Generator(device="cuda").manual_seed(7)
    dz = _bf(torch.randn((N, Cz, H, W), device="cuda", generator=g))
    w = _bf(torch.randn((Cz, Cx, 3, 3), device="cuda", generator=g) * 0.1)  # conv weight (Cout=Cz, Cin=Cx)
    want = F.conv_transpose2d(dz, w, stride=(1, 2), padding=(1, 1), output_padding=(0, 1))
    got = ops.from_nhwc_padded(ops.deconv2d_nhwc(ops.to_nhwc_padded(dz), ops.pack_deconv_weight(w)))
    assert got.shape == want.shape
    assert float((got - want).abs().max()) <= 2 ** -7 * float(want.abs().max()) + 1e-2
    # cross-check against autograd of the strided conv itself
    x = torch.zeros((N, Cx, H, 2 * W), device="cuda", requires_grad=True)
    F.conv2d(x, w, stride=(1, 2), padding=1).backward(dz)
    assert _maxrel(want, x.grad) < 1e-5


@pytest.mark.parametrize("C", [64, 128])
@pytest.mark.parametrize("mode", ["plain", "res_before", "res_after", "norelu"])
def test_bn_act_fwd_bwd_vs_torch(ops, C, mode):
    N, H, W = 2, 6, 333
    g = torch.Generator(device="cuda").manual_seed(11)
    z = _bf(torch.randn((N, C, H, W), device="cuda", generator=g) * 2 + 0.5)
    rb = _bf(torch.randn((N, C, H, W), device="cuda", generator=g)) if mode == "res_before" else None
    ra = _bf(torch.randn((N, C, H, W), device="cuda", generator=g)) if mode == "res_after" else None
    dy = _bf(torch.randn((N, C, H, W), device="cuda", generator=g))
    gamma = torch.rand(C, device="cuda", generator=g) + 0.5
    beta = torch.randn(C, device="cuda", generator=g) * 0.3
    relu = mode != "norelu"
    mm, mv = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda")

    zp = ops.to_nhwc_padded(z)
    coef = ops.bn_train_stats(zp, gamma, beta, mm, mv)
    y = ops.bn_act_fwd(zp, coef, relu=relu, res_before=None if rb is None else ops.to_nhwc_padded(rb),
                       res_after=None if ra is None else ops.to_nhwc_padded(ra))
    assert float(y[:, 0].float().abs().max()) == 0 and float(y[:, :, 0].float().abs().max()) == 0  # halo untouched

    # torch reference (fp32 autograd, batch statistics, biased variance)
    zr = z.clone().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    rbr = rb.clone().requires_grad_(True) if rb is not None else None
    mean = zr.mean((0, 2, 3), keepdim=True)
    var = zr.var((0, 2, 3), unbiased=False, keepdim=True)
    u = (zr - mean) / torch.sqrt(var + ops.BN_EPS) * gr[None, :, None, None] + br[None, :, None, None]
    if rbr is not None:
        u = u + rbr
    yr = torch.relu(u) if relu else u
    if ra is not None:
        yr = yr + ra
    yr.backward(dy)

    assert _maxrel(coef[2], mean.flatten()) < 1e-5 and _maxrel(coef[4], var.flatten()) < 1e-4
    assert _maxrel(mm, 0.1 * mean.flatten().detach()) < 1e-5
    assert _maxrel(mv, 0.9 + 0.1 * var.flatten().detach()) < 1e-4
    assert float((ops.from_nhwc_padded(y) - yr).abs().max()) <= 2 ** -8 * float(yr.abs().max()) + 1e-6

    mask_mode = 0 if not relu else (2 if ra is not None else 1)
    dz, dgamma, dbeta, gout = ops.bn_act_bwd(ops.to_nhwc_padded(dy), zp, coef, mask_mode, y_mask=y,
                                             want_g=rb is not None)
    # mask decisions can differ from the fp32 reference only where |pre-activation| is within bf16 rounding of 0
    assert _maxrel(dgamma, gr.grad) < 2e-3 and _maxrel(dbeta, br.grad) < 2e-3
    assert _maxrel(ops.from_nhwc_padded(dz), zr.grad) < 1e-2
    if rb is not None:
        assert _maxrel(ops.from_nhwc_padded(gout), rbr.grad) < 1e-2
    # phase-grouped output (W halo of 4 pixels): same values, shifted columns, zero halo
    dz4, _, _, _ = ops.bn_act_bwd(ops.to_nhwc_padded(dy), zp, coef, mask_mode, y_mask=y, dz_halo_w=4)
    assert torch.equal(dz4[:, :, 4:-4], dz[:, :, 1:-1])
    assert float(dz4[:, :, :4].float().abs().max()) == 0 and float(dz4[:, :, -4:].float().abs().max()) == 0


def test_channel_sums_and_add(ops):
    g = torch.Generator(device="cuda").manual_seed(5)
    a = _bf(torch.randn((2, 64, 5, 200), device="cuda", generator=g))
    b = _bf(torch.randn((2, 64, 5, 200), device="cuda", generator=g))
    ap, bp = ops.to_nhwc_padded(a), ops.to_nhwc_padded(b)
    assert _maxrel(ops.channel_sums(ap), a.sum((0, 2, 3))) < 1e-5
    s = ops.add_nhwc(ap, bp)
    assert torch.equal(ops.from_nhwc_padded(s), _bf(a + b))
    assert float(s[:, 0].float().abs().max()) == 0
