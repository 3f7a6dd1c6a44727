"""Host-side logic of the training step, CPU only (no kernels): the index maps that let ONE gather launch re-pack
every bf16 operand / collect every gradient must reproduce the per-tensor packing functions exactly; the SGD
restatement follows MXNet's sgd_mom_update; the synthetic roidb record has the graph-input names and shapes of
rangedet/symbol/head/builder.py:20-37 / config/rangedet/rangedet_veh_wo_aug_4_18e.py:367-378."""
import numpy as np
import pytest
import torch

from rangedet_b200 import synth, train


@pytest.mark.parametrize("kind,shape,args", [
    ("fwd", (64, 8, 3, 3), (64, 64, 1)),
    ("fwd", (128, 72, 3, 3), (128, 128, 1)),
    ("fwd", (8, 128, 1, 1), (128, 64, 1)),
    ("fwd_tapmajor", (64, 576, 1, 1), (576, 64, 1)),
    ("dgrad", (128, 64, 3, 3), (64, 128, 1)),
    ("dgrad_tapmajor", (64, 576, 1, 1), (576, 64, 1)),
    ("dgrad_s2", (128, 64, 3, 3), (64, 128, 1)),
    ("dgrad_s2", (128, 64, 1, 1), (64, 128, 1)),
    ("deconv_fwd", (128, 64, 3, 8), (128, 64, 1)),
    ("deconv_dgrad", (128, 64, 3, 8), (128, 64, 4)),
    ("deconv_dgrad", (64, 64, 3, 4), (64, 64, 2)),
])
def test_operand_gather_map_reproduces_packing(kind, shape, args):
    g = torch.Generator().manual_seed(0)
    flat = torch.randn(100 + int(np.prod(shape)) + 50, generator=g)
    off = 100
    w = flat[off:off + int(np.prod(shape))].view(shape)
    ci_p, co_p, S = args
    want = train.pack_operand(w, kind, ci_p, co_p, S)                     # bf16, per-tensor path
    ix = train.pack_operand(train._index_like(w, off), kind, ci_p, co_p, S, dtype=torch.float64)
    m = train._to_map(ix).long()
    got = torch.where(m >= 0, flat[m.clamp(min=0)], torch.zeros(())).to(torch.bfloat16).view(want.shape)
    assert torch.equal(got, want)
    assert int((m >= 0).sum()) >= int(np.prod(shape)) or kind.startswith("dgrad_s2") or kind.startswith("deconv")
    assert set(m[m >= 0].tolist()) <= set(range(off, off + int(np.prod(shape))))


def test_deconv_weight_gradient_view_is_the_inverse_of_the_phase_grouped_packing():
    """train.TrainGraph.deconv_bn extracts dW (Cin,Cout,3,KW) from G[ty][tx][ci][ph][co]; pushing indices through
    pack_operand('deconv_dgrad') and back must give each weight element exactly once."""
    for KW, S in ((8, 4), (4, 2)):
        ci, co = 128, 64
        w = torch.arange(ci * co * 3 * KW, dtype=torch.float64).view(ci, co, 3, KW) + 1
        G = train.pack_operand(w, "deconv_dgrad", 128, 64, S, dtype=torch.float64).reshape(3, 3, 128, S, 64)
        pad = KW // 4
        cols = []
        for kx in range(KW):
            tx, ph = divmod(kx - pad + S, S)
            cols.append(G[:, tx, :ci, ph, :co].permute(1, 2, 0))
        assert torch.equal(torch.stack(cols, 3), w)


def test_sgd_restatement_is_mxnet_sgd_mom_update():
    P = {"a_weight": torch.tensor([1.0, -2.0]), "a_bias": torch.tensor([0.5])}
    G = {"a_weight": torch.tensor([128.0 * 3, -128.0 * 100]), "a_bias": torch.tensor([128.0])}
    mom = {}
    train.sgd_momentum_step(P, G, mom, lr=0.1, momentum=0.9, wd=0.01, clip_gradient=35.0, rescale_grad=1 / 128)
    # g = clip(g/128, 35) + wd*w (weights only); m = -lr*g; w += m
    assert torch.allclose(P["a_weight"], torch.tensor([1.0 - 0.1 * (3 + 0.01), -2.0 - 0.1 * (-35 - 0.02)]))
    assert torch.allclose(P["a_bias"], torch.tensor([0.5 - 0.1 * 1.0]))
    train.sgd_momentum_step(P, {k: torch.zeros_like(v) for k, v in G.items()}, mom, lr=0.1, momentum=0.9, wd=0.0)
    assert torch.allclose(mom["a_bias"], torch.tensor([0.9 * -0.1]))


def test_synthetic_roidb_record_names_and_shapes():
    B = 2
    T = synth.rpn_targets(B, seed=1, n_vehicles=5, h=16, w=250, w_pad=256)
    assert T["gt_bbox_veh_for_iou_pred"].shape == (B, 200, 8)
    assert np.allclose(T["gt_bbox_veh_for_iou_pred"][:, 5:, 3:7], 1e-3) and not T["gt_bbox_veh_for_iou_pred"][:, 5:, :3].any()
    for s in (1, 2, 4):
        for k in ("rpn_reg_target", "rpn_reg_weight", "reg_normalize_weight"):
            assert T["%s_s%d" % (k, s)].shape == (B, 8, 16, 256 // s), (k, s)
        assert T["range_image_mask_s%d" % s].shape == (B, 1, 16, 256 // s)
        assert T["pc_vehicle_frame_s%d" % s].shape == (B, 16 * 256 // s, 3)
    w = T["rpn_reg_weight_s1"]
    assert 0 < (w > 0).mean() < 0.5 and not T["rpn_reg_target_s1"][w == 0].any()
    assert not T["range_image_mask_s1"][..., 250:].any()            # PadData columns carry no points
