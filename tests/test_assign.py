"""Training-target assignment: rd_assign3d_v2 / rd_get_point_num / rd_rpn_reg_target vs the
restatement of operator_cxx/src_cxx/assigner.h:11-109 (oracle/rd_oracle.cpp) and rangedet/core/input.py:430-506
(oracle/target_ref.py).  Integer outputs bit-exact; float targets |d| <= 1e-5*|ref| + 1e-5 (libm ulp differences;
the two signed-square-root offsets are compared before the root)."""
import numpy as np
import pytest
import torch

from conftest import golden
from rangedet_b200 import synth

REG_W = [1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0]


def _brute_force_ind(pc, corners, mask):
    """Independent geometric definition: a point is in a box iff strictly between bottom and top and inside the
    footprint rectangle (all four half-plane tests positive); first box wins; the reference's pre-filters restated."""
    from oracle import target_ref
    center, radius, ext, max_dist = target_ref.assigner_args(corners)
    g = corners.reshape(-1, 8, 3)
    out = np.full(pc.shape[0], -1, np.int32)
    for i, p in enumerate(pc):
        if mask[i] < 0.5 or not (ext[1] <= p[0] <= ext[0] and ext[3] <= p[1] <= ext[2] and ext[5] <= p[2] <= ext[4]):
            continue
        d = ((center - p) ** 2).sum(1)
        if d.min() > max_dist:
            continue
        for j, b in enumerate(g):
            A, B, C, D, E = b[0], b[1], b[2], b[3], b[4]
            if d[j] > radius[j] or not (A[2] < p[2] < E[2]):
                continue
            if ((A - B)[:2] @ (p - B)[:2] > 0 and (C - B)[:2] @ (p - B)[:2] > 0 and (A - D)[:2] @ (p - D)[:2] > 0
                    and (C - D)[:2] @ (p - D)[:2] > 0):
                out[i] = j
                break
    return out


def test_assign_restatement_matches_geometric_definition():
    from oracle import target_ref
    pc, mask, b7, c24 = synth.assign_frame(n_vehicles=6, seed=1, h=8, w=200)
    got = target_ref.bbox3d_ind(pc, c24, mask)
    want = _brute_force_ind(pc.astype(np.float64), c24.astype(np.float64), mask)
    # the float64 definition can differ from the fp32 comparisons only for points within rounding of a face
    assert (got != want).mean() < 2e-3
    assert (got >= 0).sum() > 20 and len(set(np.unique(got))) >= 4


def test_assign_restatement_matches_golden():
    from oracle import target_ref
    g = golden("assign.npz")
    pc, mask, b7, c24 = synth.assign_frame(n_vehicles=12, seed=2, h=16, w=400)
    ind = target_ref.bbox3d_ind(pc, c24, mask)
    assert np.array_equal(ind, g["ind"])
    assert np.array_equal(target_ref.normalization_weight(ind), g["norm_w"])
    assert np.allclose(target_ref.rpn_reg_target(pc, b7, ind), g["target"], rtol=1e-6, atol=1e-6)


def test_get_point_num_and_weights_semantics():
    from oracle import target_ref, oracle
    ind = np.array([-1, 0, 2, 2, -1, 2, 0], np.int32)
    assert oracle().get_point_num(ind.astype(np.float32)).tolist() == [-1, 2, 3, 3, -1, 3, 2]
    w = target_ref.normalization_weight(ind)
    assert np.allclose(w, [0, 0.5, 1 / 3, 1 / 3, 0, 1 / 3, 0.5])
    rw = target_ref.rpn_reg_weight(ind, REG_W)
    assert rw.shape == (7, 8) and rw[0].sum() == 0 and rw[1].sum() == 8
    # decode(encode) round trip: the regression target is the inverse of Decode3DBbox (decode_3d_bbox-inl.h:186-274)
    pc, mask, b7, c24 = synth.assign_frame(n_vehicles=5, seed=3, h=8, w=160)
    ind2 = target_ref.bbox3d_ind(pc, c24, mask)
    t = target_ref.rpn_reg_target(pc, b7, ind2)
    sel = ind2 >= 0
    dec = oracle().decode_3d_bbox(t[sel][None], pc[sel][None])[0]
    want = synth.boxes7_to_corners10(b7[ind2[sel]])
    assert np.allclose(dec, want, atol=2e-3)


# ---- GPU parity -------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(64, 2650, 30, 0), (64, 2650, 300, 1), (7, 33, 3, 2)])
def test_assign_and_targets_match_restatement(shape):
    from oracle import target_ref
    from rangedet_b200 import ops
    h, w, nv, seed = shape
    pc, mask, b7, c24 = synth.assign_frame(n_vehicles=nv, seed=seed, h=h, w=w)
    center, radius, ext, max_dist = target_ref.assigner_args(c24)
    cu = lambda a: torch.from_numpy(np.ascontiguousarray(a, np.float32)).cuda()
    ind = ops.assign3d_v2_device(cu(pc), cu(c24), cu(center), cu(radius), cu(mask), torch.zeros(pc.shape[0], device="cuda"),
                                 *ext, max_dist)
    want = target_ref.bbox3d_ind(pc, c24, mask)
    assert np.array_equal(ind.cpu().numpy(), want)          # bit-exact box indices
    assert (want >= 0).sum() > 0
    num, hist = ops.get_point_num_device(ind.float(), return_hist=True)
    from oracle import oracle
    assert np.array_equal(num.cpu().numpy(), oracle().get_point_num(want.astype(np.float32)))
    tgt, nw, rw = ops.rpn_reg_target(cu(pc), cu(b7), ind, hist, cu(np.asarray(REG_W)))
    rt = target_ref.rpn_reg_target(pc, b7, want)
    got = tgt.cpu().numpy()
    assert np.allclose(got[:, 2:], rt[:, 2:], rtol=1e-5, atol=1e-5)
    # dx, dy are signed square roots: infinitely steep at 0, so they are compared before the root (|d| <= 1e-5 m)
    assert np.allclose(got[:, :2] * np.abs(got[:, :2]), rt[:, :2] * np.abs(rt[:, :2]), rtol=1e-5, atol=1e-5)
    assert np.allclose(got[:, :2], rt[:, :2], atol=5e-3)
    assert np.array_equal(nw.cpu().numpy(), np.tile(target_ref.normalization_weight(want)[:, None], (1, 8)))
    assert np.array_equal(rw.cpu().numpy(), target_ref.rpn_reg_weight(want, REG_W))


@pytest.mark.gpu
def test_processing_cxx_assign_mirrors_and_edges():
    from oracle import target_ref
    from rangedet_b200 import processing_cxx
    pc, mask, b7, c24 = synth.assign_frame(n_vehicles=8, seed=4, h=16, w=300)
    center, radius, ext, max_dist = target_ref.assigner_args(c24)
    nlz = np.zeros((pc.shape[0], 1), np.float32)
    nlz[::7] = 1.0                                           # no-label-zone points are never assigned
    out = processing_cxx.assign3D_v2(pc, c24, center, radius.reshape(-1, 1), mask.reshape(-1, 1), nlz, *ext, max_dist)
    assert out.dtype == np.int32 and out.shape == (pc.shape[0], 1)
    from oracle import oracle
    want = oracle().assign3d_v2(pc, c24, center, radius, mask, nlz, *ext, max_dist)
    assert np.array_equal(out.reshape(-1), want) and (out[::7] == -1).all()
    n = processing_cxx.get_point_num(out.astype(np.float32))
    assert n.shape == (pc.shape[0], 1) and np.array_equal(n.reshape(-1), oracle().get_point_num(want.astype(np.float32)))
    assert processing_cxx.get_point_num(np.zeros((0, 1), np.float32)).shape == (0, 1)
    with pytest.raises(ValueError):
        processing_cxx.assign3D_v2(pc[:, :2], c24, center, radius, mask, nlz, *ext, max_dist)


REG_W_SHIPPED = [3, 1, 1, 1, 1, 1, 1, 1]    # config/rangedet/rangedet_veh_wo_aug_4_18e.py:219


def _loader_case():
    pc, mask, b7, c24 = synth.assign_frame(n_vehicles=30, seed=0)     # one full 64x2650 frame
    return pc, mask, b7, c24


def test_target_restatement_matches_the_reference_loader_code():
    """rangedet/core/input.py Bbox3dAssigner.apply + GenerateTarget.apply, imported UNMODIFIED (oracle/ref_py.py;
    processing_cxx = the C++ restatement, so the index assignment itself stays unpinned, but its ARGUMENTS -- radius,
    centres, GT extent, max_dist, input.py:296-322 -- and all of the numpy target arithmetic are the reference's):
    committed golden from that run, plus the live run where /root/reference exists.  Bit-exact."""
    from oracle import ref_py, target_ref
    pc, mask, b7, c24 = _loader_case()
    ind = target_ref.bbox3d_ind(pc, c24, mask)
    sel = ind >= 0
    tgt = target_ref.rpn_reg_target(pc, b7, ind)
    nw = target_ref.normalization_weight(ind)
    rw = target_ref.rpn_reg_weight(ind, REG_W_SHIPPED)
    g = golden("loader_targets.npz")
    assert np.array_equal(ind, g["ind"]) and sel.sum() > 1000
    assert np.array_equal(tgt[sel], g["target_fg"]) and not tgt[~sel].any()
    assert np.array_equal(nw[sel], g["norm_fg"]) and np.array_equal(rw[sel], g["weight_fg"])
    if not ref_py.available():
        pytest.skip("/root/reference not present: golden vectors only")
    r_ind, r_tgt, r_nw, r_rw = ref_py.loader_targets(pc.reshape(64, 2650, 3), mask.reshape(64, 2650, 1), c24.reshape(-1, 8, 3),
                                                     b7, REG_W_SHIPPED)
    assert np.array_equal(r_ind, ind)
    assert np.array_equal(r_tgt.reshape(-1, 8), tgt)
    assert np.array_equal(r_nw.reshape(-1, 8), np.tile(nw[:, None], (1, 8)))
    assert np.array_equal(r_rw.reshape(-1, 8), rw)
