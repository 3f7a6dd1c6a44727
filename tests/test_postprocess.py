"""Test-time post-processing (rangedet_b200/postprocess.py, mirror of tools/test.py:43-81,178-225)."""
import numpy as np
import pytest

from rangedet_b200 import postprocess, synth


def _dets(n=800, seed=5):
    c10 = synth.boxes7_to_corners10(synth.boxes7(n, seed, clustered=True))
    return c10, synth.distinct_scores(n, seed)


def test_conversions_match_the_reference_functions():
    """Against the reference's own two helper functions executed from tools/test.py (live where /root/reference
    exists) and against synth.corners10_to_dets12, the generator all wNMS vectors are built with."""
    from oracle import ref_py
    c10, score = _dets()
    d11 = postprocess.bbox3d_10dim_to_11dim(c10)
    d12 = np.concatenate([d11, score[:, None]], 1)
    assert np.array_equal(d12, synth.corners10_to_dets12(c10, score))
    d8 = postprocess.bbox3d_12dim_to_8dim(d12)
    b7 = synth.boxes7(800, 5, clustered=True)
    assert np.allclose(d8[:, :6], b7[:, :6], atol=2e-4) and np.array_equal(d8[:, 7], score)     # inverse of the generator
    if not ref_py.available():
        pytest.skip("/root/reference not present")
    f11, f8 = ref_py.test_py_functions()
    assert np.array_equal(f11(c10), d11) and np.array_equal(f8(d12), d8)


@pytest.mark.gpu
def test_frame_detections_end_to_end(orc):
    c10, score = _dets(3000, 7)
    out = postprocess.frame_detections(score, c10, min_score=0.5)
    fg = score > 0.5
    wo, wk = orc.wnms_4c(synth.corners10_to_dets12(c10[fg], score[fg]), 0.1, 0.5, False, 100)
    assert out.shape == (len(wk), 8) and out.dtype == np.float32
    assert np.array_equal(out, postprocess.bbox3d_12dim_to_8dim(wo).astype(np.float32), equal_nan=True)
    assert postprocess.frame_detections(score, c10, min_score=2.0).shape == (0, 8)
    # NMS3D branch: keep_inds select the valid rows (-1 = padding)
    keep = np.full(50, -1)
    keep[:20] = np.arange(20) * 3
    out2 = postprocess.frame_detections(score, c10[keep.clip(0)], keep_inds=keep, min_score=0.0, wnms=False)
    assert out2.shape == (20, 8) and np.array_equal(out2[:, 7], score[keep[:20]])


@pytest.mark.gpu
def test_frame_detections_device_matches_host_path():
    """The device-resident twin (no host hop between the executor and the weighted NMS): same boxes kept, same 8-dim
    detections to float32 rounding (torch's CUDA atan2 vs numpy's for the yaw column)."""
    import torch
    c10, score = _dets(5000, 11)
    host = postprocess.frame_detections(score, c10, min_score=0.4)
    dev = postprocess.frame_detections_device(torch.from_numpy(score).cuda(), torch.from_numpy(c10).cuda(), min_score=0.4)
    assert dev.is_cuda and tuple(dev.shape) == host.shape and host.shape[0] > 10
    assert np.allclose(dev.cpu().numpy(), host, rtol=1e-5, atol=1e-5)
    assert postprocess.frame_detections_device(torch.from_numpy(score).cuda(), torch.from_numpy(c10).cuda(), min_score=2.0).shape == (0, 8)
    with pytest.raises(RuntimeError):
        postprocess.frame_detections_device(torch.from_numpy(score), torch.from_numpy(c10))
