"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol the header
declares, argument validation works without a GPU, and the product fails loudly (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT


def _declared():
    txt = open(os.path.join(ROOT, "include", "rangedet_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(rd_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from rangedet_b200 import _lib, build
    build.build()
    L = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 12
    for n in names:
        assert hasattr(L, n), "librangedet_b200.so does not export %s" % n
    assert sorted(_lib.SIGNATURES) == names  # the ctypes table mirrors the header one to one
    assert _lib.lib().rd_version() == 1


def test_no_cuda_device_fails_loudly():
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from rangedet_b200 import _lib, ops, processing_cxx
    assert _lib.lib().rd_check_device() != 0
    assert "fallback" in _lib.last_error() or "CUDA" in _lib.last_error()
    with pytest.raises(RuntimeError):
        ops.decode_3d_bbox(torch.zeros(1, 4, 8), torch.zeros(1, 4, 3))
    with pytest.raises(RuntimeError):
        processing_cxx.wnms_4c(np.zeros((3, 12), np.float32), 0.1, 0.5, False, 100)
    # empty input returns two empty lists without touching the device (nms.h:464-466)
    assert processing_cxx.wnms_4c(np.zeros((0, 12), np.float32), 0.1, 0.5, False, 100) == ([], [])


def test_argument_validation_without_device():
    from rangedet_b200 import _lib
    L = _lib.lib()
    assert L.rd_rotated_iou(None, None, None, 4, 4, 6, None) != 0
    assert "box_type" in _lib.last_error()
    assert L.rd_meta_kernel_fwd(None, None, None, None, None, None, None, 1, 60, 64, 2656, 0, None) != 0
    assert "multiple of 8" in _lib.last_error()
    assert L.rd_wnms_4c_workspace_bytes(100000) > 100000 * 100
    assert L.rd_meta_kernel_bwd_workspace_bytes(4, 64, 64, 2656) > 0


def test_product_does_not_import_oracle():
    """oracle/ is test infrastructure: nothing under rangedet_b200/ may reference it."""
    pkg = os.path.join(ROOT, "rangedet_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dp, f), errors="replace").read()
                assert "import oracle" not in txt and "from oracle" not in txt and "librd_oracle" not in txt \
                    and "librd_ref" not in txt, os.path.join(dp, f)
