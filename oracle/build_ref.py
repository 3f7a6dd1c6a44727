"""Build the *reference's own* C++ for the hot path into oracle/_ref/librd_ref.so.

TEST INFRASTRUCTURE ONLY.  Nothing under rangedet_b200/ may import this.

What it does (only when /root/reference is present, i.e. in the build container;
the GPU box just uses the prebuilt oracle/_ref/librd_ref.so that travels with
the snapshot):

  * derives four headers into oracle/_ref/ (git-ignored, never committed) from
    the reference sources where they lie:
      - operator_cxx/contrib/decode_3d_bbox-inl.h : the part between
        ``const float EPS`` and ``template <typename xpu>`` (the two functor
        structs; the MXNet FCompute driver needs an MXNet source tree)
      - operator_cxx/contrib/rotated_iou-inl.h    : same cut
      - operator_cxx/contrib/nms_3d.cu            : the __device__ geometry helpers
        between ``const float EPS`` and the first ``__global__`` kernel, compiled
        for the host with ``#define __device__`` (pins the IoU of NMS3D)
      - operator_cxx/src_cxx/nms.h                : line 1 (``#include
        "overlap.h"``, which drags Eigen) replaced by overlap.h's own non-Eigen
        preamble so ``atan2(float,float)`` resolves exactly as in the original TU
  * compiles oracle/ref_shim.cpp (our C-ABI wrapper around those functors) with
    the reference's flags (operator_cxx/src_cxx/CMakeLists.txt:18-19:
    -O3 -std=c++14 -fext-numeric-literals; no -march=native, no -ffast-math, so
    no FMA contraction).

The reference's build systems are NOT run (contrib/Makefile needs an MXNet tree,
src_cxx/CMakeLists.txt needs Eigen3; neither exists here).
"""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("RD_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(HERE, "_ref")


def _cut(path, start_marker, end_marker):
    with open(path, "r", encoding="utf-8", errors="replace") as f:
        lines = f.readlines()
    s = next(i for i, l in enumerate(lines) if l.startswith(start_marker))
    e = next(i for i, l in enumerate(lines) if l.startswith(end_marker))
    return "".join(lines[s:e])


def have_reference():
    return os.path.isfile(os.path.join(REF, "operator_cxx", "src_cxx", "nms.h"))


def lib_path():
    return os.path.join(OUT, "librd_ref.so")


def build(force=False):
    """Returns the path of librd_ref.so, or None if it cannot be (re)built and
    no prebuilt copy exists."""
    so = lib_path()
    if not have_reference():
        return so if os.path.isfile(so) else None
    shim = os.path.join(HERE, "ref_shim.cpp")
    if (not force and os.path.isfile(so)
            and os.path.getmtime(so) > os.path.getmtime(shim)
            and os.path.getmtime(so) > os.path.getmtime(__file__)):
        return so
    os.makedirs(OUT, exist_ok=True)
    contrib = os.path.join(REF, "operator_cxx", "contrib")
    with open(os.path.join(OUT, "decode_extract.h"), "w") as f:
        f.write(_cut(os.path.join(contrib, "decode_3d_bbox-inl.h"),
                     "const float EPS", "template <typename xpu>"))
    with open(os.path.join(OUT, "riou_extract.h"), "w") as f:
        f.write(_cut(os.path.join(contrib, "rotated_iou-inl.h"),
                     "const float EPS", "template <typename xpu>"))
    # nms_3d.cu: the __device__ helper functions (Point ... iou_bev, iou_normal, lines 29-378) are plain C++ once
    # __device__ is defined away; the kernels / MXNet driver below them need nvcc + an MXNet tree and are not taken
    with open(os.path.join(OUT, "nms3d_extract.h"), "w") as f:
        f.write(_cut(os.path.join(contrib, "nms_3d.cu"), "const float EPS", "__global__ void nms_kernel_3d"))
        # ... and the two kernels themselves (nms_kernel_3d, prepare_output_kernel_3d, :380-468): ref_shim.cpp runs them
        # on the host by defining blockIdx / threadIdx / __shared__ / __syncthreads (block emulation, see there)
        f.write(_cut(os.path.join(contrib, "nms_3d.cu"), "__global__ void nms_kernel_3d", "template <>"))
    with open(os.path.join(REF, "operator_cxx", "src_cxx", "nms.h"), "r",
              encoding="utf-8", errors="replace") as f:
        nms = f.readlines()
    assert nms[0].startswith('#include "overlap.h"'), nms[0]
    preamble = ("#include <stdio.h>\n#include <math.h>\n#include <string.h>\n"
                "#include <assert.h>\n#include <algorithm>\n"
                # overlap.h pulls pybind11/eigen.h -> pybind11/numpy.h (py::array_t); keep numpy.h, drop Eigen
                "#include <pybind11/numpy.h>\nusing namespace std;\n")
    with open(os.path.join(OUT, "nms_extract.h"), "w") as f:
        f.write(preamble + "".join(nms[1:]))
    import pybind11
    cmd = ["g++", "-O3", "-std=c++14", "-fext-numeric-literals", "-Wno-unused-result",
           "-shared", "-fPIC", "-fopenmp",
           "-I", OUT, "-I", sysconfig.get_paths()["include"], "-I", pybind11.get_include(),
           shim, "-o", so]
    subprocess.check_call(cmd)
    return so


if __name__ == "__main__":
    p = build(force="--force" in sys.argv)
    print(p if p else "reference not available and no prebuilt oracle/_ref/librd_ref.so")
