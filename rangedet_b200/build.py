"""Build librangedet_b200.so (sm_100a) in-tree with nvcc.  Cross-compiles without a GPU.

    python -m rangedet_b200.build [--force] [--verbose]

Every .cu under rangedet_b200/csrc is compiled with
    -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17
and linked into rangedet_b200/lib/librangedet_b200.so (cudart linked statically, no libcuda
link dependency, so the library also loads on a CPU-only box for the symbol-export test).
Files that reproduce the reference's branchy fp32 geometry bit-for-bit are compiled with
-fmad=false (no FMA contraction).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "librangedet_b200.so")
OBJDIR = os.path.join(HERE, "build")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
COMMON = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
          "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]
PER_FILE = {
    "decode_iou.cu": ["-fmad=false"],
    "wnms.cu": ["-fmad=false"],
    "nms3d.cu": ["-fmad=false"],
    "rpn_loss.cu": ["-fmad=false"],
    "assign.cu": ["-fmad=false"],
}


# Sources that touch stored activations are compiled twice: as is (bf16 storage) and with -DRD_ACT_F16 (fp16 storage,
# the reference's training format) -- see csrc/act_type.cuh.
DUAL_STORAGE = ("bn_train.cu", "conv_tc.cu", "conv_t.cu", "conv_wgrad.cu", "layout.cu", "optim.cu")


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def compile_units():
    """(source file, object name, extra flags)"""
    units = []
    for src in sources():
        units.append((src, src[:-3] + ".o", []))
        if src in DUAL_STORAGE:
            units.append((src, src[:-3] + "_f16.o", ["-DRD_ACT_F16=1"]))
    return units


def _newer(target, deps):
    if not os.path.isfile(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(d) < t for d in deps)


def build(force=False, verbose=False):
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(HERE, "..", "include", "rangedet_b200.h"))
    headers.append(os.path.abspath(__file__))
    objs = []
    rebuilt = False
    procs = []
    for src, obj, extra in compile_units():
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJDIR, obj)
        objs.append(o)
        if not force and _newer(o, [s] + headers):
            continue
        cmd = [NVCC] + COMMON + PER_FILE.get(src, []) + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
        if verbose:
            print(" ".join(cmd), flush=True)
        procs.append((cmd, subprocess.Popen(cmd)))     # independent translation units: compile them concurrently
        rebuilt = True
    for cmd, pr in procs:
        if pr.wait() != 0:
            raise subprocess.CalledProcessError(pr.returncode, cmd)
    if rebuilt or not os.path.isfile(LIB):
        cmd = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
