"""rangedet_b200 -- B200-native (sm_100a) kernels for the RangeDet range-image hot path.

Submodules: ops (torch-tensor operator surface), meta_kernel (MetaKernel mirror),
processing_cxx (wnms_4c drop-in), synth (synthetic inputs), build (nvcc build of the C-ABI .so).
"""
__version__ = "0.1.0"
