"""Host-side mirror of rangedet/symbol/backbone/meta_kernel.py:7-240 (class MetaKernel).

Same constructor and the same ``meta_baseline_bias(name, data, coord_data, data_channels,
coord_channels, channel_list, norm, conv1_filter, kernel_size=3)`` signature that
DLABackboneBuilder.meta_kernel_conv selects by string (dla_backbone.py:65-89); instead of
emitting ~10 MXNet symbols it runs the single fused sm_100a kernel.  Parameters are created on
first use under the reference's names (``<name>_<W>_mlp{0,1}_{weight,bias}``, meta_kernel.py:138)
so a checkpoint keyed by those names maps one to one.
"""
import math

import torch

from . import ops


class MetaKernel(object):
    def __init__(self, num_batch, feat_height, feat_width, fp16, num_frame=1, params=None,
                 device="cuda", impl=ops.IMPL_DEFAULT):
        self.num_batch = num_batch
        self.H = feat_height
        self.W = feat_width
        self.fp16 = fp16
        self.num_frame = num_frame
        self.params = params if params is not None else {}
        self.device = device
        self.impl = impl

    def _param(self, key, shape, fan_in=None):
        if key not in self.params:
            if fan_in is None:
                v = torch.zeros(shape, device=self.device)
            else:  # mx.init.Xavier(factor_type="in", rnd_type="gaussian", magnitude=2), tools/train.py:198
                v = torch.randn(shape, device=self.device) * math.sqrt(2.0 / fan_in)
            self.params[key] = torch.nn.Parameter(v)
        return self.params[key]

    def meta_baseline_bias(self, name, data, coord_data, data_channels, coord_channels, channel_list,
                           norm, conv1_filter, kernel_size=3, **kwargs):
        if kernel_size != 3 or coord_channels != 3 or list(channel_list) != [32, data_channels]:
            raise NotImplementedError(
                "rangedet_b200 Meta-Kernel supports kernel_size=3, coord_channels=3, "
                "channel_list=[32, data_channels] (the shipped configs, config :95-103)")
        B, C, H, W = data.shape
        if (H, W) != (self.H, self.W) or C != data_channels:
            raise ValueError("data shape %s does not match MetaKernel(H=%d, W=%d, C=%d)"
                             % (tuple(data.shape), self.H, self.W, data_channels))
        pre = "%s_%d_mlp" % (name, self.W)  # name + '_' then "{W}_mlp{i}" (meta_kernel.py:198,138)
        w0 = self._param(pre + "0_weight", (32, coord_channels, 1, 1), fan_in=coord_channels)
        b0 = self._param(pre + "0_bias", (32,))
        w1 = self._param(pre + "1_weight", (data_channels, 32, 1, 1), fan_in=32)
        b1 = self._param(pre + "1_bias", (data_channels,))
        return ops.meta_kernel(data.float(), coord_data.float(), w0, b0, w1, b1, self.impl)
