"""Stand-in for the part of mxnext/complicate.py the config files touch: `normalizer_factory` (:14-95).

In the reference the factory returns a closure that emits mx.sym.BatchNorm (local), SyncBatchNorm, a frozen BN
or GroupNorm into the symbol graph.  This package's graph is not composed of symbols: the BatchNorm of every conv
layer is part of the fused layer kernels (rangedet_b200/train.py, dla.py), so the factory returns a declarative
`Normalizer` carrying the same settings; symbol.DLABackbone / RangeRpnHead check them against what the kernels
implement (per-GPU batch statistics, eps 1e-5 + 1e-10, momentum 0.9 -- config/rangedet/*.py:56 `normalizer_factory(
type="local")`) and refuse anything else.
"""
__all__ = ["normalizer_factory", "bn_count", "Normalizer"]

bn_count = [0]     # mxnext/complicate.py:11


class Normalizer(object):
    def __init__(self, type, ndev, eps, mom):
        self.type, self.ndev, self.eps, self.mom = type, ndev, eps, mom

    def __call__(self, *args, **kwargs):
        raise RuntimeError("rangedet_b200: the normalizer is fused into the layer kernels and cannot be applied to an "
                           "MXNet symbol; build the graph with rangedet_b200.symbol (see rangedet_b200.shim)")

    def __repr__(self):
        return "Normalizer(type=%r, ndev=%r, eps=%r, mom=%r)" % (self.type, self.ndev, self.eps, self.mom)


def normalizer_factory(type="local", ndev=None, eps=1e-5 + 1e-10, mom=0.9):
    """Same signature and defaults as mxnext/complicate.py:14; a pre-constructed normalizer passes through (:23-24)."""
    if callable(type):
        return type
    if type not in ("local", "localbn", "fix", "fixbn", "sync", "syncbn", "hvd", "hvd_syncbn", "in", "gn"):
        raise KeyError("Unknown norm type {}".format(type))     # the factory's own error (:148-149)
    return Normalizer(type, ndev, eps, mom)
