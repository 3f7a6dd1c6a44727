"""Drop-in import shim: makes a RangeDet checkout's own `config/*.py` build THIS package's graphs, unchanged.

The reference's config files (config/rangedet/rangedet_veh_wo_aug_4_18e.py:12-24) import

    from mxnext.complicate import normalizer_factory
    import rangedet.core.detection_metric as metric              (subclasses mx.metric.EvalMetric)
    from rangedet.core.input import LoadRecord, ...               (imports processing_cxx, utils.detection_input -> mx.io)
    from rangedet.symbol.head.builder import RangeRCNN, RangeRpnHead
    from rangedet.symbol.backbone.dla_backbone import DLABackbone

`install(reference_root)` resolves the graph-side names to this package and leaves the control plane (loader
transforms, metrics, config classes) to the checkout's own files:

    rangedet.symbol.head.builder          -> rangedet_b200.symbol.{RangeRCNN, RangeRpnHead}
    rangedet.symbol.backbone.dla_backbone -> rangedet_b200.symbol.DLABackbone
    rangedet.symbol.backbone.meta_kernel  -> rangedet_b200.meta_kernel.MetaKernel
    processing_cxx                        -> rangedet_b200.processing_cxx         (wnms_4c, assign3D_v2, get_point_num)
    mxnext / mxnext.complicate            -> normalizer_factory returning a declarative BatchNorm spec (no mx.sym)
    mxnet                                 -> ONLY when MXNet itself is not importable: the three host-side base classes
                                             the loader / metric files subclass (mxnet_host.py); nothing numeric

    import rangedet_b200.shim as shim
    shim.install("/path/to/RangeDet")
    cfg = importlib.import_module("config.rangedet.rangedet_veh_wo_aug_4_18e")
    pModel = cfg.get_config(is_train=True)[6]        # pModel.train_symbol is a rangedet_b200.symbol.TrainSymbol
    step = pModel.train_symbol.bind(params, optimizer=pOpt.optimizer, world_size=N, allreduce=...)

Nothing here touches `oracle/` (test infrastructure) and nothing computes on the CPU.
"""
import contextlib
import importlib
import os
import sys
import types

from . import mxnet_host, mxnext_complicate

_GRAPH_MODULES = ("rangedet.symbol.head.builder", "rangedet.symbol.backbone.dla_backbone",
                  "rangedet.symbol.backbone.meta_kernel", "processing_cxx", "mxnext", "mxnext.complicate")
_REF_PACKAGES = ("mxnext", "rangedet", "operator_py", "utils", "config")
_state = None


def _real_mxnet():
    try:
        return importlib.import_module("mxnet")
    except Exception:
        return None


def graph_modules():
    """name -> module object for every reference module path this package stands in for."""
    from .. import meta_kernel, processing_cxx, symbol
    builder = types.ModuleType("rangedet.symbol.head.builder")
    builder.__doc__ = "rangedet_b200 stand-in for rangedet/symbol/head/builder.py"
    builder.RangeRCNN, builder.RangeRpnHead = symbol.RangeRCNN, symbol.RangeRpnHead
    backbone = types.ModuleType("rangedet.symbol.backbone.dla_backbone")
    backbone.__doc__ = "rangedet_b200 stand-in for rangedet/symbol/backbone/dla_backbone.py"
    backbone.DLABackbone = symbol.DLABackbone
    mk = types.ModuleType("rangedet.symbol.backbone.meta_kernel")
    mk.__doc__ = "rangedet_b200 stand-in for rangedet/symbol/backbone/meta_kernel.py"
    mk.MetaKernel = meta_kernel.MetaKernel
    mxnext = types.ModuleType("mxnext")
    mxnext.__path__ = []          # a package: `from mxnext.complicate import ...`
    mxnext.complicate = mxnext_complicate
    mxnext.normalizer_factory, mxnext.bn_count = mxnext_complicate.normalizer_factory, mxnext_complicate.bn_count
    return {"rangedet.symbol.head.builder": builder, "rangedet.symbol.backbone.dla_backbone": backbone,
            "rangedet.symbol.backbone.meta_kernel": mk, "processing_cxx": processing_cxx, "mxnext": mxnext,
            "mxnext.complicate": mxnext_complicate}


def install(reference_root, processing_cxx=None):
    """Register the stand-ins in sys.modules and put the checkout on sys.path.  `processing_cxx`: optional
    replacement module for the loader-side `processing_cxx` (default: rangedet_b200.processing_cxx, CUDA).
    Idempotent; `uninstall()` restores the interpreter."""
    global _state
    if _state is not None:
        return _state["modules"]
    if not os.path.isdir(os.path.join(reference_root, "config")):
        raise FileNotFoundError("%s is not a RangeDet checkout (no config/ directory)" % reference_root)
    mods = graph_modules()
    if processing_cxx is not None:
        mods["processing_cxx"] = processing_cxx
    if _real_mxnet() is None:
        mods.update(mxnet_host.modules())
    saved = {}
    for k in list(sys.modules):                      # forget reference modules imported some other way
        if k.split(".")[0] in _REF_PACKAGES or k in mods:
            saved[k] = sys.modules.pop(k)
    sys.modules.update(mods)
    sys.path.insert(0, reference_root)
    _state = dict(root=reference_root, saved=saved, modules=mods)
    # bind every stand-in as an attribute of its (real, empty) parent package from the checkout, so that
    # `import rangedet.symbol.backbone.meta_kernel as mk` resolves like `from ... import ...` does
    for name, mod in mods.items():
        parent, _, leaf = name.rpartition(".")
        if parent and parent not in mods:
            try:
                setattr(importlib.import_module(parent), leaf, mod)
            except ImportError:
                uninstall()
                raise
    return mods


def uninstall():
    global _state
    if _state is None:
        return
    try:
        sys.path.remove(_state["root"])
    except ValueError:
        pass
    for k in list(sys.modules):
        if k.split(".")[0] in _REF_PACKAGES or k in _state["modules"]:
            sys.modules.pop(k, None)
    sys.modules.update(_state["saved"])
    _state = None


@contextlib.contextmanager
def drop_in(reference_root, processing_cxx=None):
    """with shim.drop_in(root): cfg = importlib.import_module("config.rangedet.<name>")"""
    install(reference_root, processing_cxx)
    try:
        yield importlib.import_module
    finally:
        uninstall()


def load_config(reference_root, name, is_train=True):
    """`get_config(is_train)` of config/rangedet/<name>.py from the checkout, graphs built by this package."""
    with drop_in(reference_root) as imp:
        return imp("config.rangedet." + name).get_config(is_train=is_train)
