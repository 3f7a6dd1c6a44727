"""Host-side `mxnet` names for a deployment WITHOUT MXNet: only what the reference's loader and metric files need
at import / bookkeeping time.  Installed by rangedet_b200.shim.install() when `import mxnet` fails; never used
when MXNet is present, and nothing here computes the path.

  rangedet/core/detection_metric.py:5-261   class X(mx.metric.EvalMetric): sum_metric / num_inst bookkeeping
  utils/detection_input.py:19,169-177       class PostMergeBatchLoader(mx.io.DataIter); mx.nd.array; mx.io.DataBatch

Semantics follow the MXNet 1.x Python sources of those classes (python/mxnet/metric.py `EvalMetric`,
python/mxnet/io/io.py `DataDesc` / `DataBatch` / `DataIter`): constructor arguments, attribute names, reset / get /
get_name_value / update_dict, iterator protocol.  `nd.array` returns a numpy array wrapped with `.asnumpy()` /
`.astype()` so metrics written against NDArray keep working on host data.
"""
import types
from collections import namedtuple

import numpy as np


class EvalMetric(object):
    def __init__(self, name, output_names=None, label_names=None, **kwargs):
        self.name = str(name)
        self.output_names = output_names
        self.label_names = label_names
        self._kwargs = kwargs
        self.reset()

    def __str__(self):
        return "EvalMetric: {}".format(dict(self.get_name_value()))

    def update_dict(self, label, pred):
        pred = [pred[n] for n in self.output_names] if self.output_names is not None else list(pred.values())
        label = [label[n] for n in self.label_names] if self.label_names is not None else list(label.values())
        self.update(label, pred)

    def update(self, labels, preds):
        raise NotImplementedError()

    def reset(self):
        self.num_inst = 0
        self.sum_metric = 0.0

    def get(self):
        if self.num_inst == 0:
            return (self.name, float("nan"))
        return (self.name, self.sum_metric / self.num_inst)

    def get_name_value(self):
        name, value = self.get()
        if not isinstance(name, list):
            name, value = [name], [value]
        return list(zip(name, value))


class DataDesc(namedtuple("DataDesc", ["name", "shape"])):
    def __new__(cls, name, shape, dtype=np.float32, layout="NCHW"):
        ret = super(DataDesc, cls).__new__(cls, name, shape)
        ret.dtype, ret.layout = dtype, layout
        return ret


class DataBatch(object):
    def __init__(self, data, label=None, pad=None, index=None, bucket_key=None, provide_data=None, provide_label=None):
        self.data, self.label, self.pad, self.index = data, label, pad, index
        self.bucket_key, self.provide_data, self.provide_label = bucket_key, provide_data, provide_label


class DataIter(object):
    def __init__(self, batch_size=0):
        self.batch_size = batch_size

    def __iter__(self):
        return self

    def reset(self):
        pass

    def next(self):
        raise StopIteration

    def __next__(self):
        return self.next()


class _HostArray(np.ndarray):
    """numpy array answering the two NDArray methods the loader / metrics call (asnumpy, astype returns same type)."""

    def asnumpy(self):
        return np.asarray(self)


def array(source, ctx=None, dtype=None):
    a = np.asarray(source, dtype=dtype if dtype is not None else (np.float32 if not hasattr(source, "dtype") else None))
    return a.view(_HostArray)


def modules():
    mx = types.ModuleType("mxnet")
    mx.__doc__ = "rangedet_b200.shim.mxnet_host: host-side stand-in (MXNet is not installed)"
    mx.__path__ = []
    metric = types.ModuleType("mxnet.metric")
    metric.EvalMetric = EvalMetric
    io = types.ModuleType("mxnet.io")
    io.DataIter, io.DataBatch, io.DataDesc = DataIter, DataBatch, DataDesc
    nd = types.ModuleType("mxnet.ndarray")
    nd.array = array
    mx.metric, mx.io, mx.nd, mx.ndarray = metric, io, nd, nd
    return {"mxnet": mx, "mxnet.metric": metric, "mxnet.io": io, "mxnet.ndarray": nd}
