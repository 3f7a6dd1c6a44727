"""The in-situ profiler (rangedet_b200.profiling.OpTimer) wraps every kernel-launching op with a metadata function that
receives the SAME arguments; a new keyword on an op that its metadata function does not accept would only show up on the
GPU box as a failed roofline leg of bench.py.  Host logic: checked here, without a GPU."""
import inspect

from rangedet_b200 import ops, profiling


def test_metadata_functions_accept_the_ops_arguments():
    for name, (family, meta) in profiling.OPS.items():
        assert hasattr(ops, name), "profiling.OPS names %s, which rangedet_b200.ops does not define" % name
        assert family in profiling.BOUND
        real = inspect.signature(getattr(ops, name)).parameters
        msig = inspect.signature(meta).parameters
        has_kw = any(p.kind is inspect.Parameter.VAR_KEYWORD for p in msig.values())
        has_pos = any(p.kind is inspect.Parameter.VAR_POSITIONAL for p in msig.values())
        npos_meta = sum(p.kind is inspect.Parameter.POSITIONAL_OR_KEYWORD for p in msig.values())
        required = [p for p in real.values() if p.default is inspect.Parameter.empty and p.kind is inspect.Parameter.POSITIONAL_OR_KEYWORD]
        optional = [p for p in real.values() if p.default is not inspect.Parameter.empty]
        # arguments without a default travel by position, the others by keyword (that is how train.py / dla.py call them)
        assert has_pos or len(required) <= npos_meta, "%s: metadata function takes fewer positional arguments than the op" % name
        for p in optional:
            assert p.name in msig or has_kw, "%s: metadata function does not accept keyword %r" % (name, p.name)


def test_step_ops_are_all_in_the_table():
    """Every op the training tape calls that launches a kernel has a row (else its time is missing from the table)."""
    import re
    src = open(profiling.__file__.replace("profiling.py", "train.py")).read()
    called = set(re.findall(r"ops\.([a-z_0-9]+)\(", src))
    host_only = {"pack_conv_weight", "pack_deconv_weight", "tap_major_weight", "to_nhwc_padded", "from_nhwc_padded",
                 "conv_bwdstats_supported"}
    missing = sorted(c for c in called - host_only if c not in profiling.OPS)
    assert not missing, missing
