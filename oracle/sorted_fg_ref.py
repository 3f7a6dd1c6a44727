"""oracle/ -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

numpy restatement of GetSortedFGOperator.forward, /root/reference
operator_py/get_sorted_foreground.py:11-40.  MXNet (`nd.topk`, `nd.argsort`) is not installable here ->
"parity unpinned" for the order of EQUAL scores (MXNet's tie order is an implementation detail); for
distinct scores the result is fully determined by the reference's text.  Ties here: ascending point
index (stable argsort), which is what the CUDA path guarantees too.
"""
import numpy as np


def get_sorted_foreground(cls_score, bbox_delta, pc, mask, num_fgs):
    """(B,N), (B,N,8), (B,N,3), (B,N) -> scores (B,K), deltas (B,K,8), pc (B,K,3), K = int(num_fgs)."""
    k = int(num_fgs)                                   # CustomOp kwargs arrive as strings (:51)
    score = (cls_score * mask).astype(np.float32)      # :20
    assert pc.shape[1] >= k                            # infer_shape :66
    B = score.shape[0]
    out_s = np.zeros((B, k), np.float32)               # :27-29
    out_d = np.zeros((B, k, bbox_delta.shape[2]), np.float32)
    out_p = np.zeros((B, k, 3), np.float32)
    for i in range(B):                                 # :31-37
        order = np.argsort(-(score[i] + np.float32(0.0)), kind="stable")[:k]   # topk + argsort(desc)
        out_s[i] = score[i][order]
        out_d[i] = bbox_delta[i][order]
        out_p[i] = pc[i][order]
    return out_s, out_d, out_p
