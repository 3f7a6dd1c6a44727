"""Synthetic inputs with the shapes / statistics of the reference's data path.

Pure numpy; shared by tests/ and bench.py.  Layout sources in the reference:
  * range image 64x2650 padded to 2656, normalised coords: config/rangedet/
    rangedet_veh_wo_aug_4_18e.py:38,42,257-267 and rangedet/core/input.py:200-213,522-544
  * 10-dim decoded box  [Ax,Ay,Bx,By,Cx,Cy,Dx,Dy,z0,z1]: operator_cxx/contrib/decode_3d_bbox-inl.h:244-274
  * 12-dim wNMS det     [8 corners, yaw, z0, h, score]:   tools/test.py:56-81,209
  * fixed-length GT padding [0,0,0,e,e,e,e,0], e=1e-3:     rangedet/core/input.py:264-265
"""
import numpy as np

H_RANGE = 64
W_RANGE = 2650
W_PADDED = 2656


def range_image_coords(batch, seed=0, h=H_RANGE, w=W_RANGE, w_pad=W_PADDED, missing=0.10):
    """(B,3,h,w_pad) float32 normalised xyz of a synthetic LiDAR sweep; columns w..w_pad-1 zero."""
    rng = np.random.default_rng(seed)
    incl = np.linspace(-0.31, 0.04, h, dtype=np.float64)[:, None]
    azim = np.linspace(np.pi, -np.pi, w, dtype=np.float64)[None, :]
    out = np.zeros((batch, 3, h, w_pad), np.float32)
    for b in range(batch):
        r = rng.uniform(2.0, 75.0, size=(h, w))
        x = r * np.cos(incl) * np.cos(azim)
        y = r * np.cos(incl) * np.sin(azim)
        z = r * np.sin(incl)
        # per-axis mean / std of the order used by NormData (config :257-267)
        xyz = np.stack([(x - 0.0) / 25.0, (y - 0.0) / 25.0, (z - 1.0) / 2.0], 0)
        hole = rng.uniform(size=(h, w)) < missing
        xyz[:, hole] = 0.0
        out[b, :, :, :w] = xyz.astype(np.float32)
    return out


def feature_map(batch, channels, seed=1, h=H_RANGE, w=W_RANGE, w_pad=W_PADDED):
    rng = np.random.default_rng(seed)
    out = np.zeros((batch, channels, h, w_pad), np.float32)
    out[..., :w] = rng.standard_normal((batch, channels, h, w), dtype=np.float32)
    return out


def meta_mlp_params(seed=2, coord_channels=3, hidden=32, out_channels=64):
    """Xavier(in, gaussian, 2) like tools/train.py:198; biases small non-zero so they are exercised."""
    rng = np.random.default_rng(seed)
    w0 = (rng.standard_normal((hidden, coord_channels)) * np.sqrt(2.0 / coord_channels)).astype(np.float32)
    b0 = (rng.standard_normal(hidden) * 0.1).astype(np.float32)
    w1 = (rng.standard_normal((out_channels, hidden)) * np.sqrt(2.0 / hidden)).astype(np.float32)
    b1 = (rng.standard_normal(out_channels) * 0.1).astype(np.float32)
    return w0, b0, w1, b1


def boxes7(n, seed=0, clustered=False):
    """(n,7) float32 [cx,cy,cz,l,w,h,yaw] vehicles."""
    rng = np.random.default_rng(seed)
    if not clustered:
        cx = rng.uniform(-75, 75, n)
        cy = rng.uniform(-75, 75, n)
        yaw = rng.uniform(-np.pi, np.pi, n)
    else:
        per = 50
        nc = (n + per - 1) // per
        ccx = np.repeat(rng.uniform(-75, 75, nc), per)[:n]
        ccy = np.repeat(rng.uniform(-75, 75, nc), per)[:n]
        cyaw = np.repeat(rng.uniform(-np.pi, np.pi, nc), per)[:n]
        cx = ccx + rng.normal(0, 0.3, n)
        cy = ccy + rng.normal(0, 0.3, n)
        yaw = cyaw + rng.normal(0, 0.1, n)
    cz = rng.uniform(-1, 2, n)
    l = rng.uniform(3.5, 5.5, n)
    w = rng.uniform(1.6, 2.2, n)
    h = rng.uniform(1.4, 2.0, n)
    return np.stack([cx, cy, cz, l, w, h, yaw], 1).astype(np.float32)


def boxes7_to_corners10(b7):
    """[cx,cy,cz,l,w,h,yaw] -> 10-dim corner box with the decode convention (A,B,C,D, z0, z1)."""
    b7 = np.asarray(b7, np.float32)
    cx, cy, cz, l, w, h, yaw = [b7[:, i] for i in range(7)]
    s, c = np.sin(yaw), np.cos(yaw)
    out = np.empty((b7.shape[0], 10), np.float32)
    for k, (sx, sy) in enumerate([(0.5, -0.5), (-0.5, -0.5), (-0.5, 0.5), (0.5, 0.5)]):
        x, y = sx * l, sy * w
        out[:, 2 * k] = x * c - y * s + cx
        out[:, 2 * k + 1] = x * s + y * c + cy
    out[:, 8] = cz - h / 2
    out[:, 9] = out[:, 8] + h
    return out


def corners10_to_dets12(c10, scores):
    """tools/test.py:56-81 (10 -> 11 dims) + score column (:209)."""
    c10 = np.asarray(c10, np.float32)
    yaw = np.arctan2(c10[:, 1] - c10[:, 3], c10[:, 0] - c10[:, 2]).astype(np.float32)
    return np.concatenate([c10[:, :8], yaw[:, None], c10[:, 8:9], (c10[:, 9:10] - c10[:, 8:9]),
                           np.asarray(scores, np.float32)[:, None]], 1).astype(np.float32)


def distinct_scores(n, seed=0):
    """Random permutation of {1..n}/n: all distinct (nms.h:791 uses an unstable sort)."""
    rng = np.random.default_rng(seed + 7919)
    return (rng.permutation(n).astype(np.float64) + 1.0).astype(np.float32) / np.float32(n)


def wnms_dets(n, seed=0, clustered=True):
    return corners10_to_dets12(boxes7_to_corners10(boxes7(n, seed, clustered)), distinct_scores(n, seed))


def gt_boxes8(batch, n_real=50, n_total=200, seed=3):
    """(B,200,8) BEV corner GT, first n_real real, rest padded per input.py:264-265."""
    out = np.zeros((batch, n_total, 8), np.float32)
    eps = np.float32(1e-3)
    out[:, :, 3:7] = eps
    for b in range(batch):
        out[b, :n_real] = boxes7_to_corners10(boxes7(n_real, seed + b))[:, :8]
    return out


def decode_inputs(batch, n, seed=4):
    """bbox_deltas (B,N,8) + pc_laser_frame (B,N,3) in the value ranges the head produces."""
    rng = np.random.default_rng(seed)
    delta = np.empty((batch, n, 8), np.float32)
    delta[..., 0:2] = rng.normal(0, 1.2, (batch, n, 2))
    delta[..., 2] = rng.normal(np.log(1.9), 0.15, (batch, n))
    delta[..., 3] = rng.normal(np.log(4.5), 0.15, (batch, n))
    ang = rng.uniform(-np.pi, np.pi, (batch, n))
    delta[..., 4] = np.cos(ang)
    delta[..., 5] = np.sin(ang)
    delta[..., 6] = rng.uniform(-2, 1, (batch, n))
    delta[..., 7] = rng.normal(np.log(1.7), 0.1, (batch, n))
    r = rng.uniform(2, 75, (batch, n))
    az = rng.uniform(-np.pi, np.pi, (batch, n))
    pc = np.stack([r * np.cos(az), r * np.sin(az), rng.uniform(-2, 3, (batch, n))], -1).astype(np.float32)
    return delta, pc


def encode_box8(b7, px, py):
    """Inverse of the non-bin decode (operator_cxx/contrib/decode_3d_bbox-inl.h:186-235): regression target
    [dx, dy, log w, log l, cos, sin, z0, log h] of box b7 = [cx,cy,cz,l,w,h,yaw] seen from point (px,py)."""
    cx, cy, cz, l, w, h, yaw = [np.asarray(b7[..., i], np.float64) for i in range(7)]
    az = np.arctan2(py, px)
    ox, oy = cx - px, cy - py
    rx = ox * np.cos(az) + oy * np.sin(az)
    ry = -ox * np.sin(az) + oy * np.cos(az)
    cols = [np.sign(rx) * np.sqrt(np.abs(rx)), np.sign(ry) * np.sqrt(np.abs(ry)), np.log(w), np.log(l),
            np.cos(yaw - az), np.sin(yaw - az), cz - h / 2, np.log(h)]
    out = np.stack(np.broadcast_arrays(*cols), -1)
    return out.astype(np.float32)


def rpn_targets(batch, seed=5, n_vehicles=30, strides=(1, 2, 4), h=H_RANGE, w=W_RANGE, w_pad=W_PADDED, n_gt=200,
                missing=0.10):
    """Synthetic roidb record(s) with the post-transform shapes and names of the training graph's inputs
    (config/rangedet/rangedet_veh_wo_aug_4_18e.py:367-378, rangedet/core/input.py:452-506,561-624):
    ~n_vehicles vehicles per frame whose points are assigned to them.
      rpn_reg_target_s{s}, rpn_reg_weight_s{s}, reg_normalize_weight_s{s}  (B,8,h,w_pad/s)
      range_image_mask_s{s} (B,1,h,w_pad/s)   pc_vehicle_frame_s{s} (B,h*w_pad/s,3)
      gt_bbox_veh_for_iou_pred (B,200,8), padded per input.py:264-265."""
    rng = np.random.default_rng(seed)
    incl = np.linspace(-0.31, 0.04, h, dtype=np.float64)[:, None]
    azim = np.linspace(np.pi, -np.pi, w, dtype=np.float64)[None, :]
    xyz = np.zeros((batch, 3, h, w_pad), np.float32)
    tgt = np.zeros((batch, 8, h, w_pad), np.float32)
    wgt = np.zeros((batch, 8, h, w_pad), np.float32)
    nrm = np.zeros((batch, 8, h, w_pad), np.float32)
    msk = np.zeros((batch, 1, h, w_pad), np.float32)
    gt = np.zeros((batch, n_gt, 8), np.float32)
    gt[:, :, 3:7] = np.float32(1e-3)
    for b in range(batch):
        r = rng.uniform(2.0, 75.0, size=(h, w))
        x = r * np.cos(incl) * np.cos(azim)
        y = r * np.cos(incl) * np.sin(azim)
        z = r * np.sin(incl)
        valid = rng.uniform(size=(h, w)) >= missing
        b7 = boxes7(n_vehicles, seed * 1000 + b)
        rc = rng.uniform(8.0, 60.0, n_vehicles)
        ac = rng.uniform(-np.pi, np.pi, n_vehicles)
        b7[:, 0], b7[:, 1] = rc * np.cos(ac), rc * np.sin(ac)
        gt[b, :n_vehicles] = boxes7_to_corners10(b7)[:, :8]
        for v in range(n_vehicles):
            half = max(2, int(np.arctan2(2.0, rc[v]) / (2 * np.pi) * w))
            c0 = int((np.pi - ac[v]) / (2 * np.pi) * (w - 1))
            cols = np.arange(max(c0 - half, 0), min(c0 + half + 1, w))
            nrow = min(6, h)
            r0 = int(rng.integers(min(8, h - nrow), max(h - 8 - nrow, min(8, h - nrow)) + 1))
            rows = np.arange(r0, r0 + nrow)
            hh, ww = np.meshgrid(rows, cols, indexing="ij")
            # points on the vehicle: within ~1 m of its centre
            x[hh, ww] = b7[v, 0] + rng.uniform(-1, 1, hh.shape)
            y[hh, ww] = b7[v, 1] + rng.uniform(-1, 1, hh.shape)
            z[hh, ww] = b7[v, 2] + rng.uniform(-0.5, 0.5, hh.shape)
            valid[hh, ww] = True
            tgt[b][:, hh, ww] = np.moveaxis(encode_box8(b7[v], x[hh, ww], y[hh, ww]), -1, 0)
            wgt[b][:, hh, ww] = 1.0
            nrm[b][:, hh, ww] = 1.0 / hh.size
        p = np.stack([x, y, z], 0) * valid[None]
        xyz[b, :, :, :w] = p.astype(np.float32)
        msk[b, 0, :, :w] = valid
    out = {"gt_bbox_veh_for_iou_pred": gt}
    for s in strides:
        out["rpn_reg_target_s%d" % s] = np.ascontiguousarray(tgt[..., ::s])
        out["rpn_reg_weight_s%d" % s] = np.ascontiguousarray(wgt[..., ::s])
        out["reg_normalize_weight_s%d" % s] = np.ascontiguousarray(nrm[..., ::s])
        out["range_image_mask_s%d" % s] = np.ascontiguousarray(msk[..., ::s])
        out["pc_vehicle_frame_s%d" % s] = np.ascontiguousarray(xyz[..., ::s].reshape(batch, 3, -1).transpose(0, 2, 1))
    return out


def boxes7_to_corners24(b7):
    """[cx,cy,cz,l,w,h,yaw] -> (M,24): 8 corners xyz, bottom face A,B,C,D then top face E,F,G,H (the layout
    operator_cxx/src_cxx/assigner.h:31-35 reads: A,B,C,D footprint, A.z bottom, E.z top)."""
    c10 = boxes7_to_corners10(b7)
    m = c10.shape[0]
    out = np.empty((m, 8, 3), np.float32)
    for k in range(4):
        out[:, k, 0] = out[:, k + 4, 0] = c10[:, 2 * k]
        out[:, k, 1] = out[:, k + 4, 1] = c10[:, 2 * k + 1]
        out[:, k, 2] = c10[:, 8]
        out[:, k + 4, 2] = c10[:, 9]
    return out.reshape(m, 24)


def assign_frame(n_vehicles=30, seed=0, h=H_RANGE, w=W_RANGE, missing=0.10):
    """One synthetic frame for the target-assignment path: points (h*w,3) in the vehicle frame of which a few
    hundred fall inside each of n_vehicles boxes (plus near misses around them), the validity mask, the boxes as
    7-dof and as 8 corners."""
    rng = np.random.default_rng(seed)
    incl = np.linspace(-0.31, 0.04, h, dtype=np.float64)[:, None]
    azim = np.linspace(np.pi, -np.pi, w, dtype=np.float64)[None, :]
    r = rng.uniform(2.0, 75.0, size=(h, w))
    x = r * np.cos(incl) * np.cos(azim)
    y = r * np.cos(incl) * np.sin(azim)
    z = r * np.sin(incl)
    b7 = boxes7(n_vehicles, seed + 31)
    rc, ac = rng.uniform(8.0, 60.0, n_vehicles), rng.uniform(-np.pi, np.pi, n_vehicles)
    b7[:, 0], b7[:, 1], b7[:, 2] = rc * np.cos(ac), rc * np.sin(ac), rng.uniform(0.5, 1.5, n_vehicles)
    for v in range(n_vehicles):
        half = max(2, int(np.arctan2(3.0, rc[v]) / (2 * np.pi) * w))
        c0 = int((np.pi - ac[v]) / (2 * np.pi) * (w - 1))
        cols = np.arange(max(c0 - half, 0), min(c0 + half + 1, w))
        r0 = int(rng.integers(0, max(h - 8, 1)))
        hh, ww = np.meshgrid(np.arange(r0, min(r0 + 8, h)), cols, indexing="ij")
        # scattered around the box: inside, on the faces' outside, above / below
        x[hh, ww] = b7[v, 0] + rng.uniform(-3.5, 3.5, hh.shape)
        y[hh, ww] = b7[v, 1] + rng.uniform(-3.5, 3.5, hh.shape)
        z[hh, ww] = b7[v, 2] + rng.uniform(-1.5, 1.5, hh.shape)
    mask = (rng.uniform(size=(h, w)) >= missing).astype(np.float32)
    pc = np.stack([x, y, z], -1).astype(np.float32) * mask[..., None]
    return pc.reshape(-1, 3), mask.reshape(-1), b7.astype(np.float32), boxes7_to_corners24(b7)


def shipped_config(is_train=True, hw=(H_RANGE, W_PADDED), fp16=True):
    """The VALUES of config/rangedet/rangedet_veh_wo_aug_4_18e.py:31-141,177-184 as parameter classes, for a box
    without the reference checkout (bench.py on the GPU box); with a checkout, rangedet_b200.shim.load_config()
    reads the file itself.  -> (BackboneParam, RpnParam, OptimizeParam.optimizer)"""
    class General:
        batch_image = 2 if is_train else 1           # :32
        scale_loss_shift = 128                       # :36
        class_names = ('veh',)                       # :41
        num_classes = 1

    General.fp16 = fp16                              # :35 (True as shipped)

    class BackboneParam:                             # :89-108
        fp16 = General.fp16
        normalizer = None                            # the file passes normalizer_factory(type="localbn") (:56)
        fpn_strides = (1, 2, 4)
        batch_image = General.batch_image
        range_image_shape_hw = hw
        meta_kernel_units = {'res1_unit2': dict(stride=1, meta_func_param='meta_baseline_bias', data_channels=64,
                                                coord_channels=3, channel_list=[32, 64], kernel_size=3)}
        num_block = {'res1': 2, 'res2a': 3, 'res2': 3, 'res3a': 5, 'res3': 5, 'agg1': 2, 'agg2': 2, 'agg2a': 1, 'agg3': 2}
        num_filter = {'res1': 64, 'res2a': 64, 'res2': 128, 'res3a': 128, 'res3': 128, 'agg1': 64, 'agg2': 128,
                      'agg2a': 64, 'agg3': 64}
        add_data_sc = True

    class RpnParam:                                  # :110-141
        fp16 = General.fp16
        normalizer = None
        batch_image = General.batch_image
        scale_loss_shift = General.scale_loss_shift
        class_names = General.class_names
        num_classes = General.num_classes
        fpn_strides = (1, 2, 4)
        num_reg_delta = 8
        wnms = True

        class loss:
            alpha = 1
            gamma = 2
            reg_loss_weight = 8.0
            cls_loss_weight = 10.0
            iou_type = 'bev'
            smooth_l1_scalar = 3

        class head:
            cls_conv_layers = 4
            cls_conv_channel = 128
            reg_conv_layers = 4
            reg_conv_channel = 128

        class all_proposal:
            rpn_pre_nms_top_n = {'veh': 50000, 'ped': 5000, 'cyc': 5000}
            rpn_post_nms_top_n = {'veh': 200, 'ped': 200, 'cyc': 100}
            nms_thr = {'veh': 0.2, 'ped': 0.2, 'cyc': 0.2}

    class optimizer:                                 # :178-184
        type = "sgd"
        lr = 0.01 / 8 * 1 * General.batch_image * 5
        momentum = 0.9
        wd = 0.00001
        clip_gradient = 35

    return BackboneParam, RpnParam, optimizer
