"""Seeded synthetic inputs for the hot path (no datasets or checkpoints are reachable offline).

* `synthetic_pose_maps`  -- stride-8 heat-maps / PAFs with a known number of stick-figure persons
                           (BASELINE.json configs[2]); channel order follows the reference tables
                           (modules/keypoints.py:5-8 of the reference).
* `synthetic_frames`     -- uint8 BGR frames / normalised net inputs (configs[0], configs[1]).
* `randomize_bn_`, `apply_head_gain_` -- make a random-init network exercise BN folding and produce
                           key-points (random init alone stays below the 0.1 peak threshold).
"""
import numpy as np

from .modules.keypoints import BODY_PARTS_KPT_IDS, BODY_PARTS_PAF_IDS

# 18 joints (reference order: nose, neck, r_sho, r_elb, r_wri, l_sho, l_elb, l_wri, r_hip, r_knee, r_ank,
# l_hip, l_knee, l_ank, r_eye, l_eye, r_ear, l_ear) as (x, y) fractions of the person height.
_SKELETON = np.array([
    [0.00, 0.06], [0.00, 0.19], [-0.19, 0.20], [-0.27, 0.38], [-0.30, 0.55], [0.19, 0.20], [0.27, 0.38],
    [0.30, 0.55], [-0.11, 0.55], [-0.12, 0.77], [-0.12, 0.97], [0.11, 0.55], [0.12, 0.77], [0.12, 0.97],
    [-0.045, 0.025], [0.045, 0.025], [-0.10, 0.05], [0.10, 0.05]], dtype=np.float64)


def person_joints(rng, h, w):
    """Joint coordinates (18, 2) as (x, y) in stride-8 pixels for one random person."""
    height = rng.uniform(0.35, 0.9) * h
    width = 0.6 * height
    x0 = rng.uniform(0.5 * width + 1, max(0.5 * width + 2, w - 0.5 * width - 1))
    y0 = rng.uniform(1, max(2, h - height - 1))
    jitter = rng.normal(0, 0.01, size=_SKELETON.shape)
    pts = (_SKELETON + jitter) * height
    pts[:, 0] += x0
    pts[:, 1] += y0
    return pts


def render_maps(persons, h, w, sigma=7.0 / 8.0):
    """heat-maps [19,h,w] and PAFs [38,h,w] float32 for a list of (18,2) joint arrays."""
    ys, xs = np.mgrid[0:h, 0:w].astype(np.float64)
    hm = np.zeros((19, h, w), np.float64)
    paf = np.zeros((38, h, w), np.float64)
    for pts in persons:
        for k in range(18):
            g = np.exp(-((xs - pts[k, 0]) ** 2 + (ys - pts[k, 1]) ** 2) / (2 * sigma * sigma))
            hm[k] = np.maximum(hm[k], g)
        for limb, (ka, kb) in enumerate(BODY_PARTS_KPT_IDS):
            ax, ay = pts[ka]
            bx, by = pts[kb]
            vx, vy = bx - ax, by - ay
            n = np.hypot(vx, vy)
            if n < 1e-6:
                continue
            ux, uy = vx / n, vy / n
            t = (xs - ax) * ux + (ys - ay) * uy           # along the limb
            d = np.abs((xs - ax) * uy - (ys - ay) * ux)   # across the limb
            m = (t >= -0.5) & (t <= n + 0.5) & (d <= 1.0)
            cx, cy = BODY_PARTS_PAF_IDS[limb]
            paf[cx][m] = ux
            paf[cy][m] = uy
    hm[18] = 1.0 - hm[:18].max(axis=0)
    return hm.astype(np.float32), paf.astype(np.float32)


def synthetic_pose_maps(batch, h=46, w=82, seed=0, noise=0.0, max_persons=30, persons=None):
    """Batch of stride-8 maps; frame b holds 1 + (b mod max_persons) persons unless `persons` fixes it.

    Returns (heatmaps [B,19,h,w] f32, pafs [B,38,h,w] f32, n_persons list)."""
    rng = np.random.default_rng(seed)
    hms = np.empty((batch, 19, h, w), np.float32)
    pafs = np.empty((batch, 38, h, w), np.float32)
    counts = []
    for b in range(batch):
        n = (1 + (b % max_persons)) if persons is None else int(persons)
        people = [person_joints(rng, h, w) for _ in range(n)]
        hm, paf = render_maps(people, h, w)
        if noise > 0:
            hm = hm + rng.normal(0, noise, hm.shape).astype(np.float32)
            paf = paf + rng.normal(0, noise, paf.shape).astype(np.float32)
        hms[b], pafs[b] = hm, paf
        counts.append(n)
    return hms, pafs, counts


def synthetic_frames(batch, height=720, width=1280, seed=0):
    """uint8 BGR frames like demo.py's image provider would yield."""
    rng = np.random.default_rng(seed)
    return rng.integers(0, 256, (batch, height, width, 3), dtype=np.uint8)


def synthetic_net_input(batch, height=368, width=656, seed=1):
    """Normalised NCHW float32 net input: (img - 128) / 256 lives in [-0.5, 0.498]."""
    import torch
    g = torch.Generator().manual_seed(seed)
    return torch.rand(batch, 3, height, width, generator=g) - 0.5


def randomize_bn_(net, seed=0):
    """Give every BatchNorm2d non-trivial running stats / affine so BN folding is really exercised."""
    import torch
    g = torch.Generator().manual_seed(seed)
    for m in net.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            with torch.no_grad():
                m.running_mean.copy_(torch.randn(m.num_features, generator=g) * 0.1)
                m.running_var.copy_(torch.rand(m.num_features, generator=g) * 0.5 + 0.75)
                m.weight.copy_(torch.rand(m.num_features, generator=g) * 0.5 + 0.75)
                m.bias.copy_(torch.randn(m.num_features, generator=g) * 0.1)
    return net


def apply_head_gain_(net, gain):
    """Scale the last 1x1 conv of every heat-map / PAF head (weights and bias) so a random-init net
    produces values above the key-point threshold."""
    import torch
    stages = [net.initial_stage] + list(net.refinement_stages)
    with torch.no_grad():
        for st in stages:
            for head in (st.heatmaps, st.pafs):
                last = head[1][0]
                last.weight.mul_(gain)
                if last.bias is not None:
                    last.bias.mul_(gain)
    return net
