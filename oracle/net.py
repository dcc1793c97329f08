"""oracle/net.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Plain PyTorch fp32 CPU restatement of the network forward, driven only by a state_dict with the
reference's keys (models/with_mobilenet.py:89-123, block builders modules/conv.py:4-32).  Pinned by
tests/test_oracle_golden.py against outputs of the real reference (tests/golden/net_golden.npz)."""
import torch
import torch.nn.functional as F


def _bn(sd, prefix, x):
    return F.batch_norm(x, sd[prefix + ".running_mean"], sd[prefix + ".running_var"], sd[prefix + ".weight"],
                        sd[prefix + ".bias"], False, 0.0, 1e-5)


def _conv(sd, prefix, x, stride=1, dilation=1, bn=True, relu=True, groups=1):
    """modules/conv.py:4-10: Conv2d (index 0) [+ BN (1)] [+ ReLU]; padding = dilation for 3x3, 0 for 1x1."""
    w = sd[prefix + ".0.weight"]
    pad = dilation if w.shape[-1] == 3 else 0
    x = F.conv2d(x, w, sd.get(prefix + ".0.bias"), stride, pad, dilation, groups)
    if bn:
        x = _bn(sd, prefix + ".1", x)
    return F.relu(x) if relu else x


def _conv_dw(sd, prefix, x, stride, dilation):
    """modules/conv.py:13-22"""
    w = sd[prefix + ".0.weight"]
    x = F.relu(_bn(sd, prefix + ".1", F.conv2d(x, w, None, stride, dilation, dilation, w.shape[0])))
    return F.relu(_bn(sd, prefix + ".4", F.conv2d(x, sd[prefix + ".3.weight"])))


def _conv_dw_no_bn(sd, prefix, x):
    """modules/conv.py:25-32"""
    w = sd[prefix + ".0.weight"]
    x = F.elu(F.conv2d(x, w, None, 1, 1, 1, w.shape[0]))
    return F.elu(F.conv2d(x, sd[prefix + ".2.weight"]))


_BACKBONE = [(1, 1), (2, 1), (1, 1), (2, 1), (1, 1), (1, 1), (1, 2), (1, 1), (1, 1), (1, 1), (1, 1)]


def _heads(sd, prefix, x):
    hm = _conv(sd, prefix + ".heatmaps.1", _conv(sd, prefix + ".heatmaps.0", x, bn=False), bn=False, relu=False)
    paf = _conv(sd, prefix + ".pafs.1", _conv(sd, prefix + ".pafs.0", x, bn=False), bn=False, relu=False)
    return hm, paf


def forward(state_dict, x):
    """x: float32 CPU [N,3,H,W] -> [hm_0, paf_0, ..., hm_R, paf_R] like the reference's forward (:114-123)."""
    sd = {k: v.detach().float().cpu() for k, v in state_dict.items()}
    with torch.no_grad():
        y = _conv(sd, "model.0", x.float().cpu(), stride=2)
        for i, (s, d) in enumerate(_BACKBONE):
            y = _conv_dw(sd, "model.%d" % (i + 1), y, s, d)
        a = _conv(sd, "cpm.align", y, bn=False)                                    # :19
        t = a
        for i in range(3):
            t = _conv_dw_no_bn(sd, "cpm.trunk.%d" % i, t)
        feat = _conv(sd, "cpm.conv", a + t, bn=False)                              # :20
        t = feat
        for i in range(3):
            t = _conv(sd, "initial_stage.trunk.%d" % i, t, bn=False)
        outs = list(_heads(sd, "initial_stage", t))
        s = 0
        while "refinement_stages.%d.trunk.0.initial.0.weight" % s in sd:
            t = torch.cat([feat, outs[-2], outs[-1]], 1)                           # :121
            for k in range(5):
                p = "refinement_stages.%d.trunk.%d" % (s, k)
                ini = _conv(sd, p + ".initial", t, bn=False)
                tr = _conv(sd, p + ".trunk.1", _conv(sd, p + ".trunk.0", ini), dilation=2)
                t = ini + tr                                                       # :60
            outs += list(_heads(sd, "refinement_stages.%d" % s, t))
            s += 1
    return outs
