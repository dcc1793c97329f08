"""Batched, device-resident hot path: network forward -> cubic x4 up-sample -> key-point extraction ->
PAF grouping, for a fixed (batch, height, width).  It is what demo.py's per-frame loop
(reference demo.py:91-100: infer_fast, 18 x extract_keypoints, group_keypoints) does, for a whole
batch, without the heat-maps / PAFs ever leaving the GPU.

    pipe = PosePipeline(net, batch=64, height=368, width=656, precision="bf16")
    res = pipe(frames_pinned)          # float32 [64,3,368,656] pinned host tensor -> PoseResult (host)
    poses, all_keypoints = res.frame(0)   # exactly what group_keypoints returns for that frame

All buffers (activations, up-sampled maps, key-point / pose tables, workspaces, pinned staging) are
allocated once in the constructor; a call only enqueues kernels and copies on one stream.
"""
import numpy as np
import torch

from . import _lib, postproc
from .engine import HEAD_LD


class PoseResult:
    """Host copy of one batch of results (NumPy views of pinned staging buffers; valid until the next call)."""

    def __init__(self, pose_entries, n_poses, kpts, counts, kpt_start, overflow):
        self.pose_entries, self.n_poses = pose_entries, n_poses
        self.kpts, self.counts, self.kpt_start, self.overflow = kpts, counts, kpt_start, overflow

    def check(self):
        postproc.raise_on_overflow(self.overflow)
        return self

    def keypoints_by_type(self, i):
        return postproc.keypoint_lists(self.kpts, self.counts, self.kpt_start, i)

    def frame(self, i):
        """(pose_entries, all_keypoints) of frame i with the reference's dtypes/shapes (modules/keypoints.py:201)."""
        by_type = self.keypoints_by_type(i)
        all_keypoints = np.array([item for sub in by_type for item in sub])
        return postproc.pose_entries_array(self.pose_entries, self.n_poses, i), all_keypoints

    def total_poses(self):
        return int(self.n_poses.sum())


class PosePipeline:
    def __init__(self, net, batch, height, width, precision="bf16", upsample_ratio=4, demo=True,
                 min_paf_score=0.05, cap_kpts=128, cap_candidates=2048, cap_poses=256, cap_connections=2048,
                 heads_hook=None, fused=True):
        _lib.require_cuda()
        self.net, self.precision = net, precision
        self.n, self.H, self.W = batch, height, width
        self.ratio, self.demo, self.min_paf_score = upsample_ratio, demo, min_paf_score
        self.caps = (cap_kpts, cap_candidates, cap_poses, cap_connections)
        self.heads_hook = heads_hook
        # fused: peaks / PAF samples are computed straight from the stride-8 heads (no up-sampled maps in HBM);
        # needs an up-sampling factor >= 3, otherwise the maps are materialised like the reference does
        self.fused = bool(fused) and upsample_ratio >= 3
        eng = net.engine()
        dev = eng.device
        self.device = dev
        self.plan = eng.plan(precision, batch, height, width)
        h, w = height // 8, width // 8
        self.h, self.w = h, w
        self.Hu, self.Wu = h * upsample_ratio, w * upsample_ratio
        L = _lib.load()
        self.L = L
        with torch.cuda.device(dev):
            self.x_dev = torch.empty((batch, 3, height, width), dtype=torch.float32, device=dev)
            if not self.fused:
                self.heat_up = torch.empty((batch, self.Hu, self.Wu, 19), dtype=torch.float32, device=dev)
                self.paf_up = torch.empty((batch, self.Hu, self.Wu, 38), dtype=torch.float32, device=dev)
            self.kb = postproc.KeypointBatch(batch, postproc.NUM_KPT_TYPES, cap_kpts, dev)
            self.pose_entries = torch.empty((batch, cap_poses, postproc.POSE_ENTRY), dtype=torch.float64, device=dev)
            self.n_poses = torch.empty((batch,), dtype=torch.int32, device=dev)
            self.ws_extract = torch.empty((L.lwp_extract_workspace_bytes(batch, 18, cap_candidates),),
                                          dtype=torch.uint8, device=dev)
            self.ws_group = torch.empty((L.lwp_group_workspace_bytes(batch, cap_kpts, cap_connections, cap_poses),),
                                        dtype=torch.uint8, device=dev)
            self.stream = torch.cuda.Stream(device=dev)
        # pinned staging for results
        pin = dict(pin_memory=True)
        self.h_pose_entries = torch.empty((batch, cap_poses, postproc.POSE_ENTRY), dtype=torch.float64, **pin)
        self.h_n_poses = torch.empty((batch,), dtype=torch.int32, **pin)
        self.h_kpts = torch.empty((batch, 18, cap_kpts, 4), dtype=torch.int32, **pin)
        self.h_counts = torch.empty((batch, 18), dtype=torch.int32, **pin)
        self.h_kpt_start = torch.empty((batch, 19), dtype=torch.int32, **pin)
        self.h_overflow = torch.empty((batch,), dtype=torch.int32, **pin)
        self.d2h_bytes = sum(t.numel() * t.element_size() for t in (self.h_pose_entries, self.h_n_poses, self.h_kpts,
                                                                    self.h_counts, self.h_kpt_start, self.h_overflow))
        self.h2d_bytes = self.x_dev.numel() * 4

    @property
    def heads(self):
        """float32 [n, h, w, 64]: last stage's 19 heat-map + 38 PAF channels (+7 zero) at stride 8."""
        return self.plan.heads_f32[-1].view(self.n, self.h, self.w, HEAD_LD)

    # number of kernels of this library one step launches (memsets / copies not counted)
    @property
    def launches_per_step(self):
        return self.plan.num_compute_ops + (0 if self.fused else 2) + 3 + 3

    def enqueue(self, x_dev):
        """Enqueue one pass of the hot path on the current stream; x_dev: float32 cuda [n,3,H,W]."""
        self.plan.run_compute(x_dev)
        heads = self.heads
        if self.heads_hook is not None:
            self.heads_hook(heads)
        r = self.ratio
        ck, cc, cp, cn = self.caps
        if self.fused:
            postproc.extract_keypoints_fused(heads, r, cap_kpts=ck, cap_candidates=cc, workspace=self.ws_extract,
                                             out=self.kb)
            postproc.group_keypoints_fused(self.kb, heads, r, demo=self.demo, min_paf_score=self.min_paf_score,
                                           cap_poses=cp, cap_connections=cn, workspace=self.ws_group,
                                           out=(self.pose_entries, self.n_poses))
            return
        postproc.upsample_cubic(heads, channels=19, fx=r, fy=r, out=self.heat_up, channel_offset=0)
        postproc.upsample_cubic(heads, channels=38, fx=r, fy=r, out=self.paf_up, channel_offset=19)
        postproc.extract_keypoints_batched(self.heat_up, cap_kpts=ck, cap_candidates=cc, workspace=self.ws_extract,
                                           out=self.kb)
        postproc.group_keypoints_batched(self.kb, self.paf_up, demo=self.demo, min_paf_score=self.min_paf_score,
                                         cap_poses=cp, cap_connections=cn, workspace=self.ws_group,
                                         out=(self.pose_entries, self.n_poses))

    def run_device(self, x_dev):
        """Hot path on a device-resident batch (no host traffic); results stay in self.pose_entries etc."""
        with torch.cuda.device(self.device):
            self.enqueue(x_dev)

    def _enqueue_d2h(self):
        self.h_pose_entries.copy_(self.pose_entries, non_blocking=True)
        self.h_n_poses.copy_(self.n_poses, non_blocking=True)
        self.h_kpts.copy_(self.kb.kpts, non_blocking=True)
        self.h_counts.copy_(self.kb.counts, non_blocking=True)
        self.h_kpt_start.copy_(self.kb.kpt_start, non_blocking=True)
        self.h_overflow.copy_(self.kb.overflow, non_blocking=True)

    def host_result(self):
        return PoseResult(self.h_pose_entries.numpy(), self.h_n_poses.numpy(), self.h_kpts.numpy(),
                          self.h_counts.numpy(), self.h_kpt_start.numpy(), self.h_overflow.numpy())

    def __call__(self, frames):
        """End to end: frames float32 [n,3,H,W] on the host (pinned for an asynchronous copy) or on the
        device -> PoseResult on the host.  Synchronises before returning."""
        with torch.cuda.device(self.device), torch.cuda.stream(self.stream):
            if frames.is_cuda:
                x = frames
            else:
                self.x_dev.copy_(frames, non_blocking=True)
                x = self.x_dev
            self.enqueue(x)
            self._enqueue_d2h()
        self.stream.synchronize()
        return self.host_result()

    def error_flag(self):
        return self.plan.error_flag()
