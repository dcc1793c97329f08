"""Batched, device-resident hot path: network forward -> cubic x4 up-sample -> key-point extraction ->
PAF grouping, for a fixed (batch, height, width).  It is what demo.py's per-frame loop
(reference demo.py:91-100: infer_fast, 18 x extract_keypoints, group_keypoints) does, for a whole
batch, without the heat-maps / PAFs ever leaving the GPU.

    pipe = PosePipeline(net, batch=64, height=368, width=656, precision="bf16")
    res = pipe(frames_pinned)          # float32 [64,3,368,656] pinned host tensor -> PoseResult (host)
    poses, all_keypoints = res.frame(0)   # exactly what group_keypoints returns for that frame

Streaming use: `pipe.submit(frames)` / `pipe.collect()` keep `depth` batches in flight, so the pinned
host->device copy of batch i+1 (copy stream) overlaps the kernels of batch i (compute stream).  All
buffers (activations, key-point / pose tables, workspaces, pinned staging, input double buffer) are
allocated once in the constructor; a call only enqueues kernels and copies.
"""
import numpy as np
import torch

from . import _lib, postproc
from .engine import HEAD_LD


class PoseResult:
    """Host copy of one batch of results (NumPy views of pinned staging buffers; valid until the next call)."""

    def __init__(self, pose_entries, n_poses, kpts, counts, kpt_start, overflow, pose_kpts=None, bbox=None, confidence=None):
        self.pose_entries, self.n_poses = pose_entries, n_poses
        self.kpts, self.counts, self.kpt_start, self.overflow = kpts, counts, kpt_start, overflow
        # device-side result post-conversion (PosePipeline(convert=...)): per pose the [18, 2] int32 key-points in
        # original-frame coordinates, the bounding box and the confidence
        self.pose_kpts, self.bbox, self.confidence = pose_kpts, bbox, confidence

    def poses(self, i):
        """The list of Pose objects demo.py:104-115 builds for frame i (key-points un-scaled to the original frame)."""
        if self.pose_kpts is None:
            raise RuntimeError("build the PosePipeline with convert=dict(pad=..., scale=...) to get Pose objects")
        from .modules.pose import Pose
        return [Pose(self.pose_kpts[i, j].copy(), self.confidence[i, j], bbox=self.bbox[i, j])
                for j in range(int(self.n_poses[i]))]

    retried = None   # {frame index: (pose_entries [1,cap',20], n_poses [1], kpts, counts, kpt_start)} of frames that
                     # overflowed the pipeline's fixed-capacity tables and were re-processed with larger ones by collect()

    def check(self):
        """Raises CapacityOverflow if some frame overflowed even the largest tables collect() retries with."""
        bad = np.asarray(self.overflow).copy()
        for i in (self.retried or {}):
            bad[i] = 0
        postproc.raise_on_overflow(bad)
        return self

    def _tables(self, i):
        if self.retried and i in self.retried:
            pe, npz, kp, cnt, st = self.retried[i]
            return pe, npz, kp, cnt, st, 0
        return self.pose_entries, self.n_poses, self.kpts, self.counts, self.kpt_start, i

    def keypoints_by_type(self, i):
        _, _, kp, cnt, st, j = self._tables(i)
        return postproc.keypoint_lists(kp, cnt, st, j)

    def frame(self, i):
        """(pose_entries, all_keypoints) of frame i with the reference's dtypes/shapes (modules/keypoints.py:201)."""
        by_type = self.keypoints_by_type(i)
        all_keypoints = np.array([item for sub in by_type for item in sub])
        pe, npz, _, _, _, j = self._tables(i)
        return postproc.pose_entries_array(pe, npz, j), all_keypoints

    def total_poses(self):
        extra = sum(int(t[1][0]) - int(self.n_poses[i]) for i, t in (self.retried or {}).items())
        return int(self.n_poses.sum()) + extra


class _Chunk:
    """Plan + device result tables of one slice [lo, lo + n) of the batch."""

    def __init__(self, pipe, slot, lo, n):
        self.pipe, self.lo, self.n = pipe, lo, n
        dev = pipe.device
        eng = pipe.net.engine()
        # private buffers; with the post-processing on its own stream the plan writes the last stage's heads alternately
        # into two buffers, so post-processing (batch i) and network (batch i + 1) never touch the same one -- no copy
        self.double = bool(pipe.overlap_postproc and not pipe.graph)
        self.plan = eng.new_plan(pipe.precision, n, pipe.H, pipe.W, input_u8=pipe.input_u8, double_heads=self.double)
        self._alt, self._last_alt = 0, 0
        ck, cc, cp, cn = pipe.caps
        L = pipe.L
        if not pipe.fused:
            self.heat_up = torch.empty((n, pipe.Hu, pipe.Wu, 19), dtype=torch.float32, device=dev)
            self.paf_up = torch.empty((n, pipe.Hu, pipe.Wu, 38), dtype=torch.float32, device=dev)
        self.kb = postproc.KeypointBatch(n, postproc.NUM_KPT_TYPES, ck, dev)
        self.pose_entries = torch.empty((n, cp, postproc.POSE_ENTRY), dtype=torch.float64, device=dev)
        self.n_poses = torch.empty((n,), dtype=torch.int32, device=dev)
        self.ws_extract = torch.empty((L.lwp_extract_workspace_bytes(n, 18, cc),), dtype=torch.uint8, device=dev)
        # + the optional re-packed PAF planes (fused path: a limb's two channels staged by one contiguous copy)
        pack = L.lwp_paf_pack_bytes(n, pipe.H // 8, pipe.W // 8) if pipe.fused else 0
        self.ws_group = torch.empty((L.lwp_group_workspace_bytes(n, ck, cn, cp) + pack,), dtype=torch.uint8, device=dev)
        if pipe.convert is not None:
            self.pose_kpts = torch.empty((n, cp, postproc.NUM_KPT_TYPES, 2), dtype=torch.int32, device=dev)
            self.bbox = torch.empty((n, cp, 4), dtype=torch.int32, device=dev)
            self.confidence = torch.empty((n, cp), dtype=torch.float64, device=dev)
        self.net_done = torch.cuda.Event()
        self.pp_done_buf = [torch.cuda.Event(), torch.cuda.Event()]   # post-processing that read heads buffer 0 / 1 has finished
        self.pp_done = self.pp_done_buf[0]                            # ... of the most recent batch

    def _heads_buf(self, alt):
        t = self.plan.heads_alt if alt else self.plan.heads_f32[-1]
        return t.view(self.n, self.pipe.h, self.pipe.w, HEAD_LD)

    @property
    def heads(self):
        """The heads of the most recent pass."""
        return self._heads_buf(self._last_alt)

    def enqueue(self, x_dev):
        """One pass of the hot path for this chunk: network on the current stream, post-processing on the
        pipeline's post-processing stream (ordered after it by events)."""
        pipe = self.pipe
        cur = torch.cuda.current_stream()
        alt = self._alt if self.double else 0
        if self.double:
            cur.wait_event(self.pp_done_buf[alt])   # the post-processing that last read this heads buffer (two batches ago)
        self.plan.run_compute(x_dev, alt)
        self._last_alt = alt
        heads = self._heads_buf(alt)
        if pipe.heads_hook is not None:
            pipe.heads_hook(heads, self.lo)
        if pipe.graph:   # one stream, no events: the whole pass is (or is being captured into) one CUDA graph
            self.enqueue_postproc(heads)
            return
        if not pipe.overlap_postproc:
            self.enqueue_postproc(heads)
            self.pp_done.record(cur)
            return
        self.net_done.record(cur)
        with torch.cuda.stream(pipe.pp_stream):
            pipe.pp_stream.wait_event(self.net_done)
            self.enqueue_postproc(heads)
            self.pp_done = self.pp_done_buf[alt]
            self.pp_done.record(pipe.pp_stream)
        self._alt ^= 1

    def enqueue_postproc(self, heads, stage=None):
        pipe = self.pipe
        r = pipe.ratio
        ck, cc, cp, cn = pipe.caps
        if pipe.fused:
            if stage in (None, "extract_fused"):
                postproc.extract_keypoints_fused(heads, r, cap_kpts=ck, cap_candidates=cc, workspace=self.ws_extract,
                                                 out=self.kb)
            if stage in (None, "group_fused"):
                postproc.group_keypoints_fused(self.kb, heads, r, demo=pipe.demo, min_paf_score=pipe.min_paf_score,
                                               cap_poses=cp, cap_connections=cn, workspace=self.ws_group,
                                               out=(self.pose_entries, self.n_poses))
                self.enqueue_convert()
            return
        if stage in (None, "upsample"):
            postproc.upsample_cubic(heads, channels=19, fx=r, fy=r, out=self.heat_up, channel_offset=0)
            postproc.upsample_cubic(heads, channels=38, fx=r, fy=r, out=self.paf_up, channel_offset=19)
        if stage in (None, "extract"):
            postproc.extract_keypoints_batched(self.heat_up, cap_kpts=ck, cap_candidates=cc,
                                               workspace=self.ws_extract, out=self.kb)
        if stage in (None, "group"):
            postproc.group_keypoints_batched(self.kb, self.paf_up, demo=pipe.demo, min_paf_score=pipe.min_paf_score,
                                             cap_poses=cp, cap_connections=cn, workspace=self.ws_group,
                                             out=(self.pose_entries, self.n_poses))
            self.enqueue_convert()

    def enqueue_convert(self):
        """Result post-conversion (demo.py:101-115) on the device, when the pipeline was built with convert=..."""
        cv = self.pipe.convert
        if cv is None:
            return
        if getattr(self, "_xform", None) is None:   # once: the upload is a pageable copy that blocks the host until the
            pad, scale = cv["pad"], cv["scale"]     # stream reaches it -- per batch it serialised submit() with the network
            if np.ndim(scale) > 0:   # per-frame transforms: this chunk's slice
                pad, scale = np.asarray(pad)[self.lo:self.lo + self.n], np.asarray(scale)[self.lo:self.lo + self.n]
            self._xform = postproc.pose_convert_xform(self.n, pad, scale, self.pose_entries.device)
            torch.cuda.synchronize(self.pose_entries.device)
        postproc.pose_convert(self.pose_entries, self.n_poses, self.kb, stride=cv.get("stride", 8),
                              upsample_ratio=self.pipe.ratio, out=(self.pose_kpts, self.bbox, self.confidence),
                              xform=self._xform)


class _Slot:
    """One in-flight batch: device input buffer, pinned host result tables, events."""

    def __init__(self, pipe):
        b, (ck, _, cp, _) = pipe.n, pipe.caps
        if pipe.input_u8 is None:
            self.x_dev = torch.empty((b, 3, pipe.H, pipe.W), dtype=torch.float32, device=pipe.device)
        else:
            self.x_dev = torch.empty((b, pipe.H, pipe.W, 3), dtype=torch.uint8, device=pipe.device)
        self.raw_dev = None
        if pipe.raw is not None:
            self.raw_dev = torch.empty((b,) + pipe.raw["size"] + (3,), dtype=torch.uint8, device=pipe.device)
        pin = dict(pin_memory=True)
        self.h_pose_entries = torch.empty((b, cp, postproc.POSE_ENTRY), dtype=torch.float64, **pin)
        self.h_n_poses = torch.empty((b,), dtype=torch.int32, **pin)
        self.h_kpts = torch.empty((b, 18, ck, 4), dtype=torch.int32, **pin)
        self.h_counts = torch.empty((b, 18), dtype=torch.int32, **pin)
        self.h_kpt_start = torch.empty((b, 19), dtype=torch.int32, **pin)
        self.h_overflow = torch.empty((b,), dtype=torch.int32, **pin)
        self.h_conv = None
        if pipe.convert is not None:
            self.h_conv = (torch.empty((b, cp, 18, 2), dtype=torch.int32, **pin), torch.empty((b, cp, 4), dtype=torch.int32, **pin),
                           torch.empty((b, cp), dtype=torch.float64, **pin))
        # heads of the frames whose tables overflowed (copied only for those: lwp_copy_flagged), for the retry in collect()
        self.heads_keep = torch.empty((b, pipe.h, pipe.w, HEAD_LD), dtype=torch.float32, device=pipe.device)
        self.copied = torch.cuda.Event()   # H2D of this slot's input finished
        self.consumed = torch.cuda.Event() # the network has finished reading x_dev
        self.done = torch.cuda.Event()     # compute + D2H of this slot finished
        self.busy = False

    def tables(self):
        return (self.h_pose_entries, self.h_n_poses, self.h_kpts, self.h_counts, self.h_kpt_start, self.h_overflow) + (
            self.h_conv if self.h_conv is not None else ())

    def result(self):
        return PoseResult(*[t.numpy() for t in self.tables()])


class PosePipeline:
    """`depth` batches can be in flight: submit() enqueues the host->device copy of a batch on a copy stream
    and its kernels + result read-back on the compute stream; collect() returns the oldest batch's
    PoseResult.  With depth >= 2 the copy of batch i+1 overlaps the kernels of batch i.  pipe(frames) is
    submit + collect."""

    def __init__(self, net, batch, height, width, precision="bf16", upsample_ratio=4, demo=True,
                 min_paf_score=0.05, cap_kpts=128, cap_candidates=2048, cap_poses=256, cap_connections=2048,
                 heads_hook=None, fused=True, chunk=None, depth=2, overlap_postproc=True, input_format="f32_nchw",
                 img_mean=(128, 128, 128), img_scale=1 / 256, graph=False, convert=None, raw_size=None):
        _lib.require_cuda()
        self.net, self.precision = net, precision
        self.n, self.H, self.W = batch, height, width
        self.ratio, self.demo, self.min_paf_score = upsample_ratio, demo, min_paf_score
        self.caps = (cap_kpts, cap_candidates, cap_poses, cap_connections)
        # optional callable(heads [n_chunk, h, w, 64] float32, first_frame_index) run between network and post-processing
        self.heads_hook = heads_hook
        # fused: peaks / PAF samples are computed straight from the stride-8 heads (no up-sampled maps in HBM);
        # needs an up-sampling factor >= 3, otherwise the maps are materialised like the reference does
        self.fused = bool(fused) and upsample_ratio >= 3
        # post-processing of batch i on a second stream, overlapping the network of batch i+1
        self.overlap_postproc = bool(overlap_postproc)
        # graph=True: the ~60 launches of one pass (network, post-processing, result read-back) are captured once per
        # input buffer into a CUDA graph and replayed -- for small batches (the reference's batch-1 demo loop) the
        # launches, not the kernels, are the cost.  Everything then runs on one stream.
        # convert=dict(pad=[top, left, bottom, right], scale=s[, stride=8]) (or per-frame sequences of both): the result
        # post-conversion of demo.py:101-115 + Pose.get_bbox runs on the device too and PoseResult.poses(i) is available
        self.convert = convert
        self.graph = bool(graph)
        if self.graph:
            self.overlap_postproc = False
        self._graphs = {}
        # "f32_nchw": normalised float32 [n,3,H,W] like the reference's tensor_img; "u8_nhwc": raw BGR frames uint8
        # [n,H,W,3] already at the network size -- (img - mean) * scale (val.normalize) is fused into the stem kernel
        # "u8_raw": camera frames uint8 [n, h, w, 3] at their own size raw_size=(h, w): the cubic resize to the network height
        # and the centred pad of infer_fast (demo.py:57-62) run on the GPU too (lwp_resize_pad_u8), then as "u8_nhwc"
        if input_format not in ("f32_nchw", "u8_nhwc", "u8_raw"):
            raise ValueError("input_format must be 'f32_nchw', 'u8_nhwc' or 'u8_raw'")
        self.input_u8 = (tuple(img_mean), float(img_scale)) if input_format in ("u8_nhwc", "u8_raw") else None
        self.raw = None
        if input_format == "u8_raw":
            if raw_size is None:
                raise ValueError("input_format='u8_raw' needs raw_size=(frame height, frame width)")
            scale, (rh, rw), padded, pad = postproc.infer_fast_geometry(raw_size[0], raw_size[1], height)
            if tuple(padded) != (height, width):
                raise ValueError("a %dx%d frame resized to height %d pads to %dx%d, not to the pipeline's %dx%d"
                                 % (raw_size[0], raw_size[1], height, padded[0], padded[1], height, width))
            # pad with the mean: the stem normalises it to exactly 0, the reference's pad value in the normalised image
            self.raw = dict(size=(int(raw_size[0]), int(raw_size[1])), scale=scale, resized=(rh, rw), pad=pad,
                            pad_value=tuple(int(round(float(m))) for m in img_mean))

        self._engine = net.engine()      # the packed weights this pipeline's plans were recorded with
        self.device = self._engine.device
        self.h, self.w = height // 8, width // 8
        self.Hu, self.Wu = self.h * upsample_ratio, self.w * upsample_ratio
        self.L = _lib.load()
        chunk = batch if chunk is None else max(1, min(int(chunk), batch))
        with torch.cuda.device(self.device):
            self.chunks = []
            lo = 0
            while lo < batch:
                n = min(chunk, batch - lo)
                self.chunks.append(_Chunk(self, len(self.chunks), lo, n))
                lo += n
            self.slots = [_Slot(self) for _ in range(max(1, int(depth)))]
            self.copy_stream = torch.cuda.Stream(device=self.device)
            self.stream = torch.cuda.Stream(device=self.device)
            self.pp_stream = torch.cuda.Stream(device=self.device)
        self._next, self._pending = 0, []
        self.d2h_bytes = sum(t.numel() * t.element_size() for t in self.slots[0].tables())
        self.h2d_bytes = batch * 3 * height * width * (4 if self.input_u8 is None else 1)
        if self.raw is not None:
            self.h2d_bytes = batch * 3 * self.raw["size"][0] * self.raw["size"][1]

    @property
    def heads(self):
        """float32 [n, h, w, 64]: last stage's 19 heat-map + 38 PAF channels (+7 zero) at stride 8 (a copy)."""
        return torch.cat([c.heads for c in self.chunks], 0)

    def join(self):
        """Make the current stream wait for every enqueued post-processing (call before timing / reading results
        produced by run_device)."""
        cur = torch.cuda.current_stream()
        if self.graph:
            cur.wait_stream(self.stream)
            return
        for c in self.chunks:
            cur.wait_event(c.pp_done)

    def _graphed(self, key, fn):
        """Run fn() on self.stream (the current stream): eagerly the first time (one-time kernel attribute set-up
        must not happen under capture), captured into a CUDA graph the second time, replayed from then on."""
        state = self._graphs.get(key)
        if state is None:
            fn()
            self._graphs[key] = False
            return
        if state is False:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=self.stream, capture_error_mode="thread_local"):
                fn()
            self._graphs[key] = state = g
        state.replay()

    @property
    def n_poses(self):
        return torch.cat([c.n_poses for c in self.chunks], 0)

    @property
    def pose_entries(self):
        return torch.cat([c.pose_entries for c in self.chunks], 0)

    # number of kernels of this library one step launches (memsets / copies not counted)
    @property
    def launches_per_step(self):
        return sum(c.plan.num_compute_ops + (0 if self.fused else 2) + 3 + 3 for c in self.chunks)

    def _check_engine(self):
        """net.cuda() / load_state_dict() / refresh() drop the engine whose packed weights this pipeline's plans hold:
        running on would silently use the old weights."""
        if self.net._engine is not self._engine:
            raise RuntimeError("the network's parameters changed after this PosePipeline was built (load_state / .cuda() / "
                               "refresh()): build a new PosePipeline")

    def run_device(self, x_dev):
        """Hot path on a device-resident batch (no host traffic): network on the current stream, post-processing
        on the post-processing stream (successive calls overlap); results stay on the device
        (self.pose_entries / self.n_poses / chunk.kb).  Call join() (or synchronise the device) before reading."""
        self._check_engine()
        with torch.cuda.device(self.device):
            if self.graph:
                cur = torch.cuda.current_stream()
                self.stream.wait_stream(cur)
                with torch.cuda.stream(self.stream):
                    self._graphed(("dev", x_dev.data_ptr()),
                                  lambda: [c.enqueue(x_dev[c.lo:c.lo + c.n]) for c in self.chunks])
                return
            for c in self.chunks:
                c.enqueue(x_dev[c.lo:c.lo + c.n])

    def submit(self, frames):
        """Enqueue one batch: frames float32 [n,3,H,W], pinned host memory (asynchronous copy) or device."""
        self._check_engine()
        if len(self._pending) == len(self.slots):
            raise RuntimeError("pipeline full: collect() a result first (depth=%d)" % len(self.slots))
        slot = self.slots[self._next]
        self._next = (self._next + 1) % len(self.slots)
        with torch.cuda.device(self.device):
            if self.raw is not None:
                # camera frames: H2D of the raw frames (or none, if they are on the device), then resize + pad on the GPU
                with torch.cuda.stream(self.copy_stream):
                    self.copy_stream.wait_event(slot.consumed)
                    if frames.is_cuda:
                        self.copy_stream.wait_stream(torch.cuda.current_stream())
                    slot.raw_dev.copy_(frames, non_blocking=True)
                    r = self.raw
                    postproc.resize_pad_u8(slot.raw_dev, fx=r["scale"], fy=r["scale"], padded=(self.H, self.W), top=r["pad"][0],
                                           left=r["pad"][1], pad_value=r["pad_value"], out=slot.x_dev)
                    slot.copied.record(self.copy_stream)
                x = slot.x_dev
            elif frames.is_cuda:
                x = frames
                slot.copied.record(torch.cuda.current_stream())
            else:
                with torch.cuda.stream(self.copy_stream):
                    self.copy_stream.wait_event(slot.consumed)   # the kernels that read this buffer last are finished
                    slot.x_dev.copy_(frames, non_blocking=True)
                    slot.copied.record(self.copy_stream)
                x = slot.x_dev
            def read_back(stream_waits):
                for c in self.chunks:
                    sl = slice(c.lo, c.lo + c.n)
                    if stream_waits is not None:
                        stream_waits.wait_event(c.pp_done)
                    slot.h_pose_entries[sl].copy_(c.pose_entries, non_blocking=True)
                    slot.h_n_poses[sl].copy_(c.n_poses, non_blocking=True)
                    slot.h_kpts[sl].copy_(c.kb.kpts, non_blocking=True)
                    slot.h_counts[sl].copy_(c.kb.counts, non_blocking=True)
                    slot.h_kpt_start[sl].copy_(c.kb.kpt_start, non_blocking=True)
                    slot.h_overflow[sl].copy_(c.kb.overflow, non_blocking=True)
                    src = c.heads
                    _lib.check(self.L.lwp_copy_flagged(src.data_ptr(), slot.heads_keep[sl].data_ptr(), c.kb.overflow.data_ptr(),
                                                       c.n, self.h * self.w * HEAD_LD * 4, _lib.current_stream()),
                               "lwp_copy_flagged")
                    if slot.h_conv is not None:
                        slot.h_conv[0][sl].copy_(c.pose_kpts, non_blocking=True)
                        slot.h_conv[1][sl].copy_(c.bbox, non_blocking=True)
                        slot.h_conv[2][sl].copy_(c.confidence, non_blocking=True)

            if self.graph:
                def whole_pass():
                    for c in self.chunks:
                        c.enqueue(x[c.lo:c.lo + c.n])
                    read_back(None)
                with torch.cuda.stream(self.stream):
                    self.stream.wait_event(slot.copied)
                    self._graphed(("slot", self.slots.index(slot), x.data_ptr()), whole_pass)
                    slot.consumed.record(self.stream)
                    slot.done.record(self.stream)
                self._pending.append(slot)
                return
            with torch.cuda.stream(self.stream):
                self.stream.wait_event(slot.copied)
                for c in self.chunks:
                    c.enqueue(x[c.lo:c.lo + c.n])
                slot.consumed.record(self.stream)       # the network has read this slot's input buffer
            tail = self.pp_stream if self.overlap_postproc else self.stream
            with torch.cuda.stream(tail):
                read_back(tail)
                slot.done.record(tail)
        self._pending.append(slot)

    def collect(self):
        """PoseResult of the oldest submitted batch (blocks until its read-back has finished)."""
        if not self._pending:
            raise RuntimeError("nothing submitted")
        slot = self._pending.pop(0)
        slot.done.synchronize()
        res = slot.result()
        bad = np.nonzero(res.overflow)[0]
        if bad.size:
            res.retried = {int(i): self._retry_frame(slot, int(i)) for i in bad}
            res.retried = {i: t for i, t in res.retried.items() if t is not None}
        return res

    def _retry_frame(self, slot, i):
        """One frame exceeded cap_kpts / cap_candidates / cap_connections / cap_poses: re-run its post-processing alone on
        the kept heads with tables 4x, 16x ... larger (the 63 other frames of the batch keep their results).  Returns host
        tables, or None if even the largest tables overflow (check() then raises)."""
        ck, cc, cp, cn = self.caps
        heads = slot.heads_keep[i:i + 1]
        with torch.cuda.device(self.device):
            for _ in range(3):
                ck, cc, cp, cn = ck * 4, cc * 4, cp * 4, min(cn * 4, 1 << 15)
                if self.fused:
                    kb = postproc.extract_keypoints_fused(heads, self.ratio, cap_kpts=ck, cap_candidates=cc)
                    pe, npz = postproc.group_keypoints_fused(kb, heads, self.ratio, demo=self.demo, min_paf_score=self.min_paf_score,
                                                             cap_poses=cp, cap_connections=cn)
                else:
                    heat = postproc.upsample_cubic(heads, channels=19, fx=self.ratio, fy=self.ratio, channel_offset=0)
                    paf = postproc.upsample_cubic(heads, channels=38, fx=self.ratio, fy=self.ratio, channel_offset=19)
                    kb = postproc.extract_keypoints_batched(heat, cap_kpts=ck, cap_candidates=cc)
                    pe, npz = postproc.group_keypoints_batched(kb, paf, demo=self.demo, min_paf_score=self.min_paf_score,
                                                               cap_poses=cp, cap_connections=cn)
                kp, cnt, st, ovf = kb.to_host()
                if not ovf[0]:
                    return pe.cpu().numpy(), npz.cpu().numpy(), kp, cnt, st
        return None

    def __call__(self, frames):
        """End to end, synchronous: frames -> PoseResult on the host."""
        while self._pending:
            self.collect()
        self.submit(frames)
        return self.collect()

    def error_flag(self):
        return max(c.plan.error_flag() for c in self.chunks)
