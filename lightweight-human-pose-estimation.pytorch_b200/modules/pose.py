"""Drop-in mirror of the reference's modules/pose.py: the `Pose` object demo.py builds per detected person
(:21-39 key-points, confidence, bounding box), the OKS-like `get_similarity` (:65-75) and `track_poses` (:78-118:
id propagation between frames + optional One-Euro smoothing).

This is the consumer side of the hot path.  The batched part -- coordinate un-scaling,
per-pose key-point gather and bounding boxes for a whole batch -- runs on the GPU (`lwp_pose_convert`,
`lwpose_b200.postproc.pose_convert`, `PoseResult.poses`); tracking is per-stream, sequential, stateful host logic
and stays NumPy, vectorised over the 18 key-points instead of the reference's Python loops (same float32/float64
operations per element, so the same decisions)."""
import numpy as np

from .keypoints import BODY_PARTS_KPT_IDS, BODY_PARTS_PAF_IDS
from .one_euro_filter import OneEuroFilter


class Pose:
    num_kpts = 18
    kpt_names = ['nose', 'neck', 'r_sho', 'r_elb', 'r_wri', 'l_sho', 'l_elb', 'l_wri', 'r_hip', 'r_knee', 'r_ank',
                 'l_hip', 'l_knee', 'l_ank', 'r_eye', 'l_eye', 'r_ear', 'l_ear']
    sigmas = np.array([.26, .79, .79, .72, .62, .79, .72, .62, 1.07, .87, .89, 1.07, .87, .89, .25, .25, .35, .35],
                      dtype=np.float32) / 10.0
    vars = (sigmas * 2) ** 2
    last_id = -1
    color = [0, 224, 255]

    def __init__(self, keypoints, confidence, bbox=None):
        """keypoints: int32 [18, 2], (-1, -1) for a key-point that was not found; bbox: (x, y, w, h), computed when
        not supplied (PoseResult.poses hands over the one the GPU computed)."""
        self.keypoints = keypoints
        self.confidence = confidence
        self.bbox = Pose.get_bbox(keypoints) if bbox is None else tuple(int(v) for v in bbox)
        self.id = None
        self.filters = [[OneEuroFilter(), OneEuroFilter()] for _ in range(Pose.num_kpts)]

    @staticmethod
    def get_bbox(keypoints):
        """cv2.boundingRect of the found key-points: (min x, min y, max x - min x + 1, max y - min y + 1)."""
        pts = np.asarray(keypoints)
        pts = pts[pts[:, 0] != -1].astype(np.int32)
        if pts.shape[0] == 0:
            return (0, 0, 0, 0)
        x0, y0 = int(pts[:, 0].min()), int(pts[:, 1].min())
        return (x0, y0, int(pts[:, 0].max()) - x0 + 1, int(pts[:, 1].max()) - y0 + 1)

    def update_id(self, id=None):
        if id is None:
            Pose.last_id += 1
            id = Pose.last_id
        self.id = id

    def draw(self, img):
        """Skeleton overlay (reference :47-62); needs OpenCV, which only the demo GUI uses."""
        import cv2
        assert self.keypoints.shape == (Pose.num_kpts, 2)
        have = self.keypoints[:, 0] != -1
        for ka, kb in BODY_PARTS_KPT_IDS[:len(BODY_PARTS_PAF_IDS) - 2]:
            for k in (ka, kb):
                if have[k]:
                    cv2.circle(img, (int(self.keypoints[k, 0]), int(self.keypoints[k, 1])), 3, Pose.color, -1)
            if have[ka] and have[kb]:
                cv2.line(img, (int(self.keypoints[ka, 0]), int(self.keypoints[ka, 1])),
                         (int(self.keypoints[kb, 0]), int(self.keypoints[kb, 1])), Pose.color, 2)


def get_similarity(a, b, threshold=0.5):
    """Number of key-points found in both poses whose OKS-like similarity exceeds `threshold`."""
    both = (a.keypoints[:, 0] != -1) & (b.keypoints[:, 0] != -1)
    if not both.any():
        return 0
    distance = np.sum((a.keypoints - b.keypoints) ** 2, axis=1)           # integer, like the reference's per-row np.sum
    area = max(a.bbox[2] * a.bbox[3], b.bbox[2] * b.bbox[3])
    similarity = np.exp(-distance / (2 * (area + np.spacing(1)) * Pose.vars))   # float64 / (float64 * float32 -> float64)
    return int(np.count_nonzero(both & (similarity > threshold)))


def track_poses(previous_poses, current_poses, threshold=3, smooth=False):
    """Give every current pose the id of the most similar unclaimed previous pose (at least `threshold` similar
    key-points, most confident current poses first), else a new id; with smooth=True the key-points go through the
    track's One-Euro filters and the bounding box is recomputed.  Modifies the poses in place, returns None."""
    free = [True] * len(previous_poses)
    for pose in sorted(current_poses, key=lambda q: q.confidence, reverse=True):
        best_score, best_index = 0, None
        for index, previous in enumerate(previous_poses):
            if free[index]:
                score = get_similarity(pose, previous)
                if score > best_score:
                    best_score, best_index = score, index
        matched = best_index is not None and best_score >= threshold
        if matched:
            free[best_index] = False
        pose.update_id(previous_poses[best_index].id if matched else None)
        if smooth:
            # (the reference tests `best_matched_pose_id is not None`, which it resets when the match is too weak)
            for k in range(Pose.num_kpts):
                if pose.keypoints[k, 0] == -1:
                    continue
                if matched and previous_poses[best_index].keypoints[k, 0] != -1:
                    pose.filters[k] = previous_poses[best_index].filters[k]
                pose.keypoints[k, 0] = pose.filters[k][0](pose.keypoints[k, 0])
                pose.keypoints[k, 1] = pose.filters[k][1](pose.keypoints[k, 1])
            pose.bbox = Pose.get_bbox(pose.keypoints)
