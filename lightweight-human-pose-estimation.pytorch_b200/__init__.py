"""lwpose_b200 -- B200-native (sm_100a) inference hot path of Lightweight OpenPose.

Drop-in mirror of the reference's interface for this path only:
  lwpose_b200.models.with_mobilenet.PoseEstimationWithMobileNet   (reference models/with_mobilenet.py:89-123)
  lwpose_b200.modules.keypoints.extract_keypoints / group_keypoints (reference modules/keypoints.py:16-201)
  lwpose_b200.demo.infer_fast                                      (reference demo.py:54-78)
  lwpose_b200.val.infer / normalize / pad_width                    (reference val.py:30-49,81-110)
plus the batched device pipeline `lwpose_b200.pipeline`.  All compute runs in hand-written CUDA kernels
reached through the C-ABI library declared in include/lwpose_b200.h; there is no CPU fallback.
"""
__version__ = "0.1.0"
