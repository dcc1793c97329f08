"""oracle/ -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement of the reference's algorithms for the inference hot path
(`PoseEstimationWithMobileNet.forward` -> cubic upsample -> `extract_keypoints`
-> `group_keypoints`).  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import this package,
and only as the checker / the timed CPU baseline -- never on the product path.

Parity pin: every function is checked against golden vectors produced by the
real reference (`tests/golden/make_golden.py`, run in the build container where
`/root/reference` and cv2 exist) -- see `tests/test_oracle_golden.py`.

  postproc.py  ctypes front-end of lwp_oracle.c (cubic resize, extract, group)
  net.py       torch-CPU fp32 functional restatement of the network forward
  build.py     gcc recipe for liblwp_oracle.so
"""
