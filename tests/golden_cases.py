"""Shared access to the golden fixtures (tests/golden/*.npz) and the seeded inputs they were made from."""
import importlib.util
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))


def _mg():
    # make_golden imports refimport lazily inside main(); importing the module itself is side-effect free
    import sys
    if "make_golden" not in sys.modules:
        m = importlib.util.module_from_spec(_spec)
        sys.modules["make_golden"] = m
        _spec.loader.exec_module(m)
    return sys.modules["make_golden"]


def load(name):
    return np.load(os.path.join(HERE, "golden", name), allow_pickle=False)


def resize_cases():
    mg = _mg()
    return list(enumerate(mg.RESIZE_CASES))


def resize_input(idx, h, w, C):
    return _mg().resize_input(idx, h, w, C)


def postproc_cases():
    return list(_mg().POSTPROC_CASES)


def postproc_maps(case):
    return _mg().postproc_maps(case)


def net_cases():
    return list(_mg().NET_CASES)


def sha(a):
    return _mg().sha(a)


def pack_keypoints(by_type):
    return _mg().pack_keypoints(by_type)
