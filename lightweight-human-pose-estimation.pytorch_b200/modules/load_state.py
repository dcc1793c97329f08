"""Checkpoint loading with the reference's tolerant semantics (reference modules/load_state.py:4-15):
tensors whose key and shape match are taken from the checkpoint, the rest keep their initial values
(with a warning).  Works unchanged because the module mirror has the reference's state_dict layout."""
import collections


def load_state(net, checkpoint):
    source = checkpoint["state_dict"]
    merged = collections.OrderedDict()
    for key, value in net.state_dict().items():
        if key in source and source[key].size() == value.size():
            merged[key] = source[key]
        else:
            merged[key] = value
            print("[WARNING] Not found pre-trained parameters for {}".format(key))
    net.load_state_dict(merged)
