"""Opt-in CUDA-event timing of the dominant (streaming) kernel inside `dskd_dsgfd_step`.

bench.py turns this on for its timed region: the C side records the event pair on the launch stream right
before / after the streaming kernel, so the roofline number is measured live, on the device, in the same run.
"""
import torch

enabled = False
_pairs = []


def new_event_pair(device):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # torch creates the cudaEvent lazily on first record; record once so the raw handles exist
    with torch.cuda.device(device):
        e0.record()
        e1.record()
    _pairs.append((e0, e1))
    return e0.cuda_event, e1.cuda_event


def start():
    global enabled
    _pairs.clear()
    enabled = True


def stop():
    """Returns the per-call kernel durations in ms (call after a synchronize)."""
    global enabled
    enabled = False
    out = [a.elapsed_time(b) for a, b in _pairs]
    _pairs.clear()
    return out
