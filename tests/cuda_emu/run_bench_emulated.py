"""Dry run of bench.py's control flow in the GPU-less container (TEST INFRASTRUCTURE ONLY): torch.cuda is replaced by stubs and
the library by tests/cuda_emu/libtfbs_emu.so, so that the JSON contract line, the option handling and the multi-step bookkeeping
can be checked without a GPU.  Every number it prints is meaningless (host fibers, wall-clock 'events') and is never recorded."""
import os
import runpy
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


class _Event:
    def __init__(self, enable_timing=True):
        self.t = 0.0

    def record(self, stream=None):
        self.t = time.perf_counter()

    def elapsed_time(self, other):
        return (other.t - self.t) * 1e3


torch.cuda.is_available = lambda: True
torch.cuda.set_device = lambda d: None
torch.cuda.synchronize = lambda *a, **k: None
torch.cuda.Event = _Event
torch.cuda.ExternalStream = lambda ptr, device=None: None

if __name__ == "__main__":
    sys.argv = [os.path.join(ROOT, "bench.py")] + sys.argv[1:]
    runpy.run_path(os.path.join(ROOT, "bench.py"), run_name="__main__")
