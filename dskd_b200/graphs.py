"""Replay a captured loss step with the library's launch policy: small kernels before the streaming kernel's CTAs
(include/dskd_b200.h, `dskd_graph_instantiate_prioritized`)."""
import ctypes as C

import torch

from . import _lib as L


class PrioritizedGraph:
    """Wraps a `torch.cuda.CUDAGraph(keep_graph=True)` after capture: instantiates it with per-node priorities and
    replays it on the current stream.  The torch graph object (and its memory pool) is kept alive here."""

    def __init__(self, graph: 'torch.cuda.CUDAGraph', big_grid_ctas: int = 1024):
        lib = L.load()
        self.graph = graph
        self._exec = C.c_void_p()
        small, big = C.c_int32(), C.c_int32()
        L.check(lib.dskd_graph_instantiate_prioritized(C.c_void_p(int(graph.raw_cuda_graph())), int(big_grid_ctas),
                                                       C.byref(self._exec), C.byref(small), C.byref(big)),
                'dskd_graph_instantiate_prioritized')
        self.num_small, self.num_big = small.value, big.value

    def replay(self):
        L.check(L.load().dskd_graph_launch(self._exec, C.c_void_p(torch.cuda.current_stream().cuda_stream)), 'dskd_graph_launch')

    def __del__(self):
        try:
            if self._exec:
                L.load().dskd_graph_exec_destroy(self._exec)
        except Exception:
            pass
