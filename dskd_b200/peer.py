"""Prototype exchange over NVLink / NVSwitch peer memory: `dskd_peer_allreduce` behind
`dist.allreduce_prototypes`.

One process per GPU.  Every rank owns a zero-filled "symmetric" buffer (a torch allocation), exports it as a CUDA IPC
handle, the handles travel through `torch.distributed.all_gather_object`, and every rank maps the others' buffers
(`dskd_ipc_open`).  After that one kernel launch per call publishes the local [2, classes, C + 1] table, flags the peers,
waits for their flags and adds the tables up in rank order -- no NCCL kernel, no host round trip, capturable in a CUDA
graph.  Set-up is collective and happens on the first call for a table size (outside any graph capture); it is used only
when every rank sits on the same host, every pair of devices has peer access and the process group runs on NCCL --
otherwise the caller keeps the NCCL all-reduce.  DSKD_PROTO_TRANSPORT=nccl switches it off.

One exchange object serves every table of its size on a device, and its calls must be ordered on one stream at a time
(the slots and counters belong to the sequence of calls, like an NCCL communicator's operations).
"""
import ctypes as C
import os
import socket

import torch
import torch.distributed as dist

from . import _lib as L

MAX_WORLD = 16


class PeerExchange:
    """Symmetric buffers of one process group for tables of `table_floats` floats (a multiple of 4).  Build it with
    `PeerExchange.create` (collective)."""

    def __init__(self, table_floats, device, group, buffer, ptrs, bases):
        self.table_floats, self.device, self.group = int(table_floats), device, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.buffer, self._ptrs, self._bases = buffer, ptrs, bases

    @classmethod
    def create(cls, table_floats: int, device: torch.device, group=None):
        """Collective.  Every rank runs the same two object all-gathers whatever fails locally, so a rank on which CUDA
        IPC is not permitted cannot leave the others waiting; returns None on EVERY rank if any rank failed."""
        lib = L.load()
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        buffer, mine = None, (False, b'', 0)
        try:
            with torch.cuda.device(device):
                buffer = torch.zeros(lib.dskd_peer_buffer_floats(int(table_floats)), dtype=torch.float32, device=device)
                torch.cuda.synchronize(device)          # flags and counters are zero before anybody can see them
                handle, offset = (C.c_char * 64)(), C.c_int64(0)
                L.check(lib.dskd_ipc_export(buffer.data_ptr(), C.cast(handle, C.c_void_p), C.byref(offset)), 'dskd_ipc_export')
                mine = (True, bytes(handle.raw), int(offset.value))
        except Exception:                               # noqa: BLE001 -- reported through the vote below
            pass
        everyone = [None] * world
        dist.all_gather_object(everyone, mine, group=group)
        bases, ptrs, ok = [], (C.c_void_p * world)(), all(e[0] for e in everyone)
        if ok:
            try:
                with torch.cuda.device(device):
                    for r, (_, h, off) in enumerate(everyone):
                        if r == rank:
                            ptrs[r] = buffer.data_ptr()
                            continue
                        base = C.c_void_p()
                        L.check(lib.dskd_ipc_open(C.cast(C.create_string_buffer(h, 64), C.c_void_p), C.byref(base)),
                                'dskd_ipc_open')
                        bases.append(base)
                        ptrs[r] = base.value + off
            except Exception:                           # noqa: BLE001
                ok = False
        votes = [None] * world
        dist.all_gather_object(votes, bool(ok), group=group)      # also the barrier: every mapping exists after it
        if not all(votes):
            for b in bases:
                lib.dskd_ipc_close(b)
            return None
        return cls(table_floats, device, group, buffer, ptrs, bases)

    def allreduce(self, table: torch.Tensor):
        """Replace `table` by its sum over the ranks, in place, on the current stream."""
        if table.numel() != self.table_floats or table.dtype != torch.float32 or not table.is_contiguous():
            raise L.DskdError('peer exchange: table of the wrong size / dtype / layout')
        L.check(L.load().dskd_peer_allreduce(table.data_ptr(), self.table_floats, self._ptrs, self.world, self.rank,
                                             L.stream_of(table)), 'dskd_peer_allreduce')

    def close(self):
        lib = L.load()
        for b in self._bases:
            lib.dskd_ipc_close(b)
        self._bases = []


_exchanges = {}
_unavailable = set()


def _all_agree(flag: bool, group) -> bool:
    votes = [None] * dist.get_world_size(group)
    dist.all_gather_object(votes, bool(flag), group=group)
    return all(votes)


def _usable(table: torch.Tensor, group) -> bool:
    """Collective: every rank evaluates its own view and all must agree."""
    world = dist.get_world_size(group)
    ok = (os.environ.get('DSKD_PROTO_TRANSPORT', 'nvlink').lower() != 'nccl' and table.is_cuda and
          table.dtype == torch.float32 and table.is_contiguous() and table.numel() % 4 == 0 and
          table.data_ptr() % 16 == 0 and 1 < world <= MAX_WORLD and dist.get_backend(group) == 'nccl')
    info = [None] * world
    dist.all_gather_object(info, (socket.gethostname(), int(table.device.index if table.is_cuda else -1), bool(ok)), group=group)
    same_host = len({h for h, _, _ in info}) == 1
    distinct = len({d for _, d, _ in info}) == world
    ok = ok and same_host and distinct and all(f for _, _, f in info)
    if ok:
        me = table.device.index
        ok = all(d == me or torch.cuda.can_device_access_peer(me, d) for _, d, _ in info)
    return _all_agree(ok, group)


def exchange_for(table: torch.Tensor, group=None):
    """The PeerExchange for this table size on this device, created (collectively) on first use; None when the peer path
    cannot be used -- the decision is the same on every rank."""
    key = (table.numel(), str(table.device), id(group))
    ex = _exchanges.get(key)
    if ex is not None:
        return ex
    if key in _unavailable:
        return None
    if torch.cuda.is_current_stream_capturing():
        raise L.DskdError('peer exchange: the first synchronised BCDD call sets the NVLink buffers up (a collective with '
                          'host round trips): run the step once before capturing it in a CUDA graph')
    created = PeerExchange.create(table.numel(), table.device, group) if _usable(table, group) else None
    if created is None:                                     # e.g. CUDA IPC is not permitted in this container
        _unavailable.add(key)
        return None
    _exchanges[key] = created
    return created


def close_all():
    """Unmap the peers' buffers (call before destroying the process group)."""
    for ex in _exchanges.values():
        ex.close()
    _exchanges.clear()
    _unavailable.clear()
