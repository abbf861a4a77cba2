"""CPU-only: the C-ABI library loads, exports every symbol include/dskd_b200.h declares, the ctypes
binding table matches the header, and the host-only entry points validate their arguments."""
import ctypes
import os
import re

import pytest

from dskd_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'dskd_b200.h')


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(dskd_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = declared_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f'{n} is declared in the header but not exported'


def test_binding_table_matches_header():
    names = set(declared_functions())
    bound = set(_lib.SIGNATURES) | {'dskd_last_error', 'dskd_launch_count', 'dskd_dsgfd_step_workspace_bytes',
                                   'dskd_dsgfd_step_workspace_bytes_for', 'dskd_dsgfd_kl_workspace_bytes',
                                   'dskd_qmem_workspace_bytes', 'dskd_struct_size', 'dskd_peer_buffer_floats'}
    assert names == bound, (sorted(names - bound), sorted(bound - names))


def test_header_cites_reference_lines():
    src = open(HEADER).read()
    for cite in ('head_il.py:664', 'head_il.py:525', 'gfl_hungarian_assigner.py', 'mse_loss.py', 'kd_loss.py',
                 'match_cost.py'):
        assert cite.split(':')[0] in src


def test_abi_version_and_error_string():
    lib = _lib.load()
    assert lib.dskd_abi_version() == 2
    rc = lib.dskd_lsap_f64(None, 3, 3, None, None)
    assert rc == _lib.EINVAL
    assert b'dskd_lsap_f64' in lib.dskd_last_error()
    with pytest.raises(_lib.DskdError):
        _lib.check(rc, 'dskd_lsap_f64')


def test_struct_layout_matches_c():
    # DskdLevel is {int32, int32, int64}; the arg structs start with 4 (mse) / 3 (kl) int32 then the level table
    assert ctypes.sizeof(_lib.Level) == 16
    assert _lib.DsgfdMseArgs.levels.offset == 16
    assert _lib.DsgfdKlArgs.levels.offset == 16
    assert _lib.DsgfdMseArgs.d_student.offset == 16 + 16 * _lib.MAX_LEVELS
    assert _lib.DsgfdStepArgs.levels.offset == 24
    lib = _lib.load()
    assert lib.dskd_dsgfd_step_workspace_bytes(2, 22223, 50, 256) % 256 == 0
    for which, mirror in enumerate((_lib.Level, _lib.DsgfdMseArgs, _lib.DsgfdKlArgs, _lib.DsgfdStepArgs, _lib.QmemArgs)):
        assert lib.dskd_struct_size(which) == ctypes.sizeof(mirror), mirror.__name__
    assert lib.dskd_struct_size(99) == -1
    # 300 queries -> two blocks of 160 resident rows; 100 -> one block of 112
    assert lib.dskd_qmem_workspace_bytes(2, 22223, 256, 300) > lib.dskd_qmem_workspace_bytes(2, 22223, 256, 100) > 0


def test_peer_exchange_entry_points_validate_their_arguments():
    """Host-side checks of the NVLink prototype exchange (csrc/peer.cu): buffer sizing and argument validation happen
    before any CUDA call, so they run without a GPU."""
    lib = _lib.load()
    table = 2 * 80 * 257
    assert lib.dskd_peer_buffer_floats(table) == 64 + 2 * table           # control words + two slots
    assert lib.dskd_peer_buffer_floats(41121) == 64 + 2 * 41124           # slots are padded to 4 floats
    assert lib.dskd_peer_buffer_floats(0) == -1
    assert lib.dskd_ipc_export(None, None, None) == _lib.EINVAL
    assert b'dskd_ipc_export' in lib.dskd_last_error()
    assert lib.dskd_ipc_open(None, None) == _lib.EINVAL
    assert lib.dskd_ipc_close(None) == _lib.OK
    bufs = (ctypes.c_void_p * 2)(0x1000, 0x2000)
    assert lib.dskd_peer_allreduce(None, table, bufs, 2, 0, None) == _lib.EINVAL
    assert lib.dskd_peer_allreduce(0x1000, table, bufs, 17, 0, None) == _lib.EINVAL     # world above DSKD_PEER_MAX_WORLD
    assert lib.dskd_peer_allreduce(0x1000, table, bufs, 2, 2, None) == _lib.EINVAL      # rank outside the world
    assert lib.dskd_peer_allreduce(0x1000, table + 1, bufs, 2, 0, None) == _lib.EINVAL  # not a multiple of 4 floats
    bad = (ctypes.c_void_p * 2)(0x1000, 0)
    assert lib.dskd_peer_allreduce(0x1000, table, bad, 2, 0, None) == _lib.EINVAL       # a peer that was never mapped
