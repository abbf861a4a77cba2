"""ctypes binding of libdskd_b200.so (the C ABI declared in include/dskd_b200.h).

There is no fallback: if the shared library is missing or the device is not sm_100 the import /
first call raises.  Nothing here imports the CPU oracle.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libdskd_b200.so')
MAX_LEVELS = 8
ABI_VERSION = 2

OK, EINVAL, ECUDA, EUNSUPPORTED_ARCH, EINFEASIBLE = 0, -1, -2, -3, -4
LAYOUT_NCHW, LAYOUT_SNC = 0, 1
MASK_DECODE_V1, MASK_DECODE_V2 = 0, 1
RASTER_OWNER_EXCL, RASTER_BINARY_INCL, RASTER_AREA_INCL, RASTER_AREA_FGBK = 0, 1, 2, 3
REDUCTION_MEAN, REDUCTION_SUM = 1, 2


class DskdError(RuntimeError):
    pass


class Level(C.Structure):
    _fields_ = [('H', C.c_int32), ('W', C.c_int32), ('cell_offset', C.c_int64)]


_FP = C.c_void_p


class DsgfdMseArgs(C.Structure):
    _fields_ = [('layout', C.c_int32), ('num_levels', C.c_int32), ('N', C.c_int32), ('C', C.c_int32),
                ('levels', Level * MAX_LEVELS),
                ('d_student', _FP * MAX_LEVELS), ('d_teacher', _FP * MAX_LEVELS),
                ('d_grad_student', _FP * MAX_LEVELS), ('scale', C.c_float * MAX_LEVELS),
                ('cells_per_image', C.c_int64), ('d_owner', _FP), ('d_rows', _FP), ('d_energy', _FP),
                ('num_pairs', C.c_int32), ('d_cell_weight', _FP), ('d_loss', _FP)]


class DsgfdKlArgs(C.Structure):
    _fields_ = [('num_levels', C.c_int32), ('N', C.c_int32), ('C', C.c_int32),
                ('levels', Level * MAX_LEVELS),
                ('d_student', _FP * MAX_LEVELS), ('d_teacher', _FP * MAX_LEVELS),
                ('scale', C.c_float * MAX_LEVELS), ('temperature', C.c_float),
                ('cells_per_image', C.c_int64), ('d_owner', _FP), ('d_rows', _FP), ('d_grad_rows', _FP),
                ('num_pairs', C.c_int32), ('d_cell_weight', _FP), ('d_loss', _FP),
                ('layout', C.c_int32), ('d_workspace', _FP), ('workspace_bytes', C.c_int64)]


class DsgfdStepArgs(C.Structure):
    _fields_ = [('criterion', C.c_int32), ('mask_mode', C.c_int32), ('layout', C.c_int32),
                ('num_levels', C.c_int32), ('N', C.c_int32), ('C', C.c_int32),
                ('levels', Level * MAX_LEVELS),
                ('d_student', _FP * MAX_LEVELS), ('d_teacher', _FP * MAX_LEVELS),
                ('d_grad_student', _FP * MAX_LEVELS), ('scale', C.c_float * MAX_LEVELS),
                ('temperature', C.c_float), ('cells_per_image', C.c_int64),
                ('d_hs_student', _FP), ('d_hs_teacher', _FP), ('d_grad_hs_student', _FP),
                ('num_query_rows', C.c_int32),
                ('d_teacher_keepid', _FP), ('d_student_labels', _FP), ('d_prev_mask', _FP),
                ('num_classes', C.c_int32),
                ('d_boxes', _FP), ('d_box_start', _FP), ('d_gt_boxes', _FP), ('d_gt_start', _FP), ('d_img_hw', _FP),
                ('num_pairs', C.c_int32), ('max_boxes_per_image', C.c_int32),
                ('d_loss', _FP), ('d_matched_count', _FP), ('d_workspace', _FP), ('workspace_bytes', C.c_int64),
                ('ev_kernel_begin', _FP), ('ev_kernel_end', _FP)]


class QmemArgs(C.Structure):
    _fields_ = [('N', C.c_int32), ('C', C.c_int32), ('S', C.c_int64),
                ('d_memory', _FP), ('d_hs_teacher', _FP), ('num_query_rows', C.c_int32),
                ('d_keepid', _FP), ('d_scores', _FP), ('d_box_start', _FP),
                ('num_pairs', C.c_int32), ('max_per_image', C.c_int32), ('temperature', C.c_float),
                ('d_cell_weight', _FP), ('d_workspace', _FP), ('workspace_bytes', C.c_int64)]


CRIT_MSE, CRIT_KL = 0, 1
MODE_DECODE_V1, MODE_DECODE_V2, MODE_SG_OUT, MODE_FG_ONLY, MODE_FG_BK = 0, 1, 2, 3, 4

i32, i64, f32, vp = C.c_int32, C.c_int64, C.c_float, C.c_void_p

# name -> argtypes; every entry point declared in include/dskd_b200.h (tests/test_abi.py checks both ways)
SIGNATURES = {
    'dskd_abi_version': [],
    'dskd_check_device': [],
    'dskd_mask_rows': [i32, vp, vp, vp, vp, i32, i32, vp, vp],
    'dskd_mask_rows_bwd': [vp, vp, vp, vp, vp, vp, i32, i32, vp, vp],
    'dskd_select_prev_queries': [vp, i32, vp, i32, i32, vp, vp, vp],
    'dskd_raster_cells': [i32, vp, vp, vp, vp, vp, i32, i32, C.POINTER(Level), i32, i64, vp, vp],
    'dskd_dsgfd_mse_fwd_bwd': [C.POINTER(DsgfdMseArgs), vp],
    'dskd_dsgfd_mse_finish': [vp, vp, i32, i32, vp, vp, vp],
    'dskd_dsgfd_kl_fwd_bwd': [C.POINTER(DsgfdKlArgs), vp],
    'dskd_dsgfd_rows_finish': [i32, vp, vp, vp, vp, vp, vp, i32, i32, vp, vp, vp],
    'dskd_dsgfd_step': [C.POINTER(DsgfdStepArgs), vp],
    'dskd_bcdd_loss_and_grad': [vp, i32, i32, i32, i32, f32, f32, vp, i32, vp, vp, vp, vp, vp, vp],
    'dskd_bcdd_prototypes': [vp, vp, i32, vp, vp, vp, i32, vp, i32, i32, vp, vp],
    'dskd_bcdd_distance_loss': [vp, i32, i32, i32, i32, f32, f32, vp, vp, vp, vp],
    'dskd_bcdd_scatter_grad': [vp, vp, i32, vp, i32, i32, vp, vp],
    'dskd_cost_matrix': [vp, vp, i32, i32, i32, i32, i32, vp, vp, vp, vp, i32, f32, f32, f32, vp, vp],
    'dskd_assign_targets': [vp, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp],
    'dskd_teacher_decode': [vp, vp, i32, i32, i32, i32, vp, f32, i32, vp, vp, vp, vp, vp, vp, vp],
    'dskd_teacher_compact': [vp, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp],
    'dskd_qmem_cell_weights': [C.POINTER(QmemArgs), vp],
    'dskd_graph_instantiate_prioritized': [vp, i32, C.POINTER(vp), C.POINTER(i32), C.POINTER(i32)],
    'dskd_graph_launch': [vp, vp],
    'dskd_graph_exec_destroy': [vp],
    'dskd_lsap_f64': [vp, i32, i32, vp, vp],
    'dskd_lsap_batch_f32': [vp, i32, i32, i32, vp, vp, i32],
    'dskd_lsap_batch_device': [vp, i32, i32, i32, i32, vp, i32, vp, vp, vp],
    'dskd_mse_elementwise': [vp, vp, vp, i64, f32, vp, vp, vp, vp, vp],
    'dskd_elementwise_loss': [i32, f32, vp, vp, vp, i64, f32, vp, vp, vp, vp, vp],
    'dskd_kd_kl_rows': [vp, vp, i64, i32, i64, f32, vp, f32, vp, vp, vp, vp],
    'dskd_scale_inplace': [vp, i64, vp, vp],
    'dskd_f64_to_f32': [vp, vp, i32, f32, vp],
    'dskd_msda_forward': [vp, vp, i32, vp, vp, i32, i64, i32, i32, i64, i32, vp, vp],
    'dskd_msda_backward': [vp, vp, i32, vp, vp, vp, i32, i64, i32, i32, i64, i32, vp, vp, vp, vp],
    'dskd_ipc_export': [vp, vp, C.POINTER(i64)],
    'dskd_ipc_open': [vp, C.POINTER(vp)],
    'dskd_ipc_close': [vp],
    'dskd_peer_allreduce': [vp, i64, C.POINTER(vp), i32, i32, vp],
}

_lib = None


def load():
    """Load the shared library once; raise loudly when it is absent (no CPU / eager fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DskdError(
            f'{LIB_PATH} is missing: build it with `python -c "import __graft_entry__ as g; g.build()"` '
            f'(nvcc, sm_100a).  dskd_b200 has no CPU or eager-PyTorch fallback.')
    lib = C.CDLL(LIB_PATH)
    lib.dskd_last_error.restype = C.c_char_p
    lib.dskd_last_error.argtypes = []
    lib.dskd_launch_count.restype = C.c_uint64
    lib.dskd_launch_count.argtypes = []
    lib.dskd_dsgfd_step_workspace_bytes.restype = C.c_int64
    lib.dskd_dsgfd_step_workspace_bytes.argtypes = [C.c_int32, C.c_int64, C.c_int32, C.c_int32]
    lib.dskd_dsgfd_step_workspace_bytes_for.restype = C.c_int64
    lib.dskd_dsgfd_step_workspace_bytes_for.argtypes = [C.POINTER(DsgfdStepArgs)]
    lib.dskd_struct_size.restype = C.c_int64
    lib.dskd_struct_size.argtypes = [C.c_int32]
    for which, mirror in enumerate((Level, DsgfdMseArgs, DsgfdKlArgs, DsgfdStepArgs, QmemArgs)):
        if lib.dskd_struct_size(which) != C.sizeof(mirror):
            raise DskdError(f'{mirror.__name__}: ctypes mirror is {C.sizeof(mirror)} bytes, the library says '
                            f'{lib.dskd_struct_size(which)} (include/dskd_b200.h and _lib.py are out of step)')
    lib.dskd_dsgfd_kl_workspace_bytes.restype = C.c_int64
    lib.dskd_dsgfd_kl_workspace_bytes.argtypes = [C.c_int32, C.c_int32, C.POINTER(Level), C.c_int32]
    lib.dskd_qmem_workspace_bytes.restype = C.c_int64
    lib.dskd_qmem_workspace_bytes.argtypes = [C.c_int32, C.c_int64, C.c_int32, C.c_int32]
    lib.dskd_peer_buffer_floats.restype = C.c_int64
    lib.dskd_peer_buffer_floats.argtypes = [C.c_int64]
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = C.c_int
    if lib.dskd_abi_version() != ABI_VERSION:
        raise DskdError(f'ABI version mismatch: library {lib.dskd_abi_version()} != binding {ABI_VERSION}')
    _lib = lib
    return lib


def check(rc, what=''):
    if rc != OK:
        msg = load().dskd_last_error().decode('utf-8', 'replace')
        raise DskdError(f'{what} failed with status {rc}: {msg}')


_device_checked = set()


def require_device(t: torch.Tensor):
    """Every device entry point goes through here: CUDA tensor on an sm_100 device, or raise."""
    if not t.is_cuda:
        raise DskdError('dskd_b200 runs on CUDA (sm_100a) tensors only; there is no CPU path '
                        f'(got a tensor on {t.device}).')
    idx = t.device.index if t.device.index is not None else torch.cuda.current_device()
    if idx not in _device_checked:
        with torch.cuda.device(idx):
            check(load().dskd_check_device(), 'dskd_check_device')
        _device_checked.add(idx)
    return idx


def ptr(t):
    """Device address for a `void*` argument (ctypes converts the int; None is NULL)."""
    return None if t is None else t.data_ptr()


def _first_cuda_tensor(obj, depth=0):
    if isinstance(obj, torch.Tensor):
        return obj if obj.is_cuda else None
    if depth < 3 and isinstance(obj, (tuple, list)):
        for x in obj:
            t = _first_cuda_tensor(x, depth + 1)
            if t is not None:
                return t
    return None


def guarded(fn):
    """Decorator of the public entry points: the ctypes calls launch on the CUDA device that is CURRENT for the calling
    thread, so when the tensors live on another device the call runs under `torch.cuda.device(that device)` (what
    PyTorch's own ops do with their device guard).  Costs one scan of the arguments when no switch is needed."""
    import functools

    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        t = None
        for a in args:
            t = _first_cuda_tensor(a)
            if t is not None:
                break
        if t is None:
            for a in kwargs.values():
                t = _first_cuda_tensor(a)
                if t is not None:
                    break
        if t is not None and t.device.index != torch.cuda.current_device():
            with torch.cuda.device(t.device):
                return fn(*args, **kwargs)
        return fn(*args, **kwargs)
    return wrapper


def stream_of(t: torch.Tensor):
    """Raw handle of the current stream of `t`'s device (what PyTorch ops would launch on)."""
    idx = t.device.index
    return C.c_void_p(torch._C._cuda_getCurrentRawStream(idx if idx is not None else torch.cuda.current_device()))


def f32c(t: torch.Tensor):
    """Contiguous fp32 view / copy (the reference runs this path under force_fp32, head_il.py:411)."""
    if t.dtype != torch.float32:
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()


_levels_cache = {}


def levels_struct(shapes):
    """shapes: iterable of (H, W) -> (ctypes array of Level, cells_per_image); cached per shape tuple (read-only use)."""
    key = tuple((int(h), int(w)) for h, w in shapes)
    hit = _levels_cache.get(key)
    if hit is None:
        if len(_levels_cache) > 256:
            _levels_cache.clear()
        hit = _levels_cache[key] = _levels_struct(key)
    return hit


def _levels_struct(shapes):
    arr = (Level * MAX_LEVELS)()
    off = 0
    shapes = [(int(h), int(w)) for h, w in shapes]
    if not 0 < len(shapes) <= MAX_LEVELS:
        raise DskdError(f'between 1 and {MAX_LEVELS} feature levels are supported, got {len(shapes)}')
    for l, (h, w) in enumerate(shapes):
        arr[l].H, arr[l].W, arr[l].cell_offset = h, w, off
        off += h * w
    return arr, off
