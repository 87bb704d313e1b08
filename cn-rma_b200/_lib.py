"""ctypes binding of libcnrma_b200.so (include/cnrma_b200.h) and its in-tree build.

The library is the product: there is no Python, PyTorch-eager or CPU fallback.  If it is missing, or the
current device is not a B200-class GPU, calls raise.
"""
import ctypes as C
import os
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, "csrc")
# CNRMA_LIB: load another build of the library (A/B runs of compile-time variants; see profiles/); default: the in-tree one
LIB_PATH = os.environ.get("CNRMA_LIB") or os.path.join(_HERE, "libcnrma_b200.so")
SOURCES = ["cnrma_abi.cu", "cnrma_stage_a.cu", "cnrma_stage_a_list.cu", "cnrma_stage_a_bilinear.cu", "cnrma_tsdf_head.cu", "cnrma_stage_b.cu", "cnrma_backward.cu",
           "cnrma_handoff.cu", "cnrma_fusion.cu", "cnrma_exchange.cu"]
HEADERS = ["cnrma_common.cuh", "cnrma_internal.cuh", os.path.join("..", "..", "include", "cnrma_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "--fmad=false", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]

OK = 0
F32, BF16 = 0, 1
AGG_ACCUMULATE, AGG_MEAN, AGG_COUNT_F32 = 1, 2, 4
MARCH_NEUS, MARCH_DEPTH = 0, 1
PULL_LSU = 0x10000


class CnrmaError(RuntimeError):
    pass


class Grid(C.Structure):
    _fields_ = [("nx", C.c_int32), ("ny", C.c_int32), ("nz", C.c_int32), ("voxel_size", C.c_float),
                ("origin", C.c_float * 3)]


class Box(C.Structure):
    _fields_ = [("lo", C.c_int32 * 3), ("dim", C.c_int32 * 3)]


class Features(C.Structure):
    _fields_ = [("views", C.c_int32), ("channels", C.c_int32), ("height", C.c_int32), ("width", C.c_int32),
                ("dtype", C.c_int32), ("stride_c", C.c_int64), ("stride_y", C.c_int64), ("stride_x", C.c_int64),
                ("view_ptrs_host", C.POINTER(C.c_void_p))]


class RmaResult(C.Structure):
    _fields_ = [("rows", C.c_int64), ("weight_sum", C.c_double), ("mean", C.c_float), ("overflow", C.c_int32)]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(_CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compiles csrc/*.cu for sm_100a into cn-rma_b200/libcnrma_b200.so (in-tree, travels with gpurun).  The
    translation units are compiled in parallel (one nvcc per file, objects under csrc/_build/) and linked with
    nvcc -shared; an object is rebuilt when its source or any header is newer."""
    if not force and not needs_build():
        return LIB_PATH
    from concurrent.futures import ThreadPoolExecutor
    objdir = os.path.join(_CSRC, "_build")
    os.makedirs(objdir, exist_ok=True)
    nvcc = _nvcc()
    compile_flags = [f for f in NVCC_FLAGS if f != "-shared"]
    header_time = max(os.path.getmtime(os.path.join(_CSRC, h)) for h in HEADERS)

    def compile_one(src):
        path = os.path.join(_CSRC, src)
        obj = os.path.join(objdir, src[:-3] + ".o")
        if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(path), header_time):
            return obj
        cmd = [nvcc] + compile_flags + ["-c", path, "-o", obj]
        if verbose:
            print(" ".join(cmd))
        subprocess.run(cmd, check=True)
        return obj

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 1)) as pool:
        objs = list(pool.map(compile_one, SOURCES))
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB_PATH] + objs
    if verbose:
        print(" ".join(cmd))
    subprocess.run(cmd, check=True)
    return LIB_PATH


_lib = None
_lock = threading.Lock()

_SIGNATURES = {
    "cnrma_abi_version": (C.c_int, []),
    "cnrma_status_string": (C.c_char_p, [C.c_int]),
    "cnrma_last_cuda_error": (C.c_int, []),
    "cnrma_check_device": (C.c_int, []),
    "cnrma_reload_tuning": (None, []),
    "cnrma_project_views": (C.c_int, [C.POINTER(Grid), C.c_void_p, C.c_int64, C.c_int, C.c_float, C.c_int, C.c_int,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "cnrma_aggregate_views": (C.c_int, [C.POINTER(Grid), C.POINTER(Features), C.c_void_p, C.c_int64, C.c_float,
                                        C.c_uint32, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p,
                                        C.c_void_p]),
    "cnrma_aggregate_views_box": (C.c_int, [C.POINTER(Grid), C.POINTER(Box), C.POINTER(Features), C.c_void_p, C.c_int64,
                                            C.c_float, C.c_uint32, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p,
                                            C.c_int, C.c_void_p]),
    "cnrma_mark_rows": (C.c_int, [C.POINTER(Grid), C.POINTER(Box), C.c_void_p, C.c_int64, C.c_int, C.c_float, C.c_int, C.c_int,
                                  C.c_void_p, C.c_int, C.c_int64, C.c_void_p]),
    "cnrma_pull_rows": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p), C.c_void_p,
                                  C.c_int64, C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    "cnrma_pull_default_ctas": (C.c_int, []),
    "cnrma_aggregate_views_bilinear": (C.c_int, [C.POINTER(Grid), C.POINTER(Features), C.c_void_p, C.c_int64, C.c_float,
                                                 C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "cnrma_tsdf_head_scale": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64,
                                        C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]),
    "cnrma_tsdf_head_workspace_bytes": (C.c_size_t, [C.c_int]),
    "cnrma_tsdf_head_scale_backward": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64,
                                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float,
                                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "cnrma_aggregate_views_routed": (C.c_int, [C.POINTER(Grid), C.POINTER(Features), C.c_void_p, C.c_int64, C.c_float, C.c_int,
                                               C.c_int, C.c_int, C.POINTER(C.c_void_p), C.c_void_p]),
    "cnrma_finalize_routed": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p]),
    "cnrma_selftest_count_division": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p]),
    "cnrma_to_channels_last": (C.c_int, [C.POINTER(Features), C.c_void_p, C.c_void_p]),
    "cnrma_t_one": (C.c_float, [C.POINTER(Grid), C.c_double, C.c_int]),
    "cnrma_ray_parameters": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "cnrma_rma_workspace_bytes": (C.c_int, [C.POINTER(Grid), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int,
                                            C.POINTER(C.c_size_t)]),
    "cnrma_rma_march": (C.c_int, [C.POINTER(Grid), C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                  C.c_float, C.c_int, C.c_float, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p,
                                  C.c_void_p]),
    "cnrma_rma_fill": (C.c_int, [C.POINTER(Grid), C.c_void_p, C.POINTER(Features), C.c_int, C.c_float, C.c_int,
                                 C.c_float, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p,
                                 C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]),
    "cnrma_rma_scatter": (C.c_int, [C.POINTER(Grid), C.c_void_p, C.POINTER(Features), C.c_int, C.c_float, C.c_int,
                                    C.c_float, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "cnrma_aggregate_views_backward": (C.c_int, [C.POINTER(Grid), C.POINTER(Features), C.c_void_p, C.c_int64, C.c_float,
                                                 C.c_uint32, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]),
    "cnrma_rma_fill_backward": (C.c_int, [C.POINTER(Grid), C.POINTER(Features), C.c_int, C.c_int, C.c_float, C.c_int,
                                          C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int64,
                                          C.c_void_p]),
    "cnrma_handoff_workspace_bytes": (C.c_int, [C.c_int64, C.POINTER(C.c_size_t)]),
    "cnrma_mask_prefix": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p]),
    "cnrma_select_rows": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]),
    "cnrma_rma_fill_selected": (C.c_int, [C.POINTER(Grid), C.c_void_p, C.POINTER(Features), C.c_int, C.c_float, C.c_int,
                                          C.c_float, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]),
    "cnrma_sample_workspace_bytes": (C.c_int, [C.POINTER(C.c_size_t)]),
    "cnrma_sample_mask": (C.c_int, [C.c_int64, C.c_int64, C.c_uint64, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]),
    "cnrma_sample_mask_for_result": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_uint64, C.c_void_p, C.c_size_t, C.c_void_p,
                                               C.c_void_p]),
    "cnrma_quantize_workspace_bytes": (C.c_int, [C.c_int64, C.POINTER(C.c_size_t)]),
    "cnrma_quantize_mark": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_float, C.c_void_p, C.c_size_t, C.c_void_p,
                                      C.c_void_p]),
    "cnrma_quantize_compact": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_int64, C.c_float, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p]),
    "cnrma_tsdf_integrate": (C.c_int, [C.POINTER(Grid), C.c_void_p, C.c_int64, C.c_int, C.POINTER(C.c_void_p),
                                       C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_float,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "cnrma_rma_expand": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_void_p]),
}

EXPORTS = tuple(_SIGNATURES)


def load():
    """Loads the shared library (never builds implicitly on a GPU box: the .so ships with the tree)."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise CnrmaError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                                 "(there is no fallback path)")
            lib = C.CDLL(LIB_PATH)
            for name, (res, args) in _SIGNATURES.items():
                fn = getattr(lib, name)
                fn.restype = res
                fn.argtypes = args
            if lib.cnrma_abi_version() != 2:
                raise CnrmaError("libcnrma_b200.so ABI version mismatch")
            _lib = lib
    return _lib


def reload_tuning():
    """Re-reads the CNRMA_* tuning knobs from the environment (the library reads them once; tests flip them)."""
    load().cnrma_reload_tuning()


def check(status, what=""):
    if status == OK:
        return
    lib = load()
    msg = lib.cnrma_status_string(status).decode()
    if status == -5:
        msg += f" [cudaError {lib.cnrma_last_cuda_error()}]"
    raise CnrmaError(f"{what or 'cnrma'}: {msg} (status {status})")


def make_grid(voxel_dim, voxel_size, origin):
    g = Grid()
    g.nx, g.ny, g.nz = (int(v) for v in voxel_dim)
    g.voxel_size = float(voxel_size)
    g.origin[0], g.origin[1], g.origin[2] = (float(v) for v in origin)
    return g


def make_box(lo, dim):
    b = Box()
    for a in range(3):
        b.lo[a], b.dim[a] = int(lo[a]), int(dim[a])
    return b


def empty(*size, **kw):
    """torch.empty for device outputs and scratch.  With CNRMA_POISON_OUTPUTS set (test aid) the buffer is filled with
    0xFF bytes first -- NaN for floats, -1 for integers -- so that an element a kernel fails to write cannot hide
    behind stale but plausible data left in a recycled allocator block (tests/test_gpu_poison.py)."""
    import torch
    t = torch.empty(*size, **kw)
    if os.environ.get("CNRMA_POISON_OUTPUTS") and t.is_cuda and t.numel() > 0:
        t.view(torch.uint8).fill_(0xFF)
    return t


def empty_like(x):
    import torch
    t = torch.empty_like(x)
    if os.environ.get("CNRMA_POISON_OUTPUTS") and t.is_cuda and t.numel() > 0:
        t.fill_(float("nan") if t.is_floating_point() else -1)
    return t
