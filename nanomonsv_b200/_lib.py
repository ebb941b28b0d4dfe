"""ctypes binding of libnmsv_sw.so (include/nmsv.h). Mirrors the role of the reference's ssw_lib.CSsw
(ssw_lib.py:95-199): load the shared library, declare prototypes, raise on failure.

There is deliberately NO fallback: if the CUDA library is missing or no GPU is usable, importing callers get a
RuntimeError that says so.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

from . import build as _build

NMSV_NONE = 0xFFFFFFFF
SEQ_ABSENT = 0xFFFFFFFFFFFFFFFF
MAPQ_NONE = -1
SV_SHORT_DEL_DUP = 1
SV_SBND = 2
E_INVALID, E_CUDA, E_RANGE, E_CAPACITY, E_INTERNAL = -1, -2, -3, -4, -5

# numpy mirrors of the C structs (all little-endian, naturally aligned, no padding)
ALN_DTYPE = np.dtype([("score", "<i4"), ("qstart", "<i4"), ("qend", "<i4"), ("tstart", "<i4"), ("tend", "<i4"),
                      ("strand", "<i4")])
SSW_RESULT_DTYPE = np.dtype([("score1", "<i4"), ("ref_begin1", "<i4"), ("ref_end1", "<i4"),
                             ("read_begin1", "<i4"), ("read_end1", "<i4")])
SV_DTYPE = np.dtype([("tmpl", "<u4", (4,)), ("tmpl_rc", "<u4", (4,)), ("read_begin", "<u4"), ("read_end", "<u4"),
                     ("score_min", "<i4", (2,)), ("qstart_max", "<i4", (2,)), ("qend_min", "<i4", (2,)),
                     ("margin", "<i4"), ("min_mapq", "<i4"), ("flags", "<u4"), ("reserved", "<u4")])
READ_DTYPE = np.dtype([("seq", "<u4"), ("sv", "<u4"), ("mapq1", "<i4"), ("mapq2", "<i4")])
SEAM_GROUP_DTYPE = np.dtype([("key_off", "<u8"), ("key_len", "<u4"), ("read_begin", "<u4"), ("read_end", "<u4"),
                             ("reserved", "<u4")])
SEAM_READ_DTYPE = np.dtype([("id_off", "<u8"), ("seq_off", "<u8"), ("id_len", "<u4"), ("seq_len", "<u4"),
                            ("mapq1", "<i4"), ("mapq2", "<i4")])
assert SEAM_GROUP_DTYPE.itemsize == 24 and SEAM_READ_DTYPE.itemsize == 32
RECORD_DTYPE = np.dtype([("read", "<u4"), ("sv", "<u4"), ("aln", ALN_DTYPE, (4,))])
assert SV_DTYPE.itemsize == 80 and READ_DTYPE.itemsize == 16 and RECORD_DTYPE.itemsize == 104


class Scoring(ctypes.Structure):
    _fields_ = [("match", ctypes.c_int32), ("mismatch", ctypes.c_int32), ("gap_open", ctypes.c_int32),
                ("gap_extend", ctypes.c_int32)]


class JumpParams(ctypes.Structure):
    _fields_ = [("match_score", ctypes.c_int32), ("mismatch_penalty", ctypes.c_int32), ("gap_cost", ctypes.c_int32),
                ("jump_cost", ctypes.c_int32), ("minimum_score", ctypes.c_int32)]


JUMP_RESULT_DTYPE = np.dtype([(n, "<i4") for n in ("found", "score", "i1_start", "i1_end", "i2_start", "i2_end",
                                                   "j1_start", "j1_end", "j2_start", "j2_end")] +
                             [("n_ops2", "<u4"), ("n_ops1", "<u4")])


class Stats(ctypes.Structure):
    _fields_ = [("cells_reference", ctypes.c_uint64), ("cells_executed", ctypes.c_uint64),
                ("cells_padded", ctypes.c_uint64), ("fwd_tasks", ctypes.c_uint64), ("rev_tasks", ctypes.c_uint64),
                ("kernel_launches", ctypes.c_uint32), ("n_classes", ctypes.c_uint32), ("h2d_bytes", ctypes.c_uint64),
                ("d2h_bytes", ctypes.c_uint64), ("workspace_bytes", ctypes.c_uint64),
                ("cells_dedup", ctypes.c_uint64), ("cells_pruned", ctypes.c_uint64), ("cells_begin", ctypes.c_uint64),
                ("fwd_tasks_queued", ctypes.c_uint64), ("reads_pruned", ctypes.c_uint32), ("reserved", ctypes.c_uint32)]

    def asdict(self) -> dict:
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


ABI_VERSION = 2

EXPORTS = ["nmsv_abi_version", "nmsv_device_count", "nmsv_create", "nmsv_destroy", "nmsv_last_error", "nmsv_set_stream", "nmsv_set_option",
           "nmsv_device_info", "nmsv_encode_ascii", "nmsv_encode_pack", "nmsv_encode_pack_padded", "nmsv_seam_scan", "nmsv_ssw_batch", "nmsv_ssw_check_batch", "nmsv_validate_batch",
           "nmsv_plan_create", "nmsv_plan_run", "nmsv_plan_download", "nmsv_plan_stats", "nmsv_plan_download_alns",
           "nmsv_plan_destroy", "nmsv_plan_set_profiling", "nmsv_plan_kernel_times", "nmsv_plan_copy_counts_device",
           "nmsv_measure_dpx_peak", "nmsv_sw_jump_batch", "nmsv_sw_jump_strings"]

_lib = None


def library_path() -> str:
    return _build.LIB_PATH


def load() -> ctypes.CDLL:
    """Load (building first if the sources are newer) the CUDA library and declare its prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("NMSV_LIB_PATH") or _build.LIB_PATH      # NMSV_LIB_PATH: experiment builds only
    if path == _build.LIB_PATH and _build.is_stale():
        try:
            _build.build()
        except Exception as exc:  # no silent fallback: say what is missing
            if not os.path.exists(path):
                raise RuntimeError(f"libnmsv_sw.so is not built and cannot be built here ({exc}); "
                                   "the validation engine has no CPU fallback") from exc
            # the sources are newer than the library and the rebuild failed: running the old code silently would make
            # every test of an edited kernel meaningless. Boxes without nvcc opt in explicitly.
            if os.environ.get("NMSV_ALLOW_STALE", "0") != "1":
                raise RuntimeError(f"libnmsv_sw.so is older than its sources and the rebuild failed ({exc}); "
                                   "set NMSV_ALLOW_STALE=1 to load the stale library anyway") from exc
            import warnings
            warnings.warn(f"loading a STALE libnmsv_sw.so (rebuild failed: {exc})", RuntimeWarning)
    try:
        L = ctypes.CDLL(path)
    except OSError as exc:
        raise RuntimeError(f"cannot load {path}: {exc}; the validation engine has no CPU fallback") from exc
    vp, i32, u32, u64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_uint32, ctypes.c_uint64
    L.nmsv_abi_version.restype = ctypes.c_int
    if L.nmsv_abi_version() != ABI_VERSION:
        raise RuntimeError(f"{path} has ABI version {L.nmsv_abi_version()}, this package needs {ABI_VERSION}")
    L.nmsv_device_count.restype = ctypes.c_int
    L.nmsv_create.argtypes = [ctypes.POINTER(vp), ctypes.c_int]
    L.nmsv_create.restype = ctypes.c_int
    L.nmsv_destroy.argtypes = [vp]
    L.nmsv_destroy.restype = None
    L.nmsv_last_error.argtypes = [vp]
    L.nmsv_last_error.restype = ctypes.c_char_p
    L.nmsv_set_option.argtypes = [vp, ctypes.c_char_p, ctypes.c_int]
    L.nmsv_set_option.restype = ctypes.c_int
    L.nmsv_set_stream.argtypes = [vp, vp]
    L.nmsv_set_stream.restype = ctypes.c_int
    L.nmsv_device_info.argtypes = [vp, ctypes.POINTER(i32), ctypes.POINTER(i32), ctypes.POINTER(ctypes.c_int64),
                                   ctypes.POINTER(ctypes.c_int64), ctypes.c_char_p, ctypes.c_size_t]
    L.nmsv_device_info.restype = ctypes.c_int
    L.nmsv_encode_ascii.argtypes = [ctypes.c_char_p, u64, vp]
    L.nmsv_encode_ascii.restype = None
    L.nmsv_encode_pack.argtypes = [vp, vp, vp, vp, u64, vp]
    L.nmsv_encode_pack.restype = None
    L.nmsv_encode_pack_padded.argtypes = [vp, vp, vp, vp, u64, vp]
    L.nmsv_encode_pack_padded.restype = None
    L.nmsv_seam_scan.argtypes = [vp, u64, u64, u64, u64, ctypes.c_int, vp, u32, vp, u32, ctypes.POINTER(u32), ctypes.POINTER(u32),
                                 ctypes.POINTER(u64), ctypes.POINTER(u64)]
    L.nmsv_seam_scan.restype = ctypes.c_int
    sc = ctypes.POINTER(Scoring)
    L.nmsv_ssw_batch.argtypes = [vp, vp, u64, vp, vp, u32, vp, vp, u64, sc, ctypes.c_int, vp]
    L.nmsv_ssw_batch.restype = ctypes.c_int
    L.nmsv_ssw_check_batch.argtypes = [vp, vp, u64, vp, vp, u32, vp, vp, vp, u64, sc, vp]
    L.nmsv_ssw_check_batch.restype = ctypes.c_int
    L.nmsv_validate_batch.argtypes = [vp, vp, u64, vp, vp, u32, vp, u32, vp, u32, sc, vp, vp, vp, u64,
                                      ctypes.POINTER(u64)]
    L.nmsv_validate_batch.restype = ctypes.c_int
    L.nmsv_plan_create.argtypes = [vp, ctypes.POINTER(vp), vp, u64, vp, vp, u32, vp, u32, vp, u32, sc]
    L.nmsv_plan_create.restype = ctypes.c_int
    L.nmsv_plan_run.argtypes = [vp]
    L.nmsv_plan_run.restype = ctypes.c_int
    L.nmsv_plan_download.argtypes = [vp, vp, vp, vp, u64, ctypes.POINTER(u64)]
    L.nmsv_plan_download.restype = ctypes.c_int
    L.nmsv_plan_stats.argtypes = [vp, ctypes.POINTER(Stats)]
    L.nmsv_plan_stats.restype = ctypes.c_int
    L.nmsv_plan_download_alns.argtypes = [vp, vp, vp]
    L.nmsv_plan_download_alns.restype = ctypes.c_int
    L.nmsv_plan_destroy.argtypes = [vp]
    L.nmsv_plan_destroy.restype = None
    L.nmsv_plan_set_profiling.argtypes = [vp, ctypes.c_int]
    L.nmsv_plan_set_profiling.restype = ctypes.c_int
    L.nmsv_plan_kernel_times.argtypes = [vp, vp, vp, vp, u32, ctypes.POINTER(u32)]
    L.nmsv_plan_kernel_times.restype = ctypes.c_int
    L.nmsv_plan_copy_counts_device.argtypes = [vp, vp]
    L.nmsv_plan_copy_counts_device.restype = ctypes.c_int
    L.nmsv_measure_dpx_peak.argtypes = [vp, ctypes.POINTER(ctypes.c_double)]
    L.nmsv_measure_dpx_peak.restype = ctypes.c_int
    L.nmsv_sw_jump_batch.argtypes = [vp, vp, u64, vp, vp, u32, vp, vp, vp, u32, ctypes.POINTER(JumpParams), vp, u32, vp,
                                     vp, u64]
    L.nmsv_sw_jump_batch.restype = ctypes.c_int
    L.nmsv_sw_jump_strings.argtypes = [vp, vp, vp, vp, vp, vp, u32, vp, vp, vp, vp, vp]
    L.nmsv_sw_jump_strings.restype = ctypes.c_int
    _lib = L
    return L
