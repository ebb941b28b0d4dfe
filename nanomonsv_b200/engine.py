"""Python face of the CUDA engine: packs sequences into the byte-coded batch layout and calls the C ABI.

Nothing here computes alignments; every method ends in a call into libnmsv_sw.so and raises if that fails.
"""
from __future__ import annotations

import ctypes
import os
import weakref
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import ALN_DTYPE, READ_DTYPE, RECORD_DTYPE, SSW_RESULT_DTYPE, SV_DTYPE, Scoring, Stats
from .my_seq import encode as _encode


class EngineError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"libnmsv_sw error {code}: {message}")
        self.code = code


DEFAULT_SCORING = (2, -2, 3, 1)   # matrix_create("ACGT", 2, -2); ssw(..., 3, 1, ...)  (count_sread_by_alignment.py:139,148)


def _scoring(s) -> Scoring:
    if isinstance(s, Scoring):
        return s
    match, mismatch, gap_open, gap_extend = s
    return Scoring(int(match), int(mismatch), int(gap_open), int(gap_extend))


class AsciiSlice:
    """A sequence that still lies, as ASCII, inside a larger buffer (a block of the seam TSV): it is encoded and packed in
    one pass by nmsv_encode_pack when the SeqPack is finalized. `src` is a uint8 array over the block (kept alive here)."""
    __slots__ = ("src", "start", "length")

    def __init__(self, src: np.ndarray, start: int, length: int):
        self.src, self.start, self.length = src, start, length

    def __len__(self) -> int:
        return self.length


class SeqPack:
    """All sequences of one call in one uint8 code buffer: sequence i = codes[offsets[i] : offsets[i] + lens[i]]."""

    def __init__(self, dedup: bool = False):
        self._chunks: List[np.ndarray] = []
        self._lens: List[int] = []
        self._dedup = dedup
        self._seen: Dict[str, int] = {}
        self._frozen: Optional[Tuple[np.ndarray, np.ndarray, np.ndarray]] = None

    def add(self, seq: str) -> int:
        """Add an ASCII sequence (encoded with the case-insensitive ACGT + wildcard mapper); returns its index."""
        if self._dedup:
            hit = self._seen.get(seq)
            if hit is not None:
                return hit
        idx = self.add_codes(_encode(seq))
        if self._dedup:
            self._seen[seq] = idx
        return idx

    def add_ascii_slice(self, piece: AsciiSlice) -> int:
        assert self._frozen is None, "SeqPack already finalized"
        self._chunks.append(piece)
        self._lens.append(piece.length)
        return len(self._lens) - 1

    def add_absent(self, length: int) -> int:
        """A sequence of which only the length is supplied (offset NMSV_SEQ_ABSENT): the sequence of a read that is only
        counted because its mapq test already fails."""
        assert self._frozen is None, "SeqPack already finalized"
        self._chunks.append(None)
        self._lens.append(int(length))
        return len(self._lens) - 1

    def add_codes(self, codes: np.ndarray) -> int:
        assert self._frozen is None, "SeqPack already finalized"
        codes = np.ascontiguousarray(codes, dtype=np.uint8)
        self._chunks.append(codes)
        self._lens.append(int(codes.shape[0]))
        return len(self._lens) - 1

    def __len__(self) -> int:
        return len(self._lens)

    def length(self, idx: int) -> int:
        return self._lens[idx]

    def finalize(self) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        if self._frozen is None:
            # every sequence starts on a 16-byte boundary (filled with the wildcard code): the kernel stages reads
            # with 16-byte-granular bulk copies (TMA); unaligned offsets are legal but take the plain-load path
            lens = np.asarray(self._lens, dtype=np.uint32)
            absent = np.fromiter((c is None for c in self._chunks), dtype=bool, count=len(self._chunks))
            padded = np.where(absent, np.uint64(0), (lens.astype(np.uint64) + np.uint64(15)) & ~np.uint64(15))
            offsets = np.zeros(len(lens), dtype=np.uint64)
            if len(lens) > 1:
                np.cumsum(padded[:-1], out=offsets[1:])
            total = int(offsets[-1] + padded[-1]) if len(lens) else 0
            codes = np.full(total, 4, dtype=np.uint8)
            lazy: Dict[int, list] = {}          # id(source block) -> [(source, start, length, destination offset)]
            for off, chunk in zip(offsets.tolist(), self._chunks):
                if chunk is None:
                    continue
                if isinstance(chunk, AsciiSlice):
                    lazy.setdefault(id(chunk.src), []).append((chunk.src, chunk.start, chunk.length, off))
                else:
                    codes[off:off + len(chunk)] = chunk
            for items in lazy.values():         # one library call per source block: encode + pack in a single pass
                src = items[0][0]
                so = np.array([i[1] for i in items], dtype=np.uint64)
                ln = np.array([i[2] for i in items], dtype=np.uint32)
                do = np.array([i[3] for i in items], dtype=np.uint64)
                _lib.load().nmsv_encode_pack(src.ctypes.data, so.ctypes.data, ln.ctypes.data, do.ctypes.data, len(items),
                                             codes.ctypes.data)
            offsets[absent] = _lib.SEQ_ABSENT
            self._frozen = (codes, offsets, lens)
            self._chunks = []
        return self._frozen

    @staticmethod
    def from_arrays(codes: np.ndarray, offsets: np.ndarray, lens: np.ndarray) -> "SeqPack":
        p = SeqPack()
        p._frozen = (np.ascontiguousarray(codes, dtype=np.uint8), np.ascontiguousarray(offsets, dtype=np.uint64),
                     np.ascontiguousarray(lens, dtype=np.uint32))
        p._lens = [int(x) for x in p._frozen[2]]
        return p


class Plan:
    """A validation batch resident on the GPU (nmsv_plan_*): run() is the device-resident hot path."""

    def __init__(self, engine: "Engine", handle: ctypes.c_void_p, n_svs: int, n_reads: int):
        self._e, self._h, self.n_svs, self.n_reads = engine, handle, n_svs, n_reads
        engine._plans.add(self)

    def run(self) -> None:
        self._e._check(self._e._L.nmsv_plan_run(self._h))

    def download(self) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        checked = np.zeros(self.n_svs, dtype=np.int32)
        supporting = np.zeros(self.n_svs, dtype=np.int32)
        records = np.zeros(max(self.n_reads, 1), dtype=RECORD_DTYPE)
        n = ctypes.c_uint64(0)
        self._e._check(self._e._L.nmsv_plan_download(self._h, checked.ctypes.data, supporting.ctypes.data,
                                                     records.ctypes.data, len(records), ctypes.byref(n)))
        records = records[:n.value]
        return checked, supporting, records[np.argsort(records["read"], kind="stable")]

    def download_alns(self) -> Tuple[np.ndarray, np.ndarray]:
        alns = np.zeros((self.n_reads, 4), dtype=ALN_DTYPE)
        flags = np.zeros(self.n_reads, dtype=np.uint8)
        self._e._check(self._e._L.nmsv_plan_download_alns(self._h, alns.ctypes.data, flags.ctypes.data))
        return alns, flags

    def set_profiling(self, on: bool = True) -> None:
        self._e._check(self._e._L.nmsv_plan_set_profiling(self._h, 1 if on else 0))

    def kernel_times(self):
        """[(phase, strip height R, milliseconds)] of the launches of the last run (CUDA events on the stream)."""
        cap = 64
        ms = np.zeros(cap, dtype=np.float32)
        phase = np.zeros(cap, dtype=np.uint32)
        rclass = np.zeros(cap, dtype=np.uint32)
        n = ctypes.c_uint32(0)
        self._e._check(self._e._L.nmsv_plan_kernel_times(self._h, ms.ctypes.data, phase.ctypes.data,
                                                         rclass.ctypes.data, cap, ctypes.byref(n)))
        names = ["plan", "sw_forward", "decide", "sw_begin", "final"]
        return [(names[int(phase[i])], int(rclass[i]), float(ms[i])) for i in range(n.value)]

    def copy_counts_to_device(self, device_ptr: int) -> None:
        self._e._check(self._e._L.nmsv_plan_copy_counts_device(self._h, ctypes.c_void_p(device_ptr)))

    def stats(self) -> dict:
        st = Stats()
        self._e._check(self._e._L.nmsv_plan_stats(self._h, ctypes.byref(st)))
        return st.asdict()

    def close(self) -> None:
        # the plan's buffers live on its engine's context and stream: an engine closes its plans before it goes away
        # (Engine.close), so the handle is gone by then and this is a no-op
        if self._h:
            self._e._L.nmsv_plan_destroy(self._h)
            self._h = None
            self._e._plans.discard(self)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Engine:
    """One CUDA context per process and GPU (mirrors one worker of --processes, reference run.py:361-372)."""

    def __init__(self, device: int = 0):
        self._L = _lib.load()
        h = ctypes.c_void_p()
        rc = self._L.nmsv_create(ctypes.byref(h), int(device))
        if rc != 0:
            raise EngineError(rc, (self._L.nmsv_last_error(None) or b"").decode())
        self._h = h
        self.device = int(device)
        self._plans = weakref.WeakSet()      # live plans: destroyed before the context (they use its stream and pool)
        self.prune = os.environ.get("NMSV_PRUNE", "1") != "0"      # mirrors the context's "prune" option

    # ---- plumbing
    def _check(self, rc: int) -> None:
        if rc != 0:
            raise EngineError(rc, (self._L.nmsv_last_error(self._h) or b"").decode())

    def close(self) -> None:
        if getattr(self, "_h", None):
            for plan in list(self._plans):
                plan.close()
            self._L.nmsv_destroy(self._h)
            self._h = None

    def set_option(self, name: str, value: int) -> None:
        """nmsv_set_option: "prune" = 0 computes every alignment the reference runs (parity tests)."""
        self._check(self._L.nmsv_set_option(self._h, name.encode(), int(value)))
        if name == "prune":
            self.prune = bool(value)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream_ptr: Optional[int]) -> None:
        """Launch on an existing cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream); None = own stream."""
        self._check(self._L.nmsv_set_stream(self._h, ctypes.c_void_p(cuda_stream_ptr or 0)))

    def device_info(self) -> dict:
        sm, clk = ctypes.c_int32(), ctypes.c_int32()
        l2, mem = ctypes.c_int64(), ctypes.c_int64()
        name = ctypes.create_string_buffer(256)
        self._check(self._L.nmsv_device_info(self._h, ctypes.byref(sm), ctypes.byref(clk), ctypes.byref(l2),
                                             ctypes.byref(mem), name, 256))
        return {"name": name.value.decode(), "sm_count": sm.value, "sm_clock_khz": clk.value,
                "l2_bytes": l2.value, "global_mem_bytes": mem.value}

    def measure_dpx_peak(self) -> float:
        """Sustained DPX (VIADDMNMX.S16x2) lane-instructions per second of this GPU."""
        v = ctypes.c_double(0.0)
        self._check(self._L.nmsv_measure_dpx_peak(self._h, ctypes.byref(v)))
        return v.value

    # ---- operator seam: parasail.ssw, batched
    def ssw_batch(self, pack: SeqPack, tmpl_idx: Sequence[int], read_idx: Sequence[int],
                  scoring=DEFAULT_SCORING, want_begin: bool = True) -> np.ndarray:
        codes, offsets, lens = pack.finalize()
        t = np.ascontiguousarray(tmpl_idx, dtype=np.uint32)
        r = np.ascontiguousarray(read_idx, dtype=np.uint32)
        assert t.shape == r.shape
        out = np.zeros(len(t), dtype=SSW_RESULT_DTYPE)
        sc = _scoring(scoring)
        self._check(self._L.nmsv_ssw_batch(self._h, codes.ctypes.data, codes.size, offsets.ctypes.data,
                                           lens.ctypes.data, len(lens), t.ctypes.data, r.ctypes.data, len(t),
                                           ctypes.byref(sc), 1 if want_begin else 0, out.ctypes.data))
        return out

    # ---- function seam: ssw_check, batched (template + reverse complement, strand rule, 1-based)
    def ssw_check_batch(self, pack: SeqPack, tmpl_idx: Sequence[int], read_idx: Sequence[int],
                        tmpl_rc_idx: Optional[Sequence[int]] = None, scoring=DEFAULT_SCORING) -> np.ndarray:
        codes, offsets, lens = pack.finalize()
        t = np.ascontiguousarray(tmpl_idx, dtype=np.uint32)
        r = np.ascontiguousarray(read_idx, dtype=np.uint32)
        rc = None if tmpl_rc_idx is None else np.ascontiguousarray(tmpl_rc_idx, dtype=np.uint32)
        out = np.zeros(len(t), dtype=ALN_DTYPE)
        sc = _scoring(scoring)
        self._check(self._L.nmsv_ssw_check_batch(self._h, codes.ctypes.data, codes.size, offsets.ctypes.data,
                                                 lens.ctypes.data, len(lens), t.ctypes.data,
                                                 None if rc is None else rc.ctypes.data, r.ctypes.data, len(t),
                                                 ctypes.byref(sc), out.ctypes.data))
        return out

    # ---- stage seam
    def plan(self, pack: SeqPack, svs: np.ndarray, reads: np.ndarray, scoring=DEFAULT_SCORING,
             prune: Optional[bool] = None) -> Plan:
        """prune=False: a plan that computes every alignment of the reference (download_alns then holds all tuples)."""
        if prune is not None:
            self.set_option("prune", 1 if prune else 0)
        try:
            return self._plan(pack, svs, reads, scoring)
        finally:
            if prune is not None:
                self.set_option("prune", 1)

    def _plan(self, pack: SeqPack, svs: np.ndarray, reads: np.ndarray, scoring=DEFAULT_SCORING) -> Plan:
        codes, offsets, lens = pack.finalize()
        svs = np.ascontiguousarray(svs, dtype=SV_DTYPE)
        reads = np.ascontiguousarray(reads, dtype=READ_DTYPE)
        sc = _scoring(scoring)
        h = ctypes.c_void_p()
        self._check(self._L.nmsv_plan_create(self._h, ctypes.byref(h), codes.ctypes.data, codes.size,
                                             offsets.ctypes.data, lens.ctypes.data, len(lens), svs.ctypes.data,
                                             len(svs), reads.ctypes.data, len(reads), ctypes.byref(sc)))
        return Plan(self, h, len(svs), len(reads))

    def validate(self, pack: SeqPack, svs: np.ndarray, reads: np.ndarray,
                 scoring=DEFAULT_SCORING) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        """One-shot nmsv_validate_batch: host buffers in, (checked, supporting, records sorted by read) out."""
        codes, offsets, lens = pack.finalize()
        svs = np.ascontiguousarray(svs, dtype=SV_DTYPE)
        reads = np.ascontiguousarray(reads, dtype=READ_DTYPE)
        checked = np.zeros(len(svs), dtype=np.int32)
        supporting = np.zeros(len(svs), dtype=np.int32)
        records = np.zeros(max(len(reads), 1), dtype=RECORD_DTYPE)
        n = ctypes.c_uint64(0)
        sc = _scoring(scoring)
        self._check(self._L.nmsv_validate_batch(self._h, codes.ctypes.data, codes.size, offsets.ctypes.data,
                                                lens.ctypes.data, len(lens), svs.ctypes.data, len(svs),
                                                reads.ctypes.data, len(reads), ctypes.byref(sc),
                                                checked.ctypes.data, supporting.ctypes.data, records.ctypes.data,
                                                len(records), ctypes.byref(n)))
        records = records[:n.value]
        return checked, supporting, records[np.argsort(records["read"], kind="stable")]


_default_engine: Optional[Engine] = None


def pick_device() -> int:
    """The GPU of this process: NMSV_DEVICE, else LOCAL_RANK (torchrun: one process per GPU), else the worker's place in
    a multiprocessing pool -- the reference runs its `--processes N` shards in a multiprocessing.Pool (run.py:361-372),
    whose workers are numbered 1..N: worker k takes GPU (k - 1) mod #GPUs, so N = 8 workers spread over 8 GPUs."""
    import os
    for var in ("NMSV_DEVICE", "LOCAL_RANK"):
        if os.environ.get(var, "") != "":
            return int(os.environ[var])
    import multiprocessing
    ident = getattr(multiprocessing.current_process(), "_identity", ())
    n_dev = _lib.load().nmsv_device_count()
    if ident and n_dev > 0:
        return (int(ident[0]) - 1) % n_dev
    return 0


def default_engine() -> Engine:
    """Process-wide engine on the GPU chosen by pick_device()."""
    global _default_engine
    if _default_engine is None:
        _default_engine = Engine(pick_device())
    return _default_engine
