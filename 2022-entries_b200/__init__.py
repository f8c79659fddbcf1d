"""b200msm -- host side of the B200 MSM engine (mirror of the ZPrize prize-1a Rust shim).

The reference harness is Rust (`multi_scalar_mult_init` / `multi_scalar_mult`,
P1A/storswiftlabs/src/lib.rs:43-98, P1A/yrrid/src/lib.rs:38-90; P1A = open-division/prize1-msm/prize1a-msm-gpu);
this image has no Rust toolchain, so the same two calls are mirrored here over the same `extern "C"` symbols
(include/b200msm.h).  The Rust binding a maintainer would add is in INTEGRATION.md.

Everything here is plumbing over `libb200msm.so` (hand-written sm_100a CUDA).  There is no CPU path: if the
library is missing or no CUDA device is usable, calls raise.

Layouts (arkworks in-memory): G1Affine 104 B (x, y 6xu64 Montgomery + infinity byte at 96), scalar 32 B
(4xu64 integer in [0, r)), G1Projective 144 B (Jacobian x, y, z).
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200MSM_LIB") or os.path.join(_HERE, "libb200msm.so")      # override: experiment builds (scripts/build_variants.py)
BLS12_377, BLS12_381 = 0, 1
BASES_HARNESS, BASES_DISTINCT = 0, 1
AFFINE_BYTES, SCALAR_BYTES, PROJECTIVE_BYTES = 104, 32, 144
PEER_HANDLE_BYTES = 320      # one rank's handle for the multi-GPU bucket reduce-scatter (b200msm_peer_export)
PARTIAL_BYTES = 224          # one GPU's partial result: XYZZ in the device representation (4 x 14 x 28-bit limbs)
XYZZ_ABI_BYTES = 192         # test entry points: XYZZ as 4 arkworks field elements
ENTRY_BYTES = {BLS12_377: 144, BLS12_381: 112}      # one base-table entry in device memory (csrc/grp.cuh)

_lib = None


class MsmError(RuntimeError):
    pass


class _RustError(ctypes.Structure):
    _fields_ = [("code", ctypes.c_int), ("message", ctypes.c_void_p)]


class _FfiContext(ctypes.Structure):       # #[repr(C)] struct MultiScalarMultContext { context: *mut c_void }
    _fields_ = [("context", ctypes.c_void_p)]


def lib():
    """Load libb200msm.so (built in-tree by __graft_entry__.build() / make -C csrc)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MsmError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(nvcc, sm_100a).  There is no CPU fallback.")
    L = ctypes.CDLL(LIB_PATH)
    vp, u64, i32, sz = ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int, ctypes.c_size_t
    L.b200msm_create.argtypes = [ctypes.POINTER(vp), i32, i32, i32, i32]
    L.b200msm_destroy.argtypes = [vp]
    L.b200msm_last_error.argtypes = [vp]
    L.b200msm_last_error.restype = ctypes.c_char_p
    L.b200msm_set_stream.argtypes = [vp, vp]
    L.b200msm_set_bucket_arithmetic.argtypes = [vp, i32]
    L.b200msm_set_host_chunks.argtypes = [vp, i32, i32]
    L.b200msm_set_pipelined.argtypes = [vp, i32]
    L.b200msm_result_stream.argtypes = [vp]
    L.b200msm_result_stream.restype = vp
    L.b200msm_set_bases.argtypes = [vp, vp, u64, u64]
    L.b200msm_set_bases_generated.argtypes = [vp, u64, u64, i32, u64]
    L.b200msm_get_bases.argtypes = [vp, vp, u64, u64]
    L.b200msm_generate_scalars.argtypes = [vp, vp, u64, u64]
    L.b200msm_msm_device.argtypes = [vp, vp, u64, vp]
    L.b200msm_msm_host.argtypes = [vp, vp, u64, u64, vp]
    L.b200msm_msm_partial_device.argtypes = [vp, vp, u64, vp]
    L.b200msm_combine_device.argtypes = [vp, vp, i32, vp]
    L.b200msm_msm_partial_host.argtypes = [vp, vp, u64, vp]
    L.b200msm_sync.argtypes = [vp]
    L.b200msm_peer_export.argtypes = [vp, i32, vp]
    L.b200msm_peer_attach.argtypes = [vp, i32, i32, vp]
    L.b200msm_peer_detach.argtypes = [vp]
    L.b200msm_peer_info.argtypes = [vp, ctypes.POINTER(u64)]
    L.b200msm_msm_peer_partial_device.argtypes = [vp, vp, u64, vp]
    L.b200msm_msm_peer_partial_host.argtypes = [vp, vp, u64, vp]
    L.b200msm_peer_accumulate_device.argtypes = [vp, vp, u64]
    L.b200msm_peer_reduce.argtypes = [vp, vp]
    L.b200msm_check_bases_in_g1.argtypes = [vp, ctypes.POINTER(u64)]
    L.b200msm_variable_base_msm.argtypes = [i32, i32, vp, u64, u64, vp, u64, vp]
    L.b200msm_variable_base_release.argtypes = []
    L.mult_pippenger.argtypes = [vp, vp, sz, vp]
    L.mult_pippenger.restype = _RustError
    L.b200msm_get_info.argtypes = [vp, ctypes.POINTER(u64)]
    L.b200msm_plan_info.argtypes = [i32, u64, u64, i32, i32, i32, ctypes.POINTER(u64)]
    L.b200msm_get_stage_ms.argtypes = [vp, ctypes.POINTER(ctypes.c_float)]
    L.b200msm_reset_stage_stats.argtypes = [vp]
    L.b200msm_get_stage_ms_sum.argtypes = [vp, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(u64)]
    L.b200msm_test_fp_op.argtypes = [i32, i32, vp, vp, vp, u64]
    L.b200msm_test_ec_op.argtypes = [i32, i32, vp, vp, u64, i32]
    L.b200msm_test_grp_op.argtypes = [i32, i32, vp, vp, vp, u64, i32]
    L.mult_pippenger_init.argtypes = [ctypes.POINTER(_FfiContext), vp, sz, sz]
    L.mult_pippenger_init.restype = _RustError
    L.mult_pippenger_inf.argtypes = [ctypes.POINTER(_FfiContext), vp, vp, sz, sz, vp, sz]
    L.mult_pippenger_inf.restype = _RustError
    L.mult_pippenger_free.argtypes = [ctypes.POINTER(_FfiContext)]
    L.mult_pippenger_free.restype = _RustError
    L.MSMAllocContext.argtypes = [ctypes.c_int32, ctypes.c_int32]
    L.MSMAllocContext.restype = vp
    L.MSMFreeContext.argtypes = [vp]
    L.MSMFreeContext.restype = ctypes.c_int32
    L.MSMPreprocessPoints.argtypes = [vp, vp, ctypes.c_uint32]
    L.MSMPreprocessPoints.restype = ctypes.c_int32
    L.MSMRun.argtypes = [vp, vp, vp, ctypes.c_uint32]
    L.MSMRun.restype = ctypes.c_int32
    _lib = L
    return L


EXPORTED_SYMBOLS = [
    "b200msm_create", "b200msm_destroy", "b200msm_last_error", "b200msm_set_stream", "b200msm_set_bucket_arithmetic", "b200msm_set_host_chunks", "b200msm_set_pipelined", "b200msm_result_stream", "b200msm_set_bases",
    "b200msm_set_bases_generated", "b200msm_get_bases", "b200msm_generate_scalars", "b200msm_msm_device",
    "b200msm_msm_host", "b200msm_msm_partial_device", "b200msm_msm_partial_host", "b200msm_combine_device", "b200msm_sync",
    "b200msm_peer_export", "b200msm_peer_attach", "b200msm_peer_detach", "b200msm_peer_info", "b200msm_msm_peer_partial_device", "b200msm_msm_peer_partial_host",
    "b200msm_peer_accumulate_device", "b200msm_peer_reduce", "b200msm_check_bases_in_g1", "b200msm_get_info", "b200msm_plan_info",
    "b200msm_get_stage_ms", "b200msm_reset_stage_stats", "b200msm_get_stage_ms_sum", "b200msm_test_fp_op", "b200msm_test_ec_op", "b200msm_test_grp_op", "mult_pippenger_init", "mult_pippenger_inf",
    "mult_pippenger_free", "b200msm_variable_base_msm", "b200msm_variable_base_release", "mult_pippenger", "MSMAllocContext", "MSMFreeContext", "MSMPreprocessPoints", "MSMRun",
]


def _as_u8(a, row_bytes, what):
    a = np.ascontiguousarray(a)
    flat = a.view(np.uint8).reshape(-1)
    if flat.size % row_bytes:
        raise ValueError(f"{what}: byte length {flat.size} is not a multiple of {row_bytes}")
    return flat


def _ptr(a):
    return ctypes.c_void_p(a.ctypes.data)


def _take_rust_error(err, what):
    if err.code != 0:
        msg = ctypes.string_at(err.message).decode() if err.message else ""
        if err.message:
            ctypes.CDLL(None).free(ctypes.c_void_p(err.message))     # the Rust side owns and frees the message
        raise MsmError(f"{what}: error {err.code}: {msg}")           # the Rust shim panics here


# ---------------------------------------------------------------------------------------------------------------
# The harness API (Family A symbols), same names and argument meaning as the Rust shim
class MultiScalarMultContext:
    def __init__(self):
        self.ffi = _FfiContext(None)
        self.npoints = 0

    @classmethod
    def from_engine(cls, engine):
        """A harness context over an existing native engine (bench.py: bases generated on the device instead of
        6.5 GiB of host records).  The engine keeps ownership."""
        ctx = cls()
        ctx.ffi.context = engine.h.value
        ctx.npoints = engine.n
        ctx.borrowed = True
        return ctx

    def free(self):
        if self.ffi.context and not getattr(self, "borrowed", False):
            lib().mult_pippenger_free(ctypes.byref(self.ffi))
        self.ffi.context = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def multi_scalar_mult_init(points):
    """points: n arkworks G1Affine records (uint8[n, 104]).  Uploads them and builds the base tables (untimed
    in the harness)."""
    pts = _as_u8(points, AFFINE_BYTES, "points")
    ctx = MultiScalarMultContext()
    ctx.npoints = pts.size // AFFINE_BYTES
    err = lib().mult_pippenger_init(ctypes.byref(ctx.ffi), _ptr(pts), ctx.npoints, AFFINE_BYTES)
    _take_rust_error(err, "mult_pippenger_init")
    return ctx


def multi_scalar_mult(context, points, scalars):
    """scalars: batches * n 32-byte integers.  Returns uint8[batches, 144] Jacobian results.
    points may be None for a context made by from_engine (the C side ignores the argument, as the reference does)."""
    sc = _as_u8(scalars, SCALAR_BYTES, "scalars")
    if points is None:
        pts, npoints = np.zeros(1, np.uint8), context.npoints
    else:
        pts = _as_u8(points, AFFINE_BYTES, "points")
        npoints = pts.size // AFFINE_BYTES
    nscalars = sc.size // SCALAR_BYTES
    if npoints == 0 or nscalars % npoints != 0:
        raise MsmError("length mismatch")                            # the Rust shim panics with the same message
    batches = nscalars // npoints
    out = np.zeros((batches, PROJECTIVE_BYTES), np.uint8)
    err = lib().mult_pippenger_inf(ctypes.byref(context.ffi), _ptr(out), _ptr(pts), npoints, batches, _ptr(sc), AFFINE_BYTES)
    _take_rust_error(err, "mult_pippenger_inf")
    return out


# ---------------------------------------------------------------------------------------------------------------
# The one-call form: arkworks' own trait function and the prover-side callers of it
class VariableBaseMSM:
    """`ark_ec::msm::VariableBaseMSM` (ark-ec 0.3.0; restated in the reference at P1A/speakspeak/src/util.rs:51-174):
    one static function, bases and scalars in, one projective point out, no context."""

    @staticmethod
    def multi_scalar_mul(bases, scalars, curve=BLS12_377, device=0, stride=AFFINE_BYTES):
        """bases: G1Affine records (`stride` bytes each: 104 arkworks, 96 without the infinity flag); scalars: 32-byte
        integers.  The first min(len(bases), len(scalars)) pairs are used, as arkworks does; empty input -> zero.
        Returns uint8[144] (Jacobian x, y, z)."""
        pts = _as_u8(bases, stride, "bases")
        sc = _as_u8(scalars, SCALAR_BYTES, "scalars")
        out = np.zeros(PROJECTIVE_BYTES, np.uint8)
        rc = lib().b200msm_variable_base_msm(curve, device, _ptr(pts) if pts.size else None, pts.size // stride, stride,
                                             _ptr(sc) if sc.size else None, sc.size // SCALAR_BYTES, _ptr(out))
        if rc != 0:
            raise MsmError(f"b200msm_variable_base_msm: {rc}: {lib().b200msm_last_error(None).decode()}")
        return out

    @staticmethod
    def release():
        """Free the device buffers the one-call form keeps between calls."""
        lib().b200msm_variable_base_release()


class VariableBase:
    """snarkVM's dispatcher (team-division/prize2-posw-gpu/storswiftlabs/snarkVM/algorithms/src/msm/variable_base/
    mod.rs:120-157): `VariableBase::msm(bases, scalars)`.  The reference tries the GPU and falls back to its CPU
    `batched::msm` when the GPU call fails; this mirror has no CPU path, a failed GPU call raises."""

    @staticmethod
    def msm(bases, scalars):
        return VariableBaseMSM.multi_scalar_mul(bases, scalars, BLS12_377, 0)


def multi_scalar_mult_blst(points, scalars):
    """upstream sppark `multi_scalar_mult(points: &[blst_p1_affine], scalars: &[blst_scalar]) -> blst_p1`
    (.../Jackytan2018/extern/sppark/sppark/poc/blst-cuda/src/lib.rs:38-53) over `mult_pippenger`: 96-byte points."""
    pts = _as_u8(points, 96, "points")
    sc = _as_u8(scalars, SCALAR_BYTES, "scalars")
    if pts.size // 96 != sc.size // SCALAR_BYTES:
        raise MsmError("length mismatch")                            # the Rust shim panics with the same message
    if pts.size == 0:
        raise MsmError("empty input")                                # the Rust shim indexes points[0] and panics
    out = np.zeros(PROJECTIVE_BYTES, np.uint8)
    err = lib().mult_pippenger(_ptr(out), _ptr(pts), pts.size // 96, _ptr(sc))
    _take_rust_error(err, "mult_pippenger")
    return out


# ---------------------------------------------------------------------------------------------------------------
# Family B (yrrid): thin wrappers, same call order as P1A/yrrid/src/lib.rs:38-90
class YrridContext:
    def __init__(self, max_points=1 << 26, max_batches=16):
        self.handle = lib().MSMAllocContext(max_points, max_batches)
        if not self.handle:
            raise MsmError("MSMAllocContext failed")

    def preprocess_points(self, points):
        pts = _as_u8(points, AFFINE_BYTES, "points")
        self.npoints = pts.size // AFFINE_BYTES
        rc = lib().MSMPreprocessPoints(self.handle, _ptr(pts), self.npoints)
        if rc != 0:
            raise MsmError(f"Error {rc} occurred in C code")

    def run(self, scalars):
        sc = _as_u8(scalars, SCALAR_BYTES, "scalars")
        n = sc.size // SCALAR_BYTES
        out = np.zeros((max(n // max(self.npoints, 1), 1), PROJECTIVE_BYTES), np.uint8)
        rc = lib().MSMRun(self.handle, _ptr(out), _ptr(sc), n)
        if rc != 0:
            raise MsmError(f"Error {rc} occurred in C code")
        return out

    def free(self):
        if self.handle:
            lib().MSMFreeContext(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


# ---------------------------------------------------------------------------------------------------------------
# Native interface: device-resident scalars, generated bases, partial results for point sharding, stage timings
class Engine:
    STAGES = ("digits_hist", "scan_scatter", "pieces", "accumulate", "bucket_reduce", "finalize")

    def __init__(self, curve=BLS12_377, device=0, window_bits=0, tables=0):
        self.h = ctypes.c_void_p()
        rc = lib().b200msm_create(ctypes.byref(self.h), curve, device, window_bits, tables)
        if rc != 0:
            raise MsmError(f"b200msm_create: {rc}: {lib().b200msm_last_error(None).decode()}")
        self.curve, self.device = curve, device

    def _ck(self, rc, what):
        if rc != 0:
            raise MsmError(f"{what}: {rc}: {lib().b200msm_last_error(self.h).decode()}")

    def close(self):
        if self.h:
            lib().b200msm_destroy(self.h)
            self.h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream):
        self._ck(lib().b200msm_set_stream(self.h, ctypes.c_void_p(cuda_stream)), "set_stream")

    def set_bucket_arithmetic(self, xyzz):
        """xyzz=True: short-Weierstrass XYZZ buckets (complete on every on-curve point); False: the curve's default."""
        self._ck(lib().b200msm_set_bucket_arithmetic(self.h, 1 if xyzz else 0), "set_bucket_arithmetic")

    def set_pipelined(self, on=True):
        """Consecutive device-resident MSMs overlap: results are then ordered on result_stream() (see include/b200msm.h)."""
        self._ck(lib().b200msm_set_pipelined(self.h, 1 if on else 0), "set_pipelined")

    def result_stream(self):
        """Raw cudaStream_t (int) on which results of the device-resident calls are ordered."""
        return int(lib().b200msm_result_stream(self.h) or 0)

    def set_host_chunks(self, chunks=0, ratio=0):
        """Split of the first batch of a host-memory call: `chunks` pieces of sizes 1 : ratio : ratio^2 ... (0 = defaults)."""
        self._ck(lib().b200msm_set_host_chunks(self.h, chunks, ratio), "set_host_chunks")

    def set_bases(self, points, stride=AFFINE_BYTES):
        pts = _as_u8(points, stride, "points")
        self.n = pts.size // stride
        self._ck(lib().b200msm_set_bases(self.h, _ptr(pts), self.n, stride), "set_bases")

    def set_bases_generated(self, n, seed=1, mode=BASES_HARNESS, distinct=1 << 15):
        self.n = n
        self._ck(lib().b200msm_set_bases_generated(self.h, n, seed, mode, distinct), "set_bases_generated")

    def get_bases(self, first, n):
        out = np.zeros((n, 96), np.uint8)
        self._ck(lib().b200msm_get_bases(self.h, _ptr(out), first, n), "get_bases")
        return out

    def generate_scalars(self, dev_ptr, n, seed):
        self._ck(lib().b200msm_generate_scalars(self.h, ctypes.c_void_p(dev_ptr), n, seed), "generate_scalars")

    def msm_device(self, scalars_dev_ptr, n, out_dev_ptr):
        self._ck(lib().b200msm_msm_device(self.h, ctypes.c_void_p(scalars_dev_ptr), n, ctypes.c_void_p(out_dev_ptr)), "msm_device")

    def msm_partial_device(self, scalars_dev_ptr, n, out_xyzz_dev_ptr):
        self._ck(lib().b200msm_msm_partial_device(self.h, ctypes.c_void_p(scalars_dev_ptr), n, ctypes.c_void_p(out_xyzz_dev_ptr)), "msm_partial_device")

    def msm_partial_host(self, scalars_host_ptr, n, out_xyzz_dev_ptr):
        """scalars_host_ptr: address of n * 32 bytes of (pinned) host memory, valid until sync()."""
        self._ck(lib().b200msm_msm_partial_host(self.h, ctypes.c_void_p(scalars_host_ptr), n, ctypes.c_void_p(out_xyzz_dev_ptr)), "msm_partial_host")

    # ---- multi-GPU bucket reduce-scatter (include/b200msm.h "b200msm_peer_*")
    def peer_export(self, chunks=1):
        h = np.zeros(PEER_HANDLE_BYTES, np.uint8)
        self._ck(lib().b200msm_peer_export(self.h, chunks, _ptr(h)), "peer_export")
        return h

    def peer_attach(self, rank, world, handles):
        hs = _as_u8(handles, PEER_HANDLE_BYTES, "handles")
        if hs.size != world * PEER_HANDLE_BYTES:
            raise MsmError("peer_attach: one handle per rank")
        self._ck(lib().b200msm_peer_attach(self.h, rank, world, _ptr(hs)), "peer_attach")

    def peer_detach(self):
        self._ck(lib().b200msm_peer_detach(self.h), "peer_detach")

    def peer_info(self):
        a = (ctypes.c_uint64 * 6)()
        lib().b200msm_peer_info(self.h, a)
        return dict(zip(("world", "rank", "chunks", "nvlink_bytes", "last_wait_us", "last_exchange_us"), [int(v) for v in a]))

    def msm_peer_partial_device(self, scalars_dev_ptr, n, out_xyzz_dev_ptr):
        self._ck(lib().b200msm_msm_peer_partial_device(self.h, ctypes.c_void_p(scalars_dev_ptr), n, ctypes.c_void_p(out_xyzz_dev_ptr)), "msm_peer_partial_device")

    def msm_peer_partial_host(self, scalars_host_ptr, n, out_xyzz_dev_ptr):
        self._ck(lib().b200msm_msm_peer_partial_host(self.h, ctypes.c_void_p(scalars_host_ptr), n, ctypes.c_void_p(out_xyzz_dev_ptr)), "msm_peer_partial_host")

    def peer_accumulate_device(self, scalars_dev_ptr, n):
        self._ck(lib().b200msm_peer_accumulate_device(self.h, ctypes.c_void_p(scalars_dev_ptr), n), "peer_accumulate_device")

    def peer_reduce(self, out_xyzz_dev_ptr):
        self._ck(lib().b200msm_peer_reduce(self.h, ctypes.c_void_p(out_xyzz_dev_ptr)), "peer_reduce")

    def check_bases_in_g1(self):
        """Number of resident bases outside the prime-order subgroup (r * P != infinity); opt-in, ~380 curve operations per base."""
        k = ctypes.c_uint64()
        self._ck(lib().b200msm_check_bases_in_g1(self.h, ctypes.byref(k)), "check_bases_in_g1")
        return int(k.value)

    def combine_device(self, xyzz_dev_ptr, count, out_dev_ptr):
        self._ck(lib().b200msm_combine_device(self.h, ctypes.c_void_p(xyzz_dev_ptr), count, ctypes.c_void_p(out_dev_ptr)), "combine_device")

    def msm_host(self, scalars, batches=1):
        """scalars: host buffer (numpy array or a raw address of pinned memory) of batches * n * 32 bytes."""
        out = np.zeros((batches, PROJECTIVE_BYTES), np.uint8)
        if isinstance(scalars, int):
            p = ctypes.c_void_p(scalars)
        else:
            sc = _as_u8(scalars, SCALAR_BYTES, "scalars")
            if sc.size != batches * self.n * SCALAR_BYTES:
                raise MsmError("length mismatch")
            p = _ptr(sc)
        self._ck(lib().b200msm_msm_host(self.h, p, self.n, batches, _ptr(out)), "msm_host")
        return out

    def sync(self):
        self._ck(lib().b200msm_sync(self.h), "sync")

    def info(self):
        a = (ctypes.c_uint64 * 8)()
        lib().b200msm_get_info(self.h, a)
        keys = ("n", "window_bits", "windows", "tables", "folded_windows", "buckets", "lmax", "launches")
        return dict(zip(keys, [int(v) for v in a]))

    def reset_stage_stats(self):
        self._ck(lib().b200msm_reset_stage_stats(self.h), "reset_stage_stats")

    def stage_ms_sum(self):
        """(per-stage milliseconds summed over the MSMs since reset_stage_stats, number of MSMs covered)."""
        a = (ctypes.c_float * 6)()
        k = ctypes.c_uint64()
        self._ck(lib().b200msm_get_stage_ms_sum(self.h, a, ctypes.byref(k)), "get_stage_ms_sum")
        return dict(zip(self.STAGES, [float(v) for v in a])), int(k.value)

    def stage_ms(self):
        a = (ctypes.c_float * 6)()
        self._ck(lib().b200msm_get_stage_ms(self.h, a), "get_stage_ms")
        return dict(zip(self.STAGES, [float(v) for v in a]))


def plan_info(n, curve=BLS12_377, free_bytes=170 << 30, window_bits=0, tables=0, chunks=1):
    """The window plan for n points (host arithmetic only; works without a GPU)."""
    a = (ctypes.c_uint64 * 8)()
    rc = lib().b200msm_plan_info(curve, n, free_bytes, window_bits, tables, chunks, a)
    if rc != 0:
        raise MsmError(f"b200msm_plan_info: {rc}")
    keys = ("n", "window_bits", "windows", "tables", "folded_windows", "buckets", "table_bytes", "refs")
    return dict(zip(keys, [int(v) for v in a]))


def test_fp_op(curve, op, a, b=None):
    a = _as_u8(a, 48, "a")
    b = a if b is None else _as_u8(b, 48, "b")
    r = np.zeros_like(a)
    rc = lib().b200msm_test_fp_op(curve, op, _ptr(a), _ptr(b), _ptr(r), a.size // 48)
    if rc != 0:
        raise MsmError(f"b200msm_test_fp_op: {rc}: {lib().b200msm_last_error(None).decode()}")
    return r.reshape(-1, 48)


def test_ec_op(curve, op, acc, other=None, reps=1):
    acc = _as_u8(acc, XYZZ_ABI_BYTES, "acc").copy()
    o = None if other is None else _as_u8(other, 96 if op != 1 else XYZZ_ABI_BYTES, "other")
    rc = lib().b200msm_test_ec_op(curve, op, _ptr(acc), _ptr(o) if o is not None else None, acc.size // XYZZ_ABI_BYTES, reps)
    if rc != 0:
        raise MsmError(f"b200msm_test_ec_op: {rc}: {lib().b200msm_last_error(None).decode()}")
    return acc.reshape(-1, XYZZ_ABI_BYTES)


def test_grp_op(curve, op, a, b, reps=1):
    """a, b: uint8[n, 96] arkworks (x, y); returns uint8[n, 104] G1Affine records."""
    a, b = _as_u8(a, 96, "a"), _as_u8(b, 96, "b")
    n = a.size // 96
    out = np.zeros((n, AFFINE_BYTES), np.uint8)
    rc = lib().b200msm_test_grp_op(curve, op, _ptr(a), _ptr(b), _ptr(out), n, reps)
    if rc != 0:
        raise MsmError(f"b200msm_test_grp_op: {rc}: {lib().b200msm_last_error(None).decode()}")
    return out
