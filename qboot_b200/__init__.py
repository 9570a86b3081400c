"""qboot_b200 -- B200-native engine for qboot's block-table hot path (ctypes binding of the C ABI).

The product path is the CUDA library `libqboot_b200.so` (include/qboot_b200.h).  There is no CPU fallback: importing
this package without the built library, or creating an Engine without a CUDA device, raises.
"""
from __future__ import annotations

import ctypes as C
import os
from fractions import Fraction

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("QB_LIB_PATH") or os.path.join(_HERE, "libqboot_b200.so")  # QB_LIB_PATH: development builds

# every symbol include/qboot_b200.h declares
SYMBOLS = (
    "qb_version", "qb_strerror", "qb_precision_supported", "qb_create", "qb_destroy", "qb_words",
    "qb_last_cuda_error", "qb_context_init", "qb_context_get", "qb_table_dim", "qb_block_dim", "qb_gblock_batch",
    "qb_gblock_plan", "qb_gblock_run", "qb_gblock_fetch", "qb_sync", "qb_gblock_device_tables", "qb_launch_count",
    "qb_stream", "qb_fblock_batch", "qb_op_batch", "qb_set_timing", "qb_kernel_times", "qb_imad_peak",
    "qb_vtod_batch", "qb_fblock_tables", "qb_pinned_alloc", "qb_pinned_free",
    "qb_pmp_assemble", "qb_pmp_block_add", "qb_pmp_block_fetch", "qb_pmp_pack", "qb_pmp_packed_fetch", "qb_pmp_text",
    "qb_pmp_text_declined", "qb_pmp_text_patch", "qb_pmp_text_fetch", "qb_pmp_info", "qb_dec_format", "qb_poly_eval_batch",
    "qb_set_ziv_slack", "qb_ziv_counters",
    "qb_gblock_set_sink", "qb_device_alloc", "qb_device_free", "qb_ipc_export", "qb_ipc_open", "qb_ipc_close",
)
# the C++ facade (namespace qboot, include/qboot/*.hpp) built next to the C-ABI library; Python reaches it through the
# model-level C entry points of qboot_b200/host/models_capi.cpp
FACADE_PATH = os.path.join(_HERE, "libqboot.so")


class QbError(RuntimeError):
    pass


_lib = None


def load_library():
    """Load the CUDA engine; fails loudly when it has not been built (python -m qboot_b200.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise QbError(f"{LIB_PATH} is missing: build it with `python -m qboot_b200.build` (there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    lib.qb_version.restype = C.c_char_p
    lib.qb_strerror.restype = C.c_char_p
    lib.qb_strerror.argtypes = [C.c_int]
    lib.qb_last_cuda_error.restype = C.c_char_p
    lib.qb_last_cuda_error.argtypes = [C.c_void_p]
    lib.qb_create.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_void_p)]
    lib.qb_destroy.argtypes = [C.c_void_p]
    lib.qb_destroy.restype = None
    lib.qb_words.argtypes = [C.c_void_p]
    lib.qb_context_init.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_int64, C.c_uint32]
    lib.qb_context_get.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.qb_table_dim.argtypes = [C.c_void_p]
    lib.qb_block_dim.argtypes = [C.c_void_p, C.c_int]
    lib.qb_gblock_batch.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 5
    lib.qb_gblock_plan.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 4
    lib.qb_gblock_run.argtypes = [C.c_void_p]
    lib.qb_gblock_fetch.argtypes = [C.c_void_p, C.c_void_p]
    lib.qb_sync.argtypes = [C.c_void_p]
    lib.qb_gblock_device_tables.argtypes = [C.c_void_p]
    lib.qb_gblock_device_tables.restype = C.c_void_p
    lib.qb_launch_count.argtypes = [C.c_void_p]
    lib.qb_launch_count.restype = C.c_int64
    lib.qb_stream.argtypes = [C.c_void_p]
    lib.qb_stream.restype = C.c_uint64
    lib.qb_fblock_batch.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p]
    lib.qb_op_batch.argtypes = [C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 6
    lib.qb_set_timing.argtypes = [C.c_void_p, C.c_int]
    lib.qb_kernel_times.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
    lib.qb_imad_peak.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double)]
    lib.qb_vtod_batch.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    lib.qb_fblock_tables.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 4
    lib.qb_pinned_alloc.argtypes = [C.c_size_t]
    lib.qb_pinned_alloc.restype = C.c_void_p
    lib.qb_pinned_free.argtypes = [C.c_void_p]
    lib.qb_pinned_free.restype = None
    lib.qb_pmp_assemble.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                                    C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
    lib.qb_pmp_block_add.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    lib.qb_pmp_block_fetch.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    lib.qb_pmp_pack.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.qb_pmp_packed_fetch.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    lib.qb_pmp_text.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    lib.qb_pmp_text_declined.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    lib.qb_pmp_text_patch.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    lib.qb_pmp_text_fetch.argtypes = [C.c_void_p, C.c_void_p]
    lib.qb_pmp_info.argtypes = [C.c_void_p, C.c_void_p]
    lib.qb_dec_format.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    lib.qb_poly_eval_batch.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                       C.c_void_p, C.c_void_p]
    lib.qb_gblock_set_sink.argtypes = [C.c_void_p, C.c_void_p]
    lib.qb_device_alloc.argtypes = [C.c_void_p, C.c_size_t]
    lib.qb_device_alloc.restype = C.c_void_p
    lib.qb_device_free.argtypes = [C.c_void_p, C.c_void_p]
    lib.qb_device_free.restype = None
    lib.qb_ipc_export.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.qb_ipc_open.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]
    lib.qb_ipc_close.argtypes = [C.c_void_p, C.c_void_p]
    lib.qb_set_ziv_slack.argtypes = [C.c_void_p, C.c_int]
    lib.qb_ziv_counters.argtypes = [C.c_void_p, C.c_void_p]
    _lib = lib
    return lib


_facade = None


def load_facade(path: str | None = None):
    """Load the C++ facade library (libqboot.so); it links against the CUDA engine, so the same no-CPU-path rule holds."""
    global _facade
    if path is None and _facade is not None:
        return _facade
    p = path or FACADE_PATH
    if not os.path.exists(p):
        raise QbError(f"{p} is missing: build it with `python -m qboot_b200.build`")
    lib = C.CDLL(p)
    lib.qbf_last_error.restype = C.c_char_p
    lib.qbf_build_pmp.argtypes = [C.c_int, C.c_uint32, C.c_uint32, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_char_p,
                                  C.c_char_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_char_p, C.c_void_p, C.c_void_p]
    if path is None:
        _facade = lib
    return lib


def build_pmp(prec, n_max, lam, dim="3", model=0, mode=0, finite=True, ds="0.5181489", de="1.412625", de1="3.83", numax=8,
              maxspin=24, threads=1, outdir=None, lib=None):
    """convert -> create_input -> write of one of the reference's model problems through the C++ facade
    (qboot::Context / BootstrapEquation / PolynomialProgram / SDPBInput on the GPU engine); see host/models_capi.cpp.
    Returns the phase times and sizes like oracle/ref.py::build_pmp does for the reference."""
    lib = lib or load_facade()
    times = (C.c_double * 3)()
    counts = (C.c_int64 * 3)()
    rc = lib.qbf_build_pmp(prec, n_max, lam, str(dim).encode(), model, mode, int(finite), ds.encode(), de.encode(), de1.encode(),
                           numax, maxspin, threads, (outdir or "").encode(), times, counts)
    if rc != 0:
        raise QbError("qbf_build_pmp: " + lib.qbf_last_error().decode())
    return dict(convert=times[0], create_input=times[1], write=times[2], constraints=counts[0], N=counts[1], free=counts[2])


def nlimbs(prec: int) -> int:
    return (prec + 31) // 32


def words(prec: int) -> int:
    return nlimbs(prec) + 2


# ---- numbers: exact conversions between Python values and the record format ----------------------------------------------
def from_fraction(x, prec: int) -> np.ndarray:
    """Round a rational to `prec` bits, nearest-even (what mpfr_set_q / mpfr_set_str do), as a record."""
    x = Fraction(x)
    L = nlimbs(prec)
    w = np.zeros(L + 2, dtype=np.uint32)
    if x == 0:
        return w
    sign = x < 0
    x = abs(x)
    n, d = x.numerator, x.denominator
    e = n.bit_length() - d.bit_length()  # 2^(e-1) < x < 2^(e+1)
    # scale so that the quotient has prec + 2 bits or more
    sh = prec + 2 - e
    q, r = divmod(n << sh, d) if sh >= 0 else divmod(n, d << (-sh))
    nb = q.bit_length()
    extra = nb - prec
    m, rem = q >> extra, q & ((1 << extra) - 1)
    half = 1 << (extra - 1)
    rnd = (rem >> (extra - 1)) & 1
    sticky = (rem & (half - 1)) != 0 or r != 0
    if rnd and (sticky or (m & 1)):
        m += 1
    if m.bit_length() > prec:
        m >>= 1
        extra += 1
    exp2 = extra - sh  # value = m * 2^exp2
    mm = m << (32 * L - prec)
    for i in range(L):
        w[2 + i] = (mm >> (32 * i)) & 0xFFFFFFFF
    w[1] = np.uint32((exp2 + prec) & 0xFFFFFFFF)
    w[0] = 1 | (0x80000000 if sign else 0)
    return w


def to_fraction(w, prec: int) -> Fraction:
    w = np.asarray(w, dtype=np.uint32)
    kind = int(w[0]) & 3
    if kind == 0:
        return Fraction(0)
    if kind != 1:
        raise ValueError("inf/nan record")
    L = nlimbs(prec)
    m = 0
    for i in range(L - 1, -1, -1):
        m = (m << 32) | int(w[2 + i])
    e = int(np.int32(w[1])) - 32 * L
    v = Fraction(m) * Fraction(2) ** e
    return -v if (int(w[0]) >> 31) else v


def from_str(s: str, prec: int) -> np.ndarray:
    """Decimal string ("1.412625", "3/7", "2.5e-3") -> correctly rounded record."""
    return from_fraction(Fraction(s), prec)


def zeros(n: int, prec: int) -> np.ndarray:
    return np.zeros((n, words(prec)), dtype=np.uint32)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Engine:
    """One handle = one GPU.  Mirrors what the reference's `Context` owns for this path (include/qboot/context.hpp)."""

    def __init__(self, prec: int, device: int = 0):
        self.lib = load_library()
        self.prec = int(prec)
        self.h = C.c_void_p()
        rc = self.lib.qb_create(int(device), self.prec, C.byref(self.h))
        if rc != 0:
            raise QbError(f"qb_create(device={device}, prec={prec}): {self.lib.qb_strerror(rc).decode()}")
        self.W = self.lib.qb_words(self.h)
        self.lam = None
        self.n_max = None

    def close(self):
        if getattr(self, "h", None):
            self.lib.qb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            extra = self.lib.qb_last_cuda_error(self.h).decode() if rc == -5 else ""
            raise QbError(f"{what}: {self.lib.qb_strerror(rc).decode()} {extra}")

    def context(self, n_max: int, lam: int, dim="3"):
        """Context(n_Max, lambda, dim) (src/context.cpp:47-61)."""
        eps = Fraction(dim) / 2 - 1
        self._check(self.lib.qb_context_init(self.h, n_max, lam, eps.numerator, eps.denominator), "qb_context_init")
        self.lam, self.n_max = lam, n_max
        self.dim_M = self.lib.qb_table_dim(self.h)
        return self

    def context_consts(self):
        rho = np.zeros(self.W, dtype=np.uint32)
        mat = np.zeros((self.lam + 1, self.lam + 1, self.W), dtype=np.uint32)
        self._check(self.lib.qb_context_get(self.h, _ptr(rho), _ptr(mat)), "qb_context_get")
        return rho, mat

    def _inputs(self, spin, delta, S, P):
        delta = np.ascontiguousarray(delta, dtype=np.uint32).reshape(-1, self.W)
        n = len(delta)
        S = np.ascontiguousarray(S, dtype=np.uint32).reshape(n, self.W)
        P = np.ascontiguousarray(P, dtype=np.uint32).reshape(n, self.W)
        spin = np.ascontiguousarray(spin, dtype=np.uint32).reshape(n)
        return n, spin, delta, S, P

    def gblock(self, spin, delta, S, P) -> np.ndarray:
        """gBlock(op, S, P, ctx) for a batch (src/hor_recursion.cpp:57-60): host in, host out, (n, dim_M, W)."""
        n, spin, delta, S, P = self._inputs(spin, delta, S, P)
        out = np.zeros((n, self.dim_M, self.W), dtype=np.uint32)
        self._check(self.lib.qb_gblock_batch(self.h, n, _ptr(spin), _ptr(delta), _ptr(S), _ptr(P), _ptr(out)),
                    "qb_gblock_batch")
        return out

    def plan(self, spin, delta, S, P) -> int:
        n, spin, delta, S, P = self._inputs(spin, delta, S, P)
        self._check(self.lib.qb_gblock_plan(self.h, n, _ptr(spin), _ptr(delta), _ptr(S), _ptr(P)), "qb_gblock_plan")
        self._n = n
        return n

    def run(self):
        self._check(self.lib.qb_gblock_run(self.h), "qb_gblock_run")

    def sync(self):
        self._check(self.lib.qb_sync(self.h), "qb_sync")

    def fetch(self) -> np.ndarray:
        out = np.zeros((self._n, self.dim_M, self.W), dtype=np.uint32)
        self._check(self.lib.qb_gblock_fetch(self.h, _ptr(out)), "qb_gblock_fetch")
        return out

    def fblock(self, table_idx, d, d_idx, sym):
        """F/H blocks over the resident tables: list of (dim_sym, W) arrays, one per term (context.hpp:123-139)."""
        table_idx = np.ascontiguousarray(table_idx, dtype=np.int32)
        d_idx = np.ascontiguousarray(d_idx, dtype=np.int32)
        sym = np.ascontiguousarray(sym, dtype=np.int32)
        d = np.ascontiguousarray(d, dtype=np.uint32).reshape(-1, self.W)
        m = len(table_idx)
        dims = [self.lib.qb_block_dim(self.h, int(s)) for s in sym]
        out = np.zeros((int(sum(dims)), self.W), dtype=np.uint32)
        self._check(self.lib.qb_fblock_batch(self.h, m, _ptr(table_idx), len(d), _ptr(d), _ptr(d_idx), _ptr(sym),
                                             _ptr(out)), "qb_fblock_batch")
        res, o = [], 0
        for k in dims:
            res.append(out[o:o + k])
            o += k
        return res

    def op(self, code: int, a=None, b=None, c=None, ia=None, ib=None) -> np.ndarray:
        """Diagnostic: one arithmetic primitive elementwise on the device (qb_op_batch)."""
        arrs = [None if x is None else np.ascontiguousarray(x, dtype=np.uint32).reshape(-1, self.W) for x in (a, b, c)]
        ints = [None if x is None else np.ascontiguousarray(x, dtype=np.int64).reshape(-1) for x in (ia, ib)]
        n = next(len(x) for x in arrs + ints if x is not None)
        out = np.zeros((n, self.W), dtype=np.uint32)
        self._check(self.lib.qb_op_batch(self.h, int(code), n, _ptr(arrs[0]), _ptr(arrs[1]), _ptr(arrs[2]),
                                         _ptr(ints[0]), _ptr(ints[1]), _ptr(out)), "qb_op_batch")
        return out

    def set_timing(self, on: bool = True):
        self._check(self.lib.qb_set_timing(self.h, int(on)), "qb_set_timing")

    def kernel_times(self):
        """device ms of the last run: (rec_coeffs, series, deriv, finish)"""
        ms = (C.c_float * 4)()
        self._check(self.lib.qb_kernel_times(self.h, ms), "qb_kernel_times")
        return [float(x) for x in ms]

    def imad_peak(self, iters: int = 4096) -> float:
        """measured integer multiply-add rate of this GPU in limb-MACs per second"""
        v = C.c_double()
        self._check(self.lib.qb_imad_peak(self.h, iters, C.byref(v)), "qb_imad_peak")
        return float(v.value)

    def set_ziv_slack(self, bits: int = 0):
        """width of the powers' in-doubt window in bits of the extended result's last place (0: default); test knob"""
        self._check(self.lib.qb_set_ziv_slack(self.h, bits), "qb_set_ziv_slack")
        return self

    def ziv_counters(self):
        """(powers recomputed at the second Ziv level, powers still in doubt there) since the handle was created"""
        out = (C.c_uint64 * 2)()
        self._check(self.lib.qb_ziv_counters(self.h, out), "qb_ziv_counters")
        return int(out[0]), int(out[1])

    @property
    def launches(self) -> int:
        return int(self.lib.qb_launch_count(self.h))

    @property
    def stream(self) -> int:
        return int(self.lib.qb_stream(self.h))
