"""ORACLE (test infrastructure): ctypes binding of oracle/_build/liboracle.so, the C restatement of
halo2_proofs 0.2.0 arithmetic (best_multiexp / best_fft / ... ; see oracle/c/*.h for citations).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this."""
import ctypes, os, subprocess
import numpy as np
from . import pasta

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def build():
    subprocess.run(["make", "-C", _HERE], check=True, stdout=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.orc_init()
    return _lib


class _TimedLib:
    """Proxy over the CDLL that adds up the wall time spent inside the C library (bench.py's cpu_baseline reports it as
    `native_share`: how much of the restated prover's time is C arithmetic rather than the Python protocol driver)."""

    def __init__(self, inner):
        self.inner, self.seconds, self.calls, self._cache = inner, 0.0, 0, {}

    def __getattr__(self, name):
        w = self._cache.get(name)
        if w is None:
            import time
            f = getattr(self.inner, name)

            def w(*a, _f=f, _t=time.perf_counter):
                t0 = _t()
                r = _f(*a)
                self.seconds += _t() - t0
                self.calls += 1
                return r
            self._cache[name] = w
        return w


def call_timing(on):
    """Switch the per-call timer on / off; returns the proxy (on) or the accumulated (seconds, calls) (off)."""
    global _lib
    l = lib()
    if on:
        if not isinstance(l, _TimedLib):
            _lib = _TimedLib(l)
        _lib.seconds, _lib.calls = 0.0, 0
        return _lib
    if isinstance(l, _TimedLib):
        _lib = l.inner
        return l.seconds, l.calls
    return 0.0, 0


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


FIELDS = {0: pasta.FP, 1: pasta.FQ}
# curve id -> (curve, scalar field id, base field id)
CURVES = {0: (pasta.VESTA, 0, 1), 1: (pasta.PALLAS, 1, 0)}


def ints_to_raw(vals):
    """list of canonical ints -> (n,4) uint64 little-endian limbs (NOT Montgomery)."""
    buf = b"".join(int(v).to_bytes(32, "little") for v in vals)
    return np.frombuffer(buf, dtype=np.uint64).reshape(-1, 4).copy()


def raw_to_ints(arr):
    b = np.ascontiguousarray(arr).tobytes()
    return [int.from_bytes(b[i:i + 32], "little") for i in range(0, len(b), 32)]


def to_mont(f, vals):
    """canonical ints -> Montgomery limbs via the C oracle (fast path for large arrays)."""
    raw = ints_to_raw(vals)
    out = np.empty_like(raw)
    lib().orc_field_from_repr(f, _p(raw), _p(out), ctypes.c_size_t(len(raw)))
    return out


def from_mont(f, arr):
    arr = np.ascontiguousarray(arr, dtype=np.uint64).reshape(-1, 4)
    out = np.empty_like(arr)
    lib().orc_field_to_repr(f, _p(arr), _p(out), ctypes.c_size_t(len(arr)))
    return raw_to_ints(out)


def from_u512(f, wide):
    """(n,8) uint64 -> (n,4) Montgomery, pasta `from_u512`."""
    wide = np.ascontiguousarray(wide, dtype=np.uint64).reshape(-1, 8)
    out = np.empty((len(wide), 4), dtype=np.uint64)
    lib().orc_field_from_u512(f, _p(wide), _p(out), ctypes.c_size_t(len(wide)))
    return out


def field_mul(f, a, b):
    out = np.empty_like(a)
    lib().orc_field_mul(f, _p(a), _p(b), _p(out), ctypes.c_size_t(len(a)))
    return out


def points_to_mont(curve, pts):
    """list of affine (x,y)|None -> (n,8) uint64 Montgomery x||y (identity zeros)."""
    _, _, bf = CURVES[curve]
    flat = []
    for pt in pts:
        flat += [0, 0] if pt is None else [pt[0], pt[1]]
    return to_mont(bf, flat).reshape(-1, 8)


def points_from_mont(curve, arr):
    _, _, bf = CURVES[curve]
    v = from_mont(bf, np.ascontiguousarray(arr).reshape(-1, 4))
    out = []
    for i in range(0, len(v), 2):
        out.append(None if v[i] == 0 and v[i + 1] == 0 else (v[i], v[i + 1]))
    return out


def best_multiexp(curve, scalars_mont, bases_mont):
    """-> (12,) uint64 Jacobian Montgomery."""
    n = len(scalars_mont)
    assert len(bases_mont) == n
    out = np.zeros(12, dtype=np.uint64)
    lib().orc_best_multiexp(curve, _p(scalars_mont), _p(bases_mont), ctypes.c_size_t(n), _p(out))
    return out


def to_affine(curve, jac):
    jac = np.ascontiguousarray(jac, dtype=np.uint64).reshape(-1, 12)
    out = np.zeros((len(jac), 8), dtype=np.uint64)
    lib().orc_to_affine(curve, _p(jac), _p(out), ctypes.c_size_t(len(jac)))
    return out


def best_fft(f, a_mont, omega_mont, log_n):
    a = np.ascontiguousarray(a_mont, dtype=np.uint64).copy()
    assert a.shape == (1 << log_n, 4)
    lib().orc_best_fft(f, _p(a), _p(np.ascontiguousarray(omega_mont)), ctypes.c_uint(log_n))
    return a


def set_threads(n):
    lib().orc_set_threads(int(n))


def get_threads():
    return lib().orc_get_threads()
