"""ORACLE (test infrastructure): thin numpy views over the C restatement's vector helpers
(oracle/c/oracle_api.c).  Arrays are (n,4) uint64 Montgomery limbs; scalars are canonical Python ints."""
import ctypes
import numpy as np
from . import c_oracle as co


class Vec:
    """Vector arithmetic over one Pasta field (0 = Fp, 1 = Fq)."""

    def __init__(self, f):
        self.f = f
        self.F = co.FIELDS[f]
        self.p = self.F.p

    @property
    def lib(self):
        return co.lib()          # looked up per call: c_oracle.call_timing() may have swapped in its timing proxy

    # -- conversions --
    def m(self, x):
        """canonical int -> (4,) Montgomery limbs"""
        return np.frombuffer(self.F.to_mont_bytes(x % self.p), dtype=np.uint64).copy()

    def arr(self, ints):
        return co.to_mont(self.f, [v % self.p for v in ints]) if len(ints) else np.zeros((0, 4), np.uint64)

    def ints(self, arr):
        return co.from_mont(self.f, arr)

    def int1(self, limbs):
        return self.F.from_mont_bytes(np.ascontiguousarray(limbs).tobytes())

    def zeros(self, n):
        return np.zeros((n, 4), dtype=np.uint64)

    def const(self, x, n):
        return np.repeat(self.m(x)[None, :], n, axis=0)

    # -- element-wise --
    def _op(self, op, a, b):
        a = np.ascontiguousarray(a)
        b = np.ascontiguousarray(b)
        r = np.empty_like(a)
        self.lib.orc_vec_op(self.f, op, co._p(a), co._p(b), co._p(r), ctypes.c_size_t(len(a)))
        return r

    def mul(self, a, b): return self._op(0, a, b)
    def add(self, a, b): return self._op(1, a, b)
    def sub(self, a, b): return self._op(2, a, b)
    def scale(self, a, s): return self._op(3, a, self.m(s))
    def neg(self, a): return self._op(4, a, a)
    def add_const(self, a, s): return self._op(5, a, self.m(s))

    def batch_invert(self, a):
        a = np.ascontiguousarray(a).copy()
        self.lib.orc_batch_invert(self.f, co._p(a), ctypes.c_size_t(len(a)))
        return a

    def eval_polynomial(self, poly, x):
        poly = np.ascontiguousarray(poly)
        out = np.zeros(4, dtype=np.uint64)
        self.lib.orc_eval_polynomial(self.f, co._p(poly), ctypes.c_size_t(len(poly)), co._p(self.m(x)), co._p(out))
        return self.int1(out)

    def inner_product(self, a, b):
        a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
        out = np.zeros(4, dtype=np.uint64)
        self.lib.orc_inner_product(self.f, co._p(a), co._p(b), ctypes.c_size_t(len(a)), co._p(out))
        return self.int1(out)

    def kate_division(self, a, b):
        a = np.ascontiguousarray(a)
        out = np.zeros((len(a) - 1, 4), dtype=np.uint64)
        self.lib.orc_kate_division(self.f, co._p(a), ctypes.c_size_t(len(a)), co._p(self.m(b)), co._p(out))
        return out

    def running_product(self, z0, frac):
        """z[0] = z0, z[i] = z[i-1] * frac[i-1]; len(z) == len(frac)"""
        frac = np.ascontiguousarray(frac)
        z = np.zeros_like(frac)
        self.lib.orc_running_product(self.f, co._p(self.m(z0)), co._p(frac), co._p(z), ctypes.c_size_t(len(frac)))
        return z

    def powers(self, first, base, n):
        out = np.zeros((n, 4), dtype=np.uint64)
        self.lib.orc_powers(self.f, co._p(self.m(first)), co._p(self.m(base)), co._p(out), ctypes.c_size_t(n))
        return out

    def sort_keys(self, a):
        """bytes objects whose lexicographic order == numeric order of the canonical values (pasta Ord)."""
        a = np.ascontiguousarray(a)
        out = np.zeros((len(a), 32), dtype=np.uint8)
        self.lib.orc_canonical_be(self.f, co._p(a), co._p(out), ctypes.c_size_t(len(a)))
        return out
