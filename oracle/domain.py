"""ORACLE (test infrastructure): restatement of halo2_proofs 0.2.0 `poly::EvaluationDomain`
(U: halo2_proofs/src/poly/domain.rs, crate pinned at /root/reference/Cargo.lock:382-393, not vendored;
SURVEY §8 a5, App. B) on top of the C restatement of `best_fft` (oracle/c/fft_tmpl.h).

Arrays are numpy (n,4) uint64 Montgomery limbs (pasta's in-memory form)."""
import numpy as np
from . import pasta, c_oracle as co


class EvaluationDomain:
    """`EvaluationDomain::new(j, k)`: j = constraint-system degree, n = 2^k."""

    def __init__(self, field_id, j, k):
        self.f = field_id
        F = co.FIELDS[field_id]
        self.F = F
        p = F.p
        self.k = k
        self.n = 1 << k
        self.quotient_poly_degree = j - 1
        ek = k
        while (1 << ek) < self.n * self.quotient_poly_degree:
            ek += 1
        self.extended_k = ek
        w = F.root_of_unity
        for _ in range(ek, pasta.S):
            w = w * w % p
        self.extended_omega = w
        self.extended_omega_inv = pow(w, -1, p)
        for _ in range(k, ek):
            w = w * w % p
        self.omega = w
        self.omega_inv = pow(w, -1, p)
        self.g_coset = F.zeta
        self.g_coset_inv = F.zeta * F.zeta % p
        self.ifft_divisor = pow(self.n, -1, p)
        self.extended_ifft_divisor = pow(1 << ek, -1, p)
        # t_evaluations: (zeta^n * (w_ext^n)^i - 1)^-1 for i < 2^(ek-k)
        orig = pow(F.zeta, self.n, p)
        step = pow(self.extended_omega, self.n, p)
        cur, t = orig, []
        while True:
            t.append(cur)
            cur = cur * step % p
            if cur == orig:
                break
        assert len(t) == 1 << (ek - k)
        self.t_evaluations = [pow((x - 1) % p, -1, p) for x in t]

    # -- helpers on Montgomery arrays --
    def _const(self, v, n):
        return np.repeat(co.to_mont(self.f, [v]), n, axis=0)

    def _ifft(self, a, omega_inv, log_n, divisor):
        a = co.best_fft(self.f, a, co.to_mont(self.f, [omega_inv]), log_n)
        return co.field_mul(self.f, a, self._const(divisor, len(a)))

    def lagrange_to_coeff(self, a):
        assert len(a) == self.n
        return self._ifft(a, self.omega_inv, self.k, self.ifft_divisor)

    def _distribute_powers_zeta(self, a, into_coset):
        powers = [self.g_coset, self.g_coset_inv] if into_coset else [self.g_coset_inv, self.g_coset]
        mult = co.to_mont(self.f, [1, powers[0], powers[1]])
        idx = np.arange(len(a)) % 3
        return co.field_mul(self.f, np.ascontiguousarray(a), np.ascontiguousarray(mult[idx]))

    def coeff_to_extended(self, a):
        assert len(a) == self.n
        a = self._distribute_powers_zeta(a, True)
        ext = np.zeros((1 << self.extended_k, 4), dtype=np.uint64)
        ext[: self.n] = a
        return co.best_fft(self.f, ext, co.to_mont(self.f, [self.extended_omega]), self.extended_k)

    def extended_to_coeff(self, a):
        assert len(a) == 1 << self.extended_k
        a = self._ifft(a, self.extended_omega_inv, self.extended_k, self.extended_ifft_divisor)
        a = self._distribute_powers_zeta(a, False)
        return a[: self.n * self.quotient_poly_degree]

    def divide_by_vanishing_poly(self, a):
        assert len(a) == 1 << self.extended_k
        t = co.to_mont(self.f, self.t_evaluations)
        idx = np.arange(len(a)) % len(self.t_evaluations)
        return co.field_mul(self.f, np.ascontiguousarray(a), np.ascontiguousarray(t[idx]))

    def rotate_omega(self, value, rotation):
        p = self.F.p
        if rotation >= 0:
            return value * pow(self.omega, rotation, p) % p
        return value * pow(self.omega_inv, -rotation, p) % p
