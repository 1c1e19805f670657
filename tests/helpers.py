"""ctypes bindings used by the tests: the CPU oracle (oracle/, checker only) and the
host-side FHE stand-in (hostfhe, keygen/encode/encrypt/decrypt -- not the hot path)."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "reference-seal-backend_b200")
if PKG not in sys.path:
    sys.path.insert(0, PKG)

u64p = C.POINTER(C.c_uint64)
u32p = C.POINTER(C.c_uint32)


def p64(a):
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(u64p)


_oracle = None


def oracle():
    global _oracle
    if _oracle is None:
        path = os.environ.get("B200HE_ORACLE_LIB") or os.path.join(ROOT, "oracle", "libhe_oracle.so")   # (a sanitizer build, tests/emu/run_asan.sh)
        if not os.path.exists(path):
            import subprocess
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
        o = C.CDLL(path)
        o.orc_plain_modulus_batching.restype = C.c_uint64
        o.orc_plain_modulus_batching.argtypes = [C.c_size_t, C.c_int]
        o.orc_coeff_modulus_create.argtypes = [C.c_size_t, C.POINTER(C.c_int), C.c_size_t, u64p]
        o.orc_minimal_primitive_root.argtypes = [C.c_uint64, C.c_uint64, u64p]
        o.orc_get_primes.argtypes = [C.c_uint64, C.c_int, C.c_size_t, u64p]
        o.orc_galois_elt_from_step.restype = C.c_uint32
        o.orc_galois_elt_from_step.argtypes = [C.c_int, C.c_size_t]
        o.orc_galois_elts_all.argtypes = [C.c_size_t, u32p]
        o.orc_galois_table_ntt.argtypes = [C.c_size_t, C.c_uint32, u32p]
        o.orc_ctx_create.restype = C.c_void_p
        o.orc_ctx_create.argtypes = [C.c_int, C.c_size_t, C.c_size_t, u64p, C.c_uint64]
        o.orc_ctx_destroy.argtypes = [C.c_void_p]
        o.orc_ctx_psi.restype = C.c_uint64
        o.orc_ctx_psi.argtypes = [C.c_void_p, C.c_size_t]
        o.orc_ctx_bsk_size.restype = C.c_size_t
        o.orc_ctx_bsk_size.argtypes = [C.c_void_p]
        o.orc_ctx_bsk.argtypes = [C.c_void_p, u64p]
        for f in ("orc_ntt_fwd", "orc_ntt_inv"):
            getattr(o, f).argtypes = [C.c_void_p, C.c_size_t, u64p]
        o.orc_ntt_fwd_direct.argtypes = [C.c_void_p, C.c_size_t, u64p, u64p]
        for f in ("orc_add", "orc_sub", "orc_multiply_plain", "orc_add_plain"):
            getattr(o, f).argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, u64p, u64p, u64p]
        o.orc_ckks_multiply.argtypes = [C.c_void_p, C.c_size_t, u64p, u64p, u64p]
        o.orc_bfv_multiply.argtypes = [C.c_void_p, u64p, u64p, u64p]
        o.orc_switch_key.argtypes = [C.c_void_p, C.c_size_t, u64p, u64p, u64p]
        o.orc_relinearize.argtypes = [C.c_void_p, C.c_size_t, u64p, u64p, u64p]
        o.orc_apply_galois.argtypes = [C.c_void_p, C.c_size_t, u64p, C.c_uint32, u64p]
        o.orc_rotate.argtypes = [C.c_void_p, C.c_size_t, u64p, C.c_int, u32p, C.POINTER(u64p), C.c_size_t]
        o.orc_rescale.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, u64p, u64p]
        o.orc_mod_drop.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, u64p, u64p]
        o.orc_batch_mul_relin_rescale.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, u64p, u64p, u64p, u64p, C.c_int]
        o.orc_accumulate.argtypes = [C.c_void_p, C.c_size_t, u64p, C.c_size_t, u32p, C.POINTER(u64p), C.c_size_t]
        o.orc_batch_dot.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, u64p, u64p, C.c_size_t, u64p, u32p,
                                    C.POINTER(u64p), C.c_size_t, u64p, C.c_int]
        o.orc_batch_ntt.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, u64p, C.c_int, C.c_int]
        o.orc_selftest_fastmod.restype = C.c_size_t
        o.orc_selftest_fastmod.argtypes = [C.c_uint64, C.c_size_t, C.c_uint64]
        kargs = [u64p, u32p, C.POINTER(u64p), C.c_size_t]   # relin, elts, galois keys, count
        o.orc_matmul_val.argtypes = [C.c_void_p] + [C.c_size_t] * 4 + [u64p, u64p] + kargs + [u64p, C.c_int]
        o.orc_matmul_row.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, u64p, u64p] + kargs + [u64p, C.c_int]
        o.orc_matmul_cba.argtypes = [C.c_void_p] + [C.c_size_t] * 4 + [u64p, u64p, u64p, u64p, C.c_int]
        o.orc_collapse.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, u64p, C.c_size_t, u64p, u64p] + kargs[1:] + [u64p, C.c_int]
        o.orc_horner.argtypes = [C.c_void_p, C.c_size_t, u64p, u64p, C.c_size_t, u64p, u64p, u64p]
        o.orc_logreg.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, u64p, u64p, u64p, u64p, u64p, u64p, C.c_size_t, u64p] + kargs + [u64p, C.c_int]
        _oracle = o
    return _oracle


from pyb200he.hostfhe import BFV, CKKS, Host, hostfhe  # noqa: E402,F401  (product-side host FHE binding)


class Oracle:
    """CPU oracle context (checker)."""

    def __init__(self, scheme, N, moduli, t=0):
        self.o = oracle()
        self.N, self.K, self.scheme = N, len(moduli), scheme
        self.moduli = np.ascontiguousarray(moduli, dtype=np.uint64)
        self.c = self.o.orc_ctx_create(scheme, N, self.K, p64(self.moduli), t)

    def __del__(self):
        try:
            self.o.orc_ctx_destroy(self.c)
        except Exception:
            pass

    def psi(self):
        return np.array([self.o.orc_ctx_psi(self.c, i) for i in range(self.K)], dtype=np.uint64)

    def ntt(self, limb, poly, inverse=False):
        x = np.array(poly, dtype=np.uint64, copy=True)
        (self.o.orc_ntt_inv if inverse else self.o.orc_ntt_fwd)(self.c, limb, p64(x))
        return x

    def add(self, L, size, a, b):
        out = np.empty_like(a)
        self.o.orc_add(self.c, L, size, p64(a), p64(b), p64(out))
        return out

    def ckks_multiply(self, L, a, b):
        out = np.empty(3 * L * self.N, dtype=np.uint64)
        self.o.orc_ckks_multiply(self.c, L, p64(a), p64(b), p64(out))
        return out

    def bfv_multiply(self, a, b):
        out = np.empty(3 * (self.K - 1) * self.N, dtype=np.uint64)
        self.o.orc_bfv_multiply(self.c, p64(a), p64(b), p64(out))
        return out

    def relinearize(self, L, ct3, key):
        out = np.empty(2 * L * self.N, dtype=np.uint64)
        self.o.orc_relinearize(self.c, L, p64(ct3), p64(key), p64(out))
        return out

    def switch_key(self, L, ct, target, key):
        out = np.array(ct, dtype=np.uint64, copy=True)
        self.o.orc_switch_key(self.c, L, p64(out), p64(target), p64(key))
        return out

    def apply_galois(self, L, ct, elt, key):
        out = np.array(ct, dtype=np.uint64, copy=True)
        self.o.orc_apply_galois(self.c, L, p64(out), elt, p64(key))
        return out

    @staticmethod
    def _keyargs(keys):
        elts = np.array(list(keys.keys()), dtype=np.uint32)
        arr = (u64p * len(keys))(*[p64(k) for k in keys.values()])
        return elts, arr

    def rotate(self, L, ct, step, keys):
        out = np.array(ct, dtype=np.uint64, copy=True)
        elts, arr = self._keyargs(keys)
        rc = self.o.orc_rotate(self.c, L, p64(out), step, elts.ctypes.data_as(u32p), arr, len(keys))
        assert rc == 0, rc
        return out

    def accumulate(self, L, ct, count, keys):
        out = np.array(ct, dtype=np.uint64, copy=True)
        elts, arr = self._keyargs(keys)
        rc = self.o.orc_accumulate(self.c, L, p64(out), count, elts.ctypes.data_as(u32p), arr, len(keys))
        assert rc == 0, rc
        return out

    def rescale(self, L, size, ct):
        out = np.empty(size * (L - 1) * self.N, dtype=np.uint64)
        self.o.orc_rescale(self.c, L, size, p64(ct), p64(out))
        return out

    def mod_drop(self, L, size, ct):
        out = np.empty(size * (L - 1) * self.N, dtype=np.uint64)
        self.o.orc_mod_drop(self.c, L, size, p64(ct), p64(out))
        return out

    def multiply_plain(self, L, size, ct, plain):
        out = np.empty(size * L * self.N, dtype=np.uint64)
        self.o.orc_multiply_plain(self.c, L, size, p64(ct), p64(plain), p64(out))
        return out

    def add_plain(self, L, size, ct, plain):
        out = np.empty(size * L * self.N, dtype=np.uint64)
        self.o.orc_add_plain(self.c, L, size, p64(ct), p64(plain), p64(out))
        return out

    def mul_relin_rescale(self, L, n, a, b, key, threads=0):
        out = np.empty(n * 2 * (L - 1) * self.N, dtype=np.uint64)
        self.o.orc_batch_mul_relin_rescale(self.c, L, n, p64(a), p64(b), p64(key), p64(out), threads)
        return out

    def batch_dot(self, L, n, a, b, count, relin, keys, threads=0):
        out = np.empty(n * 2 * L * self.N, dtype=np.uint64)
        elts, arr = self._keyargs(keys)
        rc = self.o.orc_batch_dot(self.c, L, n, p64(a), p64(b), count, p64(relin), elts.ctypes.data_as(u32p), arr,
                                  len(keys), p64(out), threads)
        assert rc == 0, rc
        return out


    # ---- workload bodies (oracle/he_oracle_workloads.cpp): arrays are flat uint64, ciphertexts [count][size][L][N]
    def matmul_val(self, L, r0, c0, c1, m0, m1t, relin, keys, threads=0):
        Lout = L - 1 if self.scheme == CKKS else L
        out = np.empty(r0 * c1 * 2 * Lout * self.N, dtype=np.uint64)
        elts, arr = self._keyargs(keys)
        rc = self.o.orc_matmul_val(self.c, L, r0, c0, c1, p64(m0), p64(m1t), p64(relin), elts.ctypes.data_as(u32p), arr, len(keys), p64(out), threads)
        assert rc == 0, rc
        return out

    def matmul_row(self, L, nA, dim2, spacers, A, B, relin, keys, threads=0):
        out = np.empty(nA * 2 * L * self.N, dtype=np.uint64)
        elts, arr = self._keyargs(keys)
        rc = self.o.orc_matmul_row(self.c, L, nA, dim2, spacers, p64(A), p64(B), p64(relin), elts.ctypes.data_as(u32p), arr, len(keys), p64(out), threads)
        assert rc == 0, rc
        return out

    def matmul_cba(self, L, r0, c0, c1, m0, m1, relin, threads=0):
        Lout = L - 1 if self.scheme == CKKS else L
        out = np.empty(r0 * c1 * 2 * Lout * self.N, dtype=np.uint64)
        rc = self.o.orc_matmul_cba(self.c, L, r0, c0, c1, p64(m0), p64(m1), p64(relin), p64(out), threads)
        assert rc == 0, rc
        return out

    def collapse(self, L, n, cts, first_index, masks, zero_ct, keys, threads=0):
        out = np.empty(2 * (L - 1) * self.N, dtype=np.uint64)
        elts, arr = self._keyargs(keys)
        rc = self.o.orc_collapse(self.c, L, n, p64(cts), first_index, p64(masks), p64(zero_ct) if zero_ct is not None else None,
                                 elts.ctypes.data_as(u32p), arr, len(keys), p64(out), threads)
        assert rc == 0, rc
        return out

    def horner(self, Lx, x, seed, coeffs, relin):
        """coeffs: [ncoef][Ltop][N] flat = the plaintexts of a_{d-1} .. a_0; returns (ciphertext, level)"""
        Ltop = self.K - 1
        ncoef = coeffs.size // (Ltop * self.N)
        out = np.empty(2 * Ltop * self.N, dtype=np.uint64)
        lv = self.o.orc_horner(self.c, Lx, p64(x), p64(seed), ncoef, p64(coeffs), p64(relin), p64(out))
        assert lv > 0, lv
        return out[: 2 * lv * self.N].copy(), lv

    def logreg(self, n_features, batch, W, b, X, masks, zero_ct, seed, coeffs, relin, keys, threads=0):
        Ltop = self.K - 1
        ncoef = coeffs.size // (Ltop * self.N)
        out = np.empty(2 * Ltop * self.N, dtype=np.uint64)
        elts, arr = self._keyargs(keys)
        lv = self.o.orc_logreg(self.c, n_features, batch, p64(W), p64(b), p64(X), p64(masks), p64(zero_ct), p64(seed), ncoef, p64(coeffs),
                               p64(relin), elts.ctypes.data_as(u32p), arr, len(keys), p64(out), threads)
        assert lv > 0, lv
        return out[: 2 * lv * self.N].copy(), lv


def rand_residues(rng, moduli, shape_prefix, N):
    """uniform residues: array [*shape_prefix, len(moduli), N] with limb l in [0, moduli[l])"""
    out = np.empty(tuple(shape_prefix) + (len(moduli), N), dtype=np.uint64)
    for l, q in enumerate(moduli):
        out[..., l, :] = rng.integers(0, int(q), size=tuple(shape_prefix) + (N,), dtype=np.uint64)
    return out
