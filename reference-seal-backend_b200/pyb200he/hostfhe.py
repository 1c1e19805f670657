"""ctypes binding of the host-side FHE stand-in (reference-seal-backend_b200/hostfhe): parameter chain,
keygen, encode, encrypt, decrypt, decode -- the parts of SEAL that stay on the host in this design
(R/src/engine/seal_context.cpp:46-70).  No Evaluator arithmetic lives here."""
import ctypes as C
import os

import numpy as np

PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
u64p = C.POINTER(C.c_uint64)
u32p = C.POINTER(C.c_uint32)


def p64(a):
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(u64p)


_hfhe = None


def hostfhe():
    global _hfhe
    if _hfhe is None:
        h = C.CDLL(os.path.join(PKG, "hostfhe", "libhostfhe.so"))
        h.hfhe_create.restype = C.c_void_p
        h.hfhe_create.argtypes = [C.c_int, C.c_size_t, C.c_size_t, C.c_int, C.c_int, C.c_uint64]
        h.hfhe_destroy.argtypes = [C.c_void_p]
        for f, rt in (("hfhe_N", C.c_size_t), ("hfhe_K", C.c_size_t), ("hfhe_moduli", u64p), ("hfhe_psi", u64p),
                      ("hfhe_plain_modulus", C.c_uint64), ("hfhe_scale", C.c_double), ("hfhe_relin_key", u64p),
                      ("hfhe_galois_count", C.c_size_t), ("hfhe_kswitch_key_words", C.c_size_t)):
            getattr(h, f).restype = rt
            getattr(h, f).argtypes = [C.c_void_p]
        h.hfhe_galois_elt.restype = C.c_uint32
        h.hfhe_galois_elt.argtypes = [C.c_void_p, C.c_size_t]
        h.hfhe_galois_key.restype = u64p
        h.hfhe_galois_key.argtypes = [C.c_void_p, C.c_uint32]
        h.hfhe_ckks_encode.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.c_size_t, C.c_double, u64p]
        h.hfhe_ckks_decode.argtypes = [C.c_void_p, u64p, C.c_size_t, C.c_double, C.POINTER(C.c_double)]
        h.hfhe_bfv_encode.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.c_size_t, u64p]
        h.hfhe_bfv_decode.argtypes = [C.c_void_p, u64p, C.POINTER(C.c_int64)]
        h.hfhe_encrypt.argtypes = [C.c_void_p, u64p, u64p]
        h.hfhe_decrypt.argtypes = [C.c_void_p, u64p, C.c_size_t, C.c_size_t, u64p]
        _hfhe = h
    return _hfhe


BFV, CKKS = 1, 2


class Host:
    """hostfhe context: parameter chain + keys + encode/encrypt/decrypt/decode."""

    def __init__(self, scheme, N, depth, coeff_bits, sp_bits, seed=1234):
        self.h = hostfhe()
        self.c = self.h.hfhe_create(scheme, N, depth, coeff_bits, sp_bits, seed)
        self.scheme, self.N, self.K, self.Ltop = scheme, N, depth + 1, depth
        self.moduli = np.ctypeslib.as_array(self.h.hfhe_moduli(self.c), (self.K,)).copy()
        self.psi = np.ctypeslib.as_array(self.h.hfhe_psi(self.c), (self.K,)).copy()
        self.t = self.h.hfhe_plain_modulus(self.c)
        self.scale = self.h.hfhe_scale(self.c)
        self.kwords = self.h.hfhe_kswitch_key_words(self.c)

    def __del__(self):
        try:
            self.h.hfhe_destroy(self.c)
        except Exception:
            pass

    def relin_key(self):
        return np.ctypeslib.as_array(self.h.hfhe_relin_key(self.c), (self.kwords,)).copy()

    def galois_elts(self):
        return [self.h.hfhe_galois_elt(self.c, i) for i in range(self.h.hfhe_galois_count(self.c))]

    def galois_key(self, elt):
        p = self.h.hfhe_galois_key(self.c, elt)
        assert p
        return np.ctypeslib.as_array(p, (self.kwords,)).copy()

    def encode(self, vals, scale=None):
        if self.scheme == CKKS:
            v = np.ascontiguousarray(vals, dtype=np.float64)
            out = np.empty(self.Ltop * self.N, dtype=np.uint64)
            self.h.hfhe_ckks_encode(self.c, v.ctypes.data_as(C.POINTER(C.c_double)), len(v),
                                    self.scale if scale is None else scale, p64(out))
        else:
            v = np.ascontiguousarray(vals, dtype=np.int64)
            out = np.empty(self.N, dtype=np.uint64)
            self.h.hfhe_bfv_encode(self.c, v.ctypes.data_as(C.POINTER(C.c_int64)), len(v), p64(out))
        return out

    def encrypt(self, plain):
        ct = np.empty(2 * self.Ltop * self.N, dtype=np.uint64)
        self.h.hfhe_encrypt(self.c, p64(plain), p64(ct))
        return ct

    def decrypt(self, ct, size, L):
        ct = np.ascontiguousarray(ct, dtype=np.uint64)
        out = np.empty(L * self.N if self.scheme == CKKS else self.N, dtype=np.uint64)
        self.h.hfhe_decrypt(self.c, p64(ct), size, L, p64(out))
        return out

    def decode(self, plain, L=None, scale=None):
        if self.scheme == CKKS:
            out = np.empty(self.N // 2, dtype=np.float64)
            self.h.hfhe_ckks_decode(self.c, p64(plain), L, self.scale if scale is None else scale,
                                    out.ctypes.data_as(C.POINTER(C.c_double)))
        else:
            out = np.empty(self.N, dtype=np.int64)
            self.h.hfhe_bfv_decode(self.c, p64(plain), out.ctypes.data_as(C.POINTER(C.c_int64)))
        return out

    def enc_vec(self, vals):
        return self.encrypt(self.encode(vals))

    def dec_vec(self, ct, size, L, scale=None):
        return self.decode(self.decrypt(ct, size, L), L, scale)


