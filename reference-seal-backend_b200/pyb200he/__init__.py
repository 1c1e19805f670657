"""ctypes binding of libb200he.so (include/b200he.h) -- the C-ABI drop-in boundary of the
ciphertext-evaluation hot path.  Used by the tests, bench.py and __graft_entry__.py; the C++
backend (reference-seal-backend_b200/backend) links the same library directly.

There is no CPU path: importing works anywhere (so the symbol/export checks can run without a
GPU), but Context() raises if the CUDA library is missing or no CUDA device is present.
"""
import ctypes as C
import os

import numpy as np

BFV, CKKS = 1, 2
_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "lib", "libb200he.so")

u64p = C.POINTER(C.c_uint64)
u32p = C.POINTER(C.c_uint32)

KERNEL_CLASSES = 10

# name -> (restype, argtypes); the complete export list of include/b200he.h
SIGNATURES = {
    "b200he_last_error": (C.c_char_p, []),
    "b200he_version": (C.c_char_p, []),
    "b200he_ctx_create": (C.c_int, [C.c_int, C.c_uint32, C.c_uint32, u64p, u64p, C.c_uint64, C.c_int, C.POINTER(C.c_void_p)]),
    "b200he_ctx_destroy": (None, [C.c_void_p]),
    "b200he_ctx_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "b200he_ctx_sync": (C.c_int, [C.c_void_p]),
    "b200he_ctx_set_workspace": (C.c_int, [C.c_void_p, C.c_uint64]),
    "b200he_set_relin_key": (C.c_int, [C.c_void_p, u64p]),
    "b200he_set_galois_key": (C.c_int, [C.c_void_p, C.c_uint32, u64p]),
    "b200he_has_galois_key": (C.c_int, [C.c_void_p, C.c_uint32]),
    "b200he_batch_create": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "b200he_batch_destroy": (None, [C.c_void_p]),
    "b200he_batch_resize": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_double]),
    "b200he_batch_upload": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p]),
    "b200he_batch_download": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p]),
    "b200he_batch_download_async": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p]),
    "b200he_batch_upload_scattered": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p]),
    "b200he_batch_download_scattered": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p]),
    "b200he_batch_copy_from": (C.c_int, [C.c_void_p, C.c_void_p]),
    "b200he_batch_count": (C.c_uint64, [C.c_void_p]),
    "b200he_batch_size": (C.c_int, [C.c_void_p]),
    "b200he_batch_level": (C.c_int, [C.c_void_p]),
    "b200he_batch_ntt_form": (C.c_int, [C.c_void_p]),
    "b200he_batch_scale": (C.c_double, [C.c_void_p]),
    "b200he_batch_set_scale": (C.c_int, [C.c_void_p, C.c_double]),
    "b200he_batch_device_ptr": (C.c_void_p, [C.c_void_p]),
    "b200he_add": (C.c_int, [C.c_void_p, C.c_void_p, u32p, C.c_void_p, u32p, C.c_uint64, C.c_void_p]),
    "b200he_sub": (C.c_int, [C.c_void_p, C.c_void_p, u32p, C.c_void_p, u32p, C.c_uint64, C.c_void_p]),
    "b200he_multiply": (C.c_int, [C.c_void_p, C.c_void_p, u32p, C.c_void_p, u32p, C.c_uint64, C.c_void_p]),
    "b200he_matmul_accumulate": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p]),
    "b200he_relinearize": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "b200he_rotate": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "b200he_rotate_columns": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "b200he_rotate_each": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_int32), C.c_void_p]),
    "b200he_sum": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "b200he_apply_galois": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]),
    "b200he_rescale_to_next": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "b200he_relinearize_rescale": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "b200he_mod_drop": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "b200he_multiply_plain": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, u32p, C.c_void_p]),
    "b200he_add_plain": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, u32p, C.c_void_p]),
    "b200he_accumulate": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64]),
    "b200he_gather": (C.c_int, [C.c_void_p, C.c_void_p, u32p, C.c_uint64, C.c_void_p]),
    "b200he_ntt_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "b200he_ntt_inverse": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "b200he_launch_count": (C.c_uint64, [C.c_void_p]),
    "b200he_profile_begin": (C.c_int, [C.c_void_p]),
    "b200he_profile_end": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), u64p]),
    "b200he_profile_work": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "b200he_kernel_name": (C.c_char_p, [C.c_int]),
}


class B200HEError(RuntimeError):
    pass


def declare(lib):
    """attach the header's signatures to a loaded CDLL; raises AttributeError on a missing export"""
    for name, (res, args) in SIGNATURES.items():
        f = getattr(lib, name)
        f.restype = res
        f.argtypes = args
    return lib


_lib = None


def load_library(path=None):
    """load the CUDA library (never the host-C++ emulation build used by the CPU-side tests)"""
    global _lib
    if _lib is None or path:
        p = path or LIB_PATH
        if not os.path.exists(p):
            raise B200HEError(f"{p} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(there is no CPU fallback)")
        lib = declare(C.CDLL(p))
        if b"EMU" in lib.b200he_version():
            raise B200HEError("refusing to load the host emulation build as the product library")
        if path:
            return lib
        _lib = lib
    return _lib


def _idx(a):
    if a is None:
        return None, None
    arr = np.ascontiguousarray(a, dtype=np.uint32)
    return arr, arr.ctypes.data_as(u32p)


class Batch:
    """device-resident vector of ciphertexts / plaintexts: uint64[count][size][L][N] in HBM"""

    def __init__(self, ctx):
        self.ctx = ctx
        self.lib = ctx.lib
        h = C.c_void_p()
        ctx._ck(self.lib.b200he_batch_create(ctx.h, C.byref(h)))
        self.h = h

    def __del__(self):
        try:
            if self.h:   # safe after Context.close(): the library orphans the batches of a destroyed context
                self.lib.b200he_batch_destroy(self.h)
        except Exception:
            pass
        self.h = None

    count = property(lambda s: int(s.lib.b200he_batch_count(s.h)))
    size = property(lambda s: s.lib.b200he_batch_size(s.h))
    L = property(lambda s: s.lib.b200he_batch_level(s.h))
    ntt_form = property(lambda s: bool(s.lib.b200he_batch_ntt_form(s.h)))
    scale = property(lambda s: s.lib.b200he_batch_scale(s.h))

    def set_scale(self, v):
        self.ctx._ck(self.lib.b200he_batch_set_scale(self.h, float(v)))

    def resize(self, count, size, L, ntt_form, scale=1.0):
        self.ctx._ck(self.lib.b200he_batch_resize(self.h, count, size, L, int(ntt_form), float(scale)))
        return self

    def upload(self, host, first=0):
        """host: uint64 array of n*size*L*N words (any shape)"""
        a = np.ascontiguousarray(host, dtype=np.uint64)
        words = self.size * self.L * self.ctx.N
        assert a.size % words == 0, (a.size, words)
        self.ctx._ck(self.lib.b200he_batch_upload(self.h, first, a.size // words, a.ctypes.data))
        self.ctx.sync()   # the host array may be a temporary
        return self

    def copy_from(self, other):
        """this batch = a copy of a batch of another context / GPU (device to device, stream-ordered)"""
        self.ctx._ck(self.lib.b200he_batch_copy_from(self.h, other.h))
        return self

    def upload_from(self, ptr, first, n):
        """raw pointer upload (pinned host memory), asynchronous"""
        self.ctx._ck(self.lib.b200he_batch_upload(self.h, first, n, ptr))

    def download_to(self, ptr, first, n, wait=True):
        fn = self.lib.b200he_batch_download if wait else self.lib.b200he_batch_download_async
        self.ctx._ck(fn(self.h, first, n, ptr))

    def upload_scattered(self, arrays, first=0):
        """load(): separately allocated host ciphertexts (one contiguous uint64 array each) -> ciphertexts
        [first, first+len) of the batch, through the context's pinned staging"""
        ptrs = (C.c_void_p * len(arrays))(*[a.ctypes.data for a in arrays])
        self.ctx._ck(self.lib.b200he_batch_upload_scattered(self.h, first, len(arrays), ptrs))

    def download_scattered(self, first=0, n=None):
        """store(): ciphertexts [first, first+n) -> n separately allocated host arrays"""
        n = self.count - first if n is None else n
        outs = [np.empty((self.size, self.L, self.ctx.N), dtype=np.uint64) for _ in range(n)]
        ptrs = (C.c_void_p * n)(*[a.ctypes.data for a in outs])
        self.ctx._ck(self.lib.b200he_batch_download_scattered(self.h, first, n, ptrs))
        return outs

    def download(self, first=0, n=None):
        n = self.count - first if n is None else n
        out = np.empty((n, self.size, self.L, self.ctx.N), dtype=np.uint64)
        if n:
            self.ctx._ck(self.lib.b200he_batch_download(self.h, first, n, out.ctypes.data))
        return out

    def device_ptr(self):
        return self.lib.b200he_batch_device_ptr(self.h)


class Context:
    """one GPU + one stream + tables + keys (replaces seal::SEALContext + seal::Evaluator, include/b200he.h)"""

    def __init__(self, scheme, N, moduli, psi, plain_modulus=0, device=0, lib=None):
        self.lib = lib or load_library()
        self.h = None
        self.scheme, self.N = scheme, N
        self.moduli = np.ascontiguousarray(moduli, dtype=np.uint64)
        self.psi = np.ascontiguousarray(psi, dtype=np.uint64)
        self.K = len(self.moduli)
        h = C.c_void_p()
        rc = self.lib.b200he_ctx_create(scheme, N, self.K, self.moduli.ctypes.data_as(u64p), self.psi.ctypes.data_as(u64p),
                                        int(plain_modulus), device, C.byref(h))
        if rc:
            raise B200HEError(self.lib.b200he_last_error().decode())
        self.h = h

    def close(self):
        if self.h:
            self.lib.b200he_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc:
            raise B200HEError(self.lib.b200he_last_error().decode())

    def sync(self):
        self._ck(self.lib.b200he_ctx_sync(self.h))

    def set_stream(self, cuda_stream):
        self._ck(self.lib.b200he_ctx_set_stream(self.h, cuda_stream))

    def set_workspace(self, nbytes):
        self._ck(self.lib.b200he_ctx_set_workspace(self.h, nbytes))

    def set_relin_key(self, key):
        k = np.ascontiguousarray(key, dtype=np.uint64)
        assert k.size == (self.K - 1) * 2 * self.K * self.N
        self._ck(self.lib.b200he_set_relin_key(self.h, k.ctypes.data_as(u64p)))

    def set_galois_key(self, elt, key):
        k = np.ascontiguousarray(key, dtype=np.uint64)
        assert k.size == (self.K - 1) * 2 * self.K * self.N
        self._ck(self.lib.b200he_set_galois_key(self.h, int(elt), k.ctypes.data_as(u64p)))

    def has_galois_key(self, elt):
        return bool(self.lib.b200he_has_galois_key(self.h, int(elt)))

    # ---- batches
    def batch(self, host=None, size=2, L=None, ntt_form=None, scale=1.0):
        b = Batch(self)
        if host is not None:
            a = np.ascontiguousarray(host, dtype=np.uint64)
            L = L if L is not None else self.K - 1
            ntt_form = (self.scheme == CKKS) if ntt_form is None else ntt_form
            words = size * L * self.N
            assert a.size % words == 0, (a.size, words)
            b.resize(a.size // words, size, L, ntt_form, scale)
            if a.size:
                b.upload(a)
        return b

    # ---- evaluator (each returns the output batch)
    def _binary(self, fn, a, b, ai, bi, n, out):
        out = out or Batch(self)
        ka, pa = _idx(ai)
        kb, pb = _idx(bi)
        if n is None:
            n = len(ka) if ka is not None else (len(kb) if kb is not None else min(a.count, b.count))
        self._ck(fn(self.h, a.h, pa, b.h, pb, n, out.h))
        return out

    def add(self, a, b, ai=None, bi=None, n=None, out=None):
        return self._binary(self.lib.b200he_add, a, b, ai, bi, n, out)

    def sub(self, a, b, ai=None, bi=None, n=None, out=None):
        return self._binary(self.lib.b200he_sub, a, b, ai, bi, n, out)

    def multiply(self, a, b, ai=None, bi=None, n=None, out=None):
        return self._binary(self.lib.b200he_multiply, a, b, ai, bi, n, out)

    def matmul_accumulate(self, a, b, rows, inner, cols, out=None):
        """out[i * cols + j] = sum_k a[i * inner + k] (x) b[j * inner + k] (size 3; b holds the right-hand matrix by columns):
        MatMult CipherBatchAxis' inner loop"""
        out = out or Batch(self)
        self._ck(self.lib.b200he_matmul_accumulate(self.h, a.h, b.h, rows, inner, cols, out.h))
        return out

    def _unary(self, fn, a, out, *args):
        out = out or Batch(self)
        self._ck(fn(self.h, a.h, *args, out.h))
        return out

    def relinearize(self, a, out=None):
        return self._unary(self.lib.b200he_relinearize, a, out)

    def rotate(self, a, step, out=None):
        return self._unary(self.lib.b200he_rotate, a, out, int(step))

    def rotate_each(self, a, steps, out=None):
        out = out or Batch(self)
        st = np.ascontiguousarray(steps, dtype=np.int32)
        assert len(st) == a.count
        self._ck(self.lib.b200he_rotate_each(self.h, a.h, st.ctypes.data_as(C.POINTER(C.c_int32)), out.h))
        return out

    def sum(self, a, out=None):
        return self._unary(self.lib.b200he_sum, a, out)

    def rotate_columns(self, a, out=None):
        return self._unary(self.lib.b200he_rotate_columns, a, out)

    def apply_galois(self, a, elt, out=None):
        return self._unary(self.lib.b200he_apply_galois, a, out, int(elt))

    def rescale_to_next(self, a, out=None):
        return self._unary(self.lib.b200he_rescale_to_next, a, out)

    def relinearize_rescale(self, a, out=None):
        return self._unary(self.lib.b200he_relinearize_rescale, a, out)

    def mod_drop(self, a, L_target, out=None):
        return self._unary(self.lib.b200he_mod_drop, a, out, int(L_target))

    def multiply_plain(self, ct, plain, pi=None, out=None):
        out = out or Batch(self)
        k, p = _idx(pi)
        self._ck(self.lib.b200he_multiply_plain(self.h, ct.h, plain.h, p, out.h))
        return out

    def add_plain(self, ct, plain, pi=None, out=None):
        out = out or Batch(self)
        k, p = _idx(pi)
        self._ck(self.lib.b200he_add_plain(self.h, ct.h, plain.h, p, out.h))
        return out

    def accumulate(self, a, count):
        self._ck(self.lib.b200he_accumulate(self.h, a.h, int(count)))
        return a

    def gather(self, a, idx=None, n=None, out=None):
        out = out or Batch(self)
        k, p = _idx(idx)
        n = (len(k) if k is not None else a.count) if n is None else n
        self._ck(self.lib.b200he_gather(self.h, a.h, p, n, out.h))
        return out

    def ntt_forward(self, a, out=None):
        return self._unary(self.lib.b200he_ntt_forward, a, out)

    def ntt_inverse(self, a, out=None):
        return self._unary(self.lib.b200he_ntt_inverse, a, out)

    # ---- measurement
    def launch_count(self):
        return int(self.lib.b200he_launch_count(self.h))

    def profile_begin(self):
        self._ck(self.lib.b200he_profile_begin(self.h))

    def profile_end(self):
        ms = (C.c_double * KERNEL_CLASSES)()
        n = (C.c_uint64 * KERNEL_CLASSES)()
        self._ck(self.lib.b200he_profile_end(self.h, ms, n))
        return {self.lib.b200he_kernel_name(i).decode(): (ms[i], int(n[i])) for i in range(KERNEL_CLASSES) if n[i]}

    def profile_work(self):
        """algorithmic work of the launches of the last profile: {kernel class: (integer-pipe butterfly equivalents,
        FP64-pipe butterfly equivalents, HBM bytes)}"""
        bi, bd, by = ((C.c_double * KERNEL_CLASSES)() for _ in range(3))
        self._ck(self.lib.b200he_profile_work(self.h, bi, bd, by))
        return {self.lib.b200he_kernel_name(i).decode(): (bi[i], bd[i], by[i]) for i in range(KERNEL_CLASSES) if bi[i] or bd[i] or by[i]}
