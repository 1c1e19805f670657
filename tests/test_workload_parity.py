"""Workload-level bit-exact parity (SURVEY.md §8c "workload-level"): the ciphertexts the plugin's store() returns vs the
CPU oracle's restatement of the reference's operate() bodies (oracle/he_oracle_workloads.cpp), on the SAME inputs.

The mini harness drives libhebench_seal_backend.so in HEBench order with HEB_B200_TRACE_DIR set, so the plugin writes the
ciphertexts that crossed load()/store() and the fresh encryptions it drew inside operate() (the reference's randomised
encrypt_zero / encrypt, R/src/engine/seal_context.cpp:360,440).  HEB_B200_SEED makes the host stand-in's keys
reproducible, so the test regenerates the same relinearization / Galois keys, replays the inputs through the oracle and
compares every output word.

  * CPU (not gpu): the same backend sources linked against the host-C++ emulation of the kernels, N = 2048;
  * GPU: the real plugin at the BASELINE.json chains -- reduced shapes compared in full, and the FULL shapes
    (C3 100x100, C4 100x100x100 for all three algorithms, C5 batch 1024) validated at value level by the harness with
    a few result cells traced (HEB_B200_TRACE_PICK_*) and compared bit for bit.
"""
import os
import struct
import subprocess

import numpy as np
import pytest

from helpers import BFV, CKKS, Host, Oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BACKEND = os.path.join(ROOT, "reference-seal-backend_b200", "backend")
PLUGIN = os.path.join(BACKEND, "libhebench_seal_backend.so")
EMU_PLUGIN = os.path.join(ROOT, "tests", "emu", "libhebench_seal_backend_emu.so")
HARNESS = os.path.join(BACKEND, "mini_harness")
SEED = 20261018
SIGMOID = {3: [0.5, 0.15012, 0.0, -0.0015930078125],
           5: [0.5, 0.19131, 0.0, -0.0045963, 0.0, 0.0000412332],
           7: [0.5, 0.21687, 0.0, -0.0081918, 0.0, 0.000165838, 0.0, -0.00000119581]}


def read_trace(d, tag):
    """-> (uint64 array [count][size][L][N], scale, items in the traced vector)"""
    with open(os.path.join(d, tag + ".bin"), "rb") as f:
        magic, count, size, L, N, ntt, scale_bits, total = struct.unpack("<8Q", f.read(64))
        assert magic == 0x3143525430303242, hex(magic)
        data = np.fromfile(f, dtype=np.uint64)
    assert data.size == count * size * L * N, (tag, data.size, count, size, L, N)
    return data.reshape(count, size, L, N), struct.unpack("<d", struct.pack("<Q", scale_bits))[0], total


GPUS = [None]   # HEB_B200_GPUS of the harness runs (None: the environment's / one GPU)


def run_harness(plugin, args, trace_dir, picks=None, timeout=1500):
    # HEB_B200_POPULATE_MIN_MB=0: the result arenas of store() (made and touched during operate(), for results of a gigabyte
    # and more by default) are used at every size, so these tests cover that path
    env = dict(os.environ, HEB_B200_TRACE_DIR=str(trace_dir), HEB_B200_SEED=str(SEED), HEB_B200_POPULATE_MIN_MB="0")
    if GPUS[0]:
        env["HEB_B200_GPUS"] = str(GPUS[0])
    for tag, ids in (picks or {}).items():
        env["HEB_B200_TRACE_PICK_" + tag] = ",".join(str(i) for i in ids)
    p = subprocess.run([HARNESS, "--backend_lib_path", plugin, "--iterations", "1"] + args, capture_output=True, text=True, env=env, timeout=timeout)
    assert "[ Info    ] Failed: 0" in p.stdout and "Total: 1" in p.stdout, p.stdout[-3000:] + p.stderr[-2000:]
    return p.stdout


class Keys:
    """the plugin's keys, regenerated from the same seed (every key has its own keystream in hostfhe)"""

    def __init__(self, scheme, N, depth, coeff_bits, sp_bits, galois=True):
        self.host = Host(scheme, N, depth, coeff_bits, sp_bits, seed=SEED)
        self.orc = Oracle(scheme, N, self.host.moduli, self.host.t)
        self.relin = self.host.relin_key()
        self.gal = {e: self.host.galois_key(e) for e in self.host.galois_elts()} if galois else {}
        self.N, self.Ltop = N, depth


def eq(got, want, what):
    got, want = np.asarray(got).reshape(-1), np.asarray(want).reshape(-1)
    assert got.shape == want.shape, (what, got.shape, want.shape)
    if not np.array_equal(got, want):
        bad = np.nonzero(got != want)[0]
        raise AssertionError(f"{what}: {len(bad)} of {got.size} words differ, first at {bad[0]}")


# ------------------------------------------------------------------------------------------- per-workload checkers
def check_vector(plugin, tmp, scheme, op, N, depth, coeff_bits, sp_bits, n, s0, s1, pick_out=None, sub=None):
    """sub = (v0, b0, v1, b1): operate() on the ParameterIndexer range [v0, v0+b0) x [v1, v1+b1) of the loaded samples
    (R/src/benchmarks/ckks/seal_ckks_element_wise_benchmark.cpp:334-336); result r is sample pair (v0 + r // b1, v1 + r % b1)"""
    name = {"add": "EltwiseAdd", "mul": "EltwiseMultiply", "dot": "DotProduct"}[op]
    args = ["--filter", f"{name} {'CKKS' if scheme == CKKS else 'BFV'} Offline", "--n", str(n), "--samples", f"{s0},{s1}", "--poly", str(N), "--depth", str(depth)]
    v0, b0, v1, b1 = sub if sub else (0, s0, 0, s1)
    if sub:
        args += ["--sub", ",".join(str(v) for v in sub)]
    run_harness(plugin, args, tmp, {"out": pick_out} if pick_out else None)
    k = Keys(scheme, N, depth, coeff_bits, sp_bits, galois=(op == "dot"))
    a, _, _ = read_trace(tmp, "in0")
    b, _, _ = read_trace(tmp, "in1")
    out, _, total = read_trace(tmp, "out")
    assert a.shape[0] == s0 and b.shape[0] == s1 and total == b0 * b1
    cells = list(pick_out) if pick_out else list(range(b0 * b1))
    L = depth
    for r, cell in enumerate(cells):
        x, y = a[v0 + cell // b1].reshape(-1), b[v1 + cell % b1].reshape(-1)
        if op == "add":
            want = k.orc.add(L, 2, x, y)
        elif op == "mul":
            want = k.orc.ckks_multiply(L, x, y) if scheme == CKKS else k.orc.bfv_multiply(x, y)
        else:
            want = k.orc.batch_dot(L, 1, x, y, n, k.relin, k.gal)
        eq(out[r], want, f"{name} result cell {cell}")


def check_matmul(plugin, tmp, scheme, algo, N, depth, coeff_bits, sp_bits, dims, cells=None):
    """algo: 0 Val, 1 CipherBatchAxis, 2 Row.  cells: result cells (i, j) to trace and compare (None = everything)"""
    r0, c0, c1 = dims
    args = ["--filter", f"MatrixMultiply {'CKKS' if scheme == CKKS else 'BFV'} Latency other={algo}", "--dims", f"{r0},{c0},{c1}", "--poly", str(N), "--depth", str(depth)]
    picks = None
    if cells is not None:
        rows, cols = sorted({i for i, _ in cells}), sorted({j for _, j in cells})
        if algo == 0:
            picks = {"in0": rows, "in1": cols, "out": [i * c1 + j for i in rows for j in cols]}
        elif algo == 1:
            picks = {"in0": [i * c0 + k for i in rows for k in range(c0)], "in1": [k * c1 + j for k in range(c0) for j in cols],
                     "out": [i * c1 + j for i in rows for j in cols]}
        else:   # Row: one ciphertext per row of M0 (CKKS)
            picks = {"in0": rows, "out": rows}
    run_harness(plugin, args, tmp, picks)
    k = Keys(scheme, N, depth, coeff_bits, sp_bits, galois=(algo != 1))
    m0, _, _ = read_trace(tmp, "in0")
    m1, _, _ = read_trace(tmp, "in1")
    out, _, _ = read_trace(tmp, "out")
    L = depth
    if algo == 0:
        nr, nc = (len(rows), len(cols)) if cells is not None else (r0, c1)
        want = k.orc.matmul_val(L, nr, c0, nc, m0.reshape(-1), m1.reshape(-1), k.relin, k.gal)
    elif algo == 1:
        nr, nc = (len(rows), len(cols)) if cells is not None else (r0, c1)
        want = k.orc.matmul_cba(L, nr, c0, nc, m0.reshape(-1), m1.reshape(-1), k.relin)
    else:
        slots = N // 2
        spacers = (slots if scheme == CKKS else slots) // c0   # BFV: row size N/2 as well
        want = k.orc.matmul_row(L, m0.shape[0], c0, spacers, m0.reshape(-1), m1.reshape(-1), k.relin, k.gal)
    eq(out, want, f"MatMult algo {algo} {dims}")


def check_logreg(plugin, tmp, N, depth, degree, batch, n_features=16):
    args = ["--filter", f"LogisticRegression_PolyD{degree} CKKS Offline", "--poly", str(N), "--depth", str(depth), "--batch", str(batch), "--n", str(n_features)]
    run_harness(plugin, args, tmp)
    k = Keys(CKKS, N, depth, 45, 45)
    h, Ltop = k.host, depth
    W, _, _ = read_trace(tmp, "W")
    b, _, _ = read_trace(tmp, "b")
    X, _, _ = read_trace(tmp, "X")
    zero, _, _ = read_trace(tmp, "collapse_zero")
    seed, _, _ = read_trace(tmp, "horner_seed")
    out, _, _ = read_trace(tmp, "out")
    assert X.shape[0] == batch
    # the injected plaintexts are deterministic encodings: regenerated here, independently of the plugin
    masks = np.empty((batch, Ltop - 1, N), dtype=np.uint64)
    for i in range(batch):
        e = np.zeros(batch)
        e[i] = 1.0
        masks[i] = h.encode(e).reshape(Ltop, N)[: Ltop - 1]
    coeff = SIGMOID[degree]
    cf = np.stack([h.encode(np.full(N // 2, coeff[j])) for j in range(degree - 1, -1, -1)])
    seed_plain_check = h.decode(h.decrypt(seed[0].reshape(-1), 2, Ltop), Ltop)
    assert abs(seed_plain_check[0] - coeff[degree]) < 1e-6, "the traced Horner seed does not encrypt the leading coefficient"
    want, lv = k.orc.logreg(n_features, batch, W.reshape(-1), b.reshape(-1), X.reshape(-1), masks.reshape(-1), zero.reshape(-1), seed.reshape(-1),
                            cf.reshape(-1), k.relin, k.gal)
    assert out.shape[2] == lv, (out.shape, lv)
    eq(out, want, f"LogReg D{degree} batch {batch}")


# ------------------------------------------------------------------------------------------- CPU: emulation plugin
@pytest.fixture(scope="module")
def emu_plugin(emu_lib):
    if os.environ.get("B200HE_EMU_PLUGIN"):   # a sanitizer build of the plugin (tests/emu/run_asan.sh)
        return os.environ["B200HE_EMU_PLUGIN"]
    import __graft_entry__ as g
    g.build_host()
    subprocess.check_call(["make", "-s", "-C", BACKEND, "emu"])
    return EMU_PLUGIN


def test_emu_vector_workloads(emu_plugin, tmp_path):
    check_vector(emu_plugin, tmp_path, CKKS, "add", 2048, 2, 45, 45, n=16, s0=2, s1=3)
    check_vector(emu_plugin, tmp_path, CKKS, "mul", 2048, 2, 45, 45, n=16, s0=2, s1=2)
    check_vector(emu_plugin, tmp_path, CKKS, "dot", 2048, 2, 40, 40, n=10, s0=2, s1=2)
    check_vector(emu_plugin, tmp_path, BFV, "mul", 2048, 2, 40, 20, n=16, s0=2, s1=1)
    check_vector(emu_plugin, tmp_path, BFV, "dot", 2048, 2, 45, 20, n=9, s0=1, s1=2)


def test_emu_parameter_indexers(emu_plugin, tmp_path):
    """operate() honours ParameterIndexer sub-ranges for the element-wise and dot-product workloads (result order
    (i - v0) * b1 + (j - v1)), on one GPU and with the result space split over several, and rejects ranges that leave the
    loaded samples; the matrix and logistic-regression workloads accept the full range only
    (R/src/benchmarks/ckks/seal_ckks_element_wise_benchmark.cpp:321-336, ...matmultval...:451-461, ...logreg_horner.cpp:413-419)"""
    check_vector(emu_plugin, tmp_path, CKKS, "mul", 2048, 2, 45, 45, n=16, s0=4, s1=3, sub=(1, 2, 1, 2))
    check_vector(emu_plugin, tmp_path, CKKS, "dot", 2048, 2, 40, 40, n=10, s0=3, s1=5, sub=(2, 1, 1, 3))
    check_vector(emu_plugin, tmp_path, BFV, "mul", 2048, 2, 40, 20, n=16, s0=3, s1=2, sub=(0, 3, 1, 1))
    GPUS[0] = 3
    try:
        check_vector(emu_plugin, tmp_path, CKKS, "add", 2048, 2, 45, 45, n=16, s0=5, s1=2, sub=(1, 3, 0, 2))   # range spans two GPUs' blocks
        check_vector(emu_plugin, tmp_path, CKKS, "dot", 2048, 2, 40, 40, n=10, s0=2, s1=7, sub=(0, 2, 3, 3))   # parameter 1 split
    finally:
        GPUS[0] = 0
    base = ["--backend_lib_path", emu_plugin, "--iterations", "1", "--expect-operate-error", "--poly", "2048", "--depth", "2"]
    env = dict(os.environ, HEB_B200_SEED=str(SEED))
    for extra, message in (
            (["--filter", "EltwiseAdd CKKS Offline", "--n", "8", "--samples", "3,2", "--sub", "2,2,0,2"], "Invalid parameter indexer"),     # 2 + 2 > 3
            (["--filter", "DotProduct CKKS Offline", "--n", "8", "--samples", "2,2", "--sub", "0,2,1,2"], "Invalid parameter indexer"),     # 1 + 2 > 2
            (["--filter", "MatrixMultiply CKKS Latency other=0", "--dims", "2,3,2", "--sub", "1,1,0,1"], "Unexpected index"),              # value_index > 0
            (["--filter", "MatrixMultiply CKKS Latency other=1", "--dims", "2,3,2", "--depth", "3", "--sub", "0,2,0,1"], "Batch size"),      # batch_size > 1
            (["--filter", "LogisticRegression_PolyD3 CKKS Offline", "--batch", "4", "--depth", "6", "--sub", "1,3,0,0"], "indexer")):        # not the whole batch
        p = subprocess.run([HARNESS] + base[:-2] + extra + ([] if "--depth" in extra else ["--depth", "2"]), capture_output=True, text=True, env=env, timeout=600)
        assert "rejected the indexers as expected" in p.stdout and "Failed: 0" in p.stdout, p.stdout[-2000:] + p.stderr[-1000:]
        assert message in p.stdout, p.stdout[-2000:]


@pytest.mark.parametrize("scheme", [CKKS, BFV], ids=["ckks", "bfv"])
def test_emu_matmul_workloads(emu_plugin, tmp_path, scheme):
    bits = (45, 45) if scheme == CKKS else (40, 20)
    check_matmul(emu_plugin, tmp_path, scheme, 0, 2048, 2, *bits, dims=(3, 5, 2))
    check_matmul(emu_plugin, tmp_path, scheme, 1, 2048, 3, *bits, dims=(2, 3, 2))
    check_matmul(emu_plugin, tmp_path, scheme, 2, 2048, 3, *bits, dims=(3, 4, 2))


def test_emu_logreg_workload(emu_plugin, tmp_path):
    check_logreg(emu_plugin, tmp_path, 2048, 6, 3, batch=5, n_features=6)


def test_emu_result_space_partitioning(emu_plugin, tmp_path):
    """with several GPUs (here: several contexts of the emulation) the result space is partitioned -- the longer
    operand split in blocks, the other replicated; matrix rows or columns; logistic-regression samples with the partial
    collapse sums exchanged device to device -- and store() still returns the reference's bits in the reference's order"""
    GPUS[0] = 3
    try:
        check_vector(emu_plugin, tmp_path, CKKS, "mul", 2048, 2, 45, 45, n=16, s0=4, s1=2)    # parameter 0 split
        check_vector(emu_plugin, tmp_path, CKKS, "dot", 2048, 2, 40, 40, n=10, s0=2, s1=5)    # parameter 1 split: results interleave
        check_vector(emu_plugin, tmp_path, BFV, "dot", 2048, 2, 45, 20, n=9, s0=1, s1=2)      # fewer samples than GPUs
        check_matmul(emu_plugin, tmp_path, CKKS, 0, 2048, 2, 45, 45, dims=(2, 5, 4))          # Val, columns of M1 split
        check_matmul(emu_plugin, tmp_path, CKKS, 1, 2048, 3, 45, 45, dims=(4, 3, 2))          # CipherBatchAxis, rows of M0 split
        check_matmul(emu_plugin, tmp_path, CKKS, 1, 2048, 3, 45, 45, dims=(2, 3, 5))          # CipherBatchAxis, columns of M1 split
        check_matmul(emu_plugin, tmp_path, BFV, 1, 2048, 3, 40, 20, dims=(2, 3, 4))
        check_matmul(emu_plugin, tmp_path, CKKS, 2, 2048, 3, 45, 45, dims=(4, 4, 2))          # Row
        check_logreg(emu_plugin, tmp_path, 2048, 6, 3, batch=7, n_features=6)
    finally:
        GPUS[0] = None


# ------------------------------------------------------------------------------------------- GPU: the real plugin
@pytest.fixture(scope="module")
def plugin():
    assert os.path.exists(PLUGIN) and os.path.exists(HARNESS)
    return PLUGIN


@pytest.mark.gpu
def test_gpu_vector_workloads(plugin, tmp_path):
    """C1 (BFV eltwise multiply n = 100, N = 8192) and C2 shapes compared in full; C3 (dot n = 100, N = 16384) on a 3 x 2 grid"""
    check_vector(plugin, tmp_path, BFV, "mul", 8192, 2, 40, 20, n=100, s0=3, s1=2)
    check_vector(plugin, tmp_path, CKKS, "add", 8192, 2, 45, 45, n=1000, s0=4, s1=3)
    check_vector(plugin, tmp_path, CKKS, "mul", 8192, 2, 45, 45, n=1000, s0=4, s1=3)
    check_vector(plugin, tmp_path, CKKS, "dot", 16384, 2, 40, 40, n=100, s0=3, s1=2)
    check_vector(plugin, tmp_path, BFV, "dot", 8192, 2, 45, 20, n=100, s0=2, s1=2)


@pytest.mark.gpu
def test_gpu_parameter_indexers(plugin, tmp_path):
    """ParameterIndexer sub-ranges through the CUDA path: the C2 and C3 parameter sets on a window of the loaded samples"""
    check_vector(plugin, tmp_path, CKKS, "mul", 8192, 2, 45, 45, n=1000, s0=5, s1=4, sub=(1, 3, 2, 2))
    check_vector(plugin, tmp_path, CKKS, "dot", 16384, 2, 40, 40, n=100, s0=3, s1=4, sub=(2, 1, 1, 2))


@pytest.mark.gpu
def test_gpu_matmul_workloads_reduced(plugin, tmp_path):
    """the C4 chain (N = 16384, {60, 45 x 5, 60}) at shapes the oracle finishes in seconds, every result ciphertext compared;
    MatMultRow at its own N = 32768; the BFV twins at their defaults"""
    check_matmul(plugin, tmp_path, CKKS, 0, 16384, 6, 45, 45, dims=(3, 9, 2))
    check_matmul(plugin, tmp_path, CKKS, 1, 16384, 6, 45, 45, dims=(2, 7, 3))
    check_matmul(plugin, tmp_path, CKKS, 2, 32768, 3, 45, 45, dims=(2, 6, 5))
    check_matmul(plugin, tmp_path, BFV, 0, 8192, 2, 40, 20, dims=(3, 9, 2))
    check_matmul(plugin, tmp_path, BFV, 1, 8192, 3, 40, 20, dims=(2, 3, 2))
    check_matmul(plugin, tmp_path, BFV, 2, 8192, 3, 40, 20, dims=(4, 5, 3))


@pytest.mark.gpu
def test_gpu_logreg_workloads_reduced(plugin, tmp_path):
    """C5 chain (N = 32768, K = 7): collapse + bias + Horner composites bit for bit, degree 3 and degree 7"""
    check_logreg(plugin, tmp_path, 32768, 6, 3, batch=9)
    check_logreg(plugin, tmp_path, 32768, 10, 7, batch=3)


@pytest.mark.gpu
def test_gpu_c3_dot_full_shape(plugin, tmp_path):
    """BASELINE configs[2] at its stated shape: 100 x 100 = 10^4 dot products of n = 100 at N = 16384; all validated at value
    level by the harness, 4 result cells bit for bit"""
    check_vector(plugin, tmp_path, CKKS, "dot", 16384, 2, 40, 40, n=100, s0=100, s1=100, pick_out=[0, 4242, 7077, 9999])


@pytest.mark.gpu
def test_gpu_c4_matmul_val_full_shape(plugin, tmp_path):
    """BASELINE configs[3]: MatMultVal 100 x 100 x 100, N = 16384, 6 data limbs"""
    check_matmul(plugin, tmp_path, CKKS, 0, 16384, 6, 45, 45, dims=(100, 100, 100), cells=[(0, 0), (57, 99)])


@pytest.mark.gpu
def test_gpu_c4_matmul_cipherbatchaxis_full_shape(plugin, tmp_path):
    """BASELINE configs[3]: CipherBatchAxis 100 x 100 x 100 (2 x 10^4 input ciphertexts, 31.5 GB in HBM)"""
    check_matmul(plugin, tmp_path, CKKS, 1, 16384, 6, 45, 45, dims=(100, 100, 100), cells=[(3, 7), (99, 0)])


@pytest.mark.gpu
def test_gpu_c4_matmul_row_full_shape(plugin, tmp_path):
    """BASELINE configs[3]: MatMultRow 100 x 100 x 100 needs cols_M0 * cols_M1 <= N / 2, i.e. N = 32768
    (R/src/benchmarks/ckks/seal_ckks_matmult_row_benchmark.cpp:142)"""
    check_matmul(plugin, tmp_path, CKKS, 2, 32768, 3, 45, 45, dims=(100, 100, 100), cells=[(0, 0), (63, 0)])


@pytest.mark.gpu
def test_gpu_c5_logreg_full_batch(plugin, tmp_path):
    """BASELINE configs[4] at batch 1024 (N = 32768): validated at value level by the harness (the single result ciphertext
    depends on every sample, so the bit-level comparison lives in the reduced-batch test above)"""
    env = dict(os.environ, HEB_B200_SEED=str(SEED))
    p = subprocess.run([HARNESS, "--backend_lib_path", plugin, "--iterations", "1", "--filter", "LogisticRegression_PolyD3 CKKS Offline", "--poly", "32768",
                        "--batch", "1024"], capture_output=True, text=True, env=env, timeout=1500)
    assert "[ Info    ] Failed: 0" in p.stdout and "Total: 1" in p.stdout, p.stdout[-3000:]
