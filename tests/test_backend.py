"""The HEBench-facing plugin (libhebench_seal_backend.so): exports, descriptor set, and the full
encode -> encrypt -> load -> operate -> store -> decrypt -> decode flow of all 24 benchmarks (the reference's 20 + logistic regression with degree-5 / degree-7 sigmoids), validated
at value level by the mini harness like the reference's CI does with test_harness
(R/.github/workflows/cmake.yml:40-49: grep "Failed: 0")."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BACKEND = os.path.join(ROOT, "reference-seal-backend_b200", "backend")
PLUGIN = os.path.join(BACKEND, "libhebench_seal_backend.so")
HARNESS = os.path.join(BACKEND, "mini_harness")
API_SYMBOLS = ["initEngine", "destroyHandle", "subscribeBenchmarksCount", "subscribeBenchmarks", "getWorkloadParamsDetails",
               "describeBenchmark", "createBenchmark", "initBenchmark", "encode", "decode", "encrypt", "decrypt", "load", "store",
               "operate", "getSchemeName", "getSchemeSecurityName", "getBenchmarkDescriptionEx", "getErrorDescription",
               "getLastErrorDescription"]


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as g
    g.build_cuda()
    g.build_host()
    assert os.path.exists(PLUGIN) and os.path.exists(HARNESS)


def test_cuda_library_exports_every_header_symbol(built):
    """every function declared in include/b200he.h is exported by libb200he.so and bound by the Python layer"""
    import pyb200he
    hdr = open(os.path.join(ROOT, "include", "b200he.h")).read()
    declared = set(re.findall(r"\b(b200he_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(pyb200he.SIGNATURES), declared ^ set(pyb200he.SIGNATURES)
    lib = ctypes.CDLL(pyb200he.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert b"sm_100a" in pyb200he.declare(lib).b200he_version()


def test_plugin_exports_api_bridge(built):
    lib = ctypes.CDLL(PLUGIN)
    for name in API_SYMBOLS:
        assert hasattr(lib, name), name


def test_plugin_descriptor_set(built):
    """the reference's 20 descriptors in the reference's order (R/src/engine/seal_engine.cpp:108-151), then the
    api-bridge's degree-5 / degree-7 logistic-regression workloads (BASELINE.json configs[4]; not in the reference)"""
    out = subprocess.run([HARNESS, "--backend_lib_path", PLUGIN, "--list"], capture_output=True, text=True, check=True).stdout
    lines = [l for l in out.splitlines() if re.match(r"\s*\d+:", l)]
    assert len(lines) == 24
    want = (["EltwiseAdd BFV Latency", "EltwiseAdd CKKS Latency", "EltwiseAdd BFV Offline", "EltwiseAdd CKKS Offline",
             "EltwiseMultiply BFV Latency", "EltwiseMultiply CKKS Latency", "EltwiseMultiply BFV Offline", "EltwiseMultiply CKKS Offline",
             "DotProduct BFV Latency", "DotProduct CKKS Latency", "DotProduct BFV Offline", "DotProduct CKKS Offline",
             "MatrixMultiply BFV Latency other=1", "MatrixMultiply CKKS Latency other=1", "MatrixMultiply BFV Latency other=0",
             "MatrixMultiply CKKS Latency other=0", "MatrixMultiply BFV Latency other=2", "MatrixMultiply CKKS Latency other=2",
             "LogisticRegression_PolyD3 CKKS Latency other=1", "LogisticRegression_PolyD3 CKKS Offline other=1",
             "LogisticRegression_PolyD5 CKKS Latency other=1", "LogisticRegression_PolyD5 CKKS Offline other=1",
             "LogisticRegression_PolyD7 CKKS Latency other=1", "LogisticRegression_PolyD7 CKKS Offline other=1"])
    for line, w in zip(lines, want):
        assert w in line, (line, w)
    assert "n=1000 PolyModulusDegree=8192 MultiplicativeDepth=2 CoefficientModulusBits=45 ScaleBits=45" in lines[1]
    assert "PolyModulusDegree=16384 MultiplicativeDepth=6 CoefficientModulusBits=45" in lines[19]
    assert "PolyModulusDegree=32768 MultiplicativeDepth=8 CoefficientModulusBits=45" in lines[21]
    assert "PolyModulusDegree=32768 MultiplicativeDepth=10 CoefficientModulusBits=45" in lines[23]


def test_plugin_fails_loudly_without_gpu(built):
    """no CPU fallback: on a machine without a CUDA device createBenchmark reports the error through the API"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    p = subprocess.run([HARNESS, "--backend_lib_path", PLUGIN, "--filter", "EltwiseAdd CKKS Latency"], capture_output=True, text=True)
    assert p.returncode != 0 and "no CUDA device" in p.stdout and "Failed: 1" in p.stdout


def test_workload_logic_on_emulation(emu_lib):
    """TEST INFRASTRUCTURE: the same backend sources linked against the host-C++ emulation of the kernels, all 24
    benchmarks at N = 2048, validated against cleartext ground truth"""
    subprocess.check_call(["make", "-s", "-C", BACKEND, "emu"])
    emu_plugin = os.path.join(ROOT, "tests", "emu", "libhebench_seal_backend_emu.so")
    p = subprocess.run([HARNESS, "--backend_lib_path", emu_plugin, "--poly", "2048", "--n", "16", "--dims", "4,3,2", "--batch", "5",
                        "--iterations", "1"], capture_output=True, text=True, timeout=900)
    assert "[ Info    ] Total: 24" in p.stdout and "[ Info    ] Failed: 0" in p.stdout, p.stdout[-3000:]


@pytest.mark.gpu
def test_all_benchmarks_default_parameters_on_gpu(built):
    """the reference CI's check: every benchmark with its default parameters, decoded results validated"""
    p = subprocess.run([HARNESS, "--backend_lib_path", PLUGIN, "--random_seed", "1234"], capture_output=True, text=True, timeout=1500)
    assert "[ Info    ] Total: 24" in p.stdout and "[ Info    ] Failed: 0" in p.stdout, p.stdout[-4000:]


@pytest.mark.gpu
def test_baseline_config_shapes_on_gpu(built):
    """BASELINE.json configs at their stated shapes (scaled where the full batch only adds time):
    C1 BFV eltwise multiply n=100 N=8192; C3 CKKS dot n=100 N=16384; C5 CKKS logreg N=32768"""
    runs = [["--filter", "EltwiseMultiply BFV Offline", "--n", "100", "--samples", "10,10"],
            ["--filter", "DotProduct CKKS Offline", "--n", "100", "--poly", "16384", "--samples", "20,10"],
            ["--filter", "LogisticRegression_PolyD3 CKKS Offline", "--poly", "32768", "--batch", "64"],
            ["--filter", "LogisticRegression_PolyD7 CKKS Offline", "--batch", "64"]]
    for extra in runs:
        p = subprocess.run([HARNESS, "--backend_lib_path", PLUGIN] + extra, capture_output=True, text=True, timeout=1500)
        assert "[ Info    ] Failed: 0" in p.stdout and "Total: 1" in p.stdout, p.stdout[-3000:]


def test_cmake_project_configures_offline_and_rejects_missing_seal(tmp_path):
    """the CMake project with the reference's -D{SEAL,API_BRIDGE}_INSTALL_DIR options
    (R/cmake/utils/import-library.cmake:54-66).  Offline it configures with the stand-ins; a SEAL prefix that holds no SEAL
    is a hard error, as in the reference (the real-SEAL adapter hostfhe/seal_adapter.cpp cannot be compiled here)."""
    import shutil
    if not shutil.which("cmake"):
        pytest.skip("cmake not installed")
    src = os.path.join(ROOT, "reference-seal-backend_b200")
    env = dict(os.environ, CXX="g++", CUDAHOSTCXX="g++")
    p = subprocess.run(["cmake", "-S", src, "-B", str(tmp_path / "b"), "-G", "Ninja"], capture_output=True, text=True, env=env, timeout=600)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    assert "host crypto: hostfhe stand-in" in p.stdout and "api-bridge: backend/compat restatement" in p.stdout
    q = subprocess.run(["cmake", "-S", src, "-B", str(tmp_path / "c"), "-G", "Ninja", "-DSEAL_INSTALL_DIR=" + str(tmp_path / "nowhere")],
                       capture_output=True, text=True, env=env, timeout=600)
    assert q.returncode != 0 and "FAILED TO FIND PRE-INSTALLED SEAL" in q.stderr
    adapter = open(os.path.join(src, "hostfhe", "seal_adapter.cpp")).read()
    hdr = open(os.path.join(src, "hostfhe", "hostfhe.h")).read()
    for fn in re.findall(r"\b(hfhe_[a-z0-9_]+)\s*\(", hdr):   # the adapter implements the whole hostfhe.h interface
        assert re.search(r'extern "C" [^;{]*\b' + fn + r"\s*\(", adapter), fn
