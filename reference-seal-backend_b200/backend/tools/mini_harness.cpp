// mini_harness.cpp -- a small stand-in for HEBench's test_harness (upstream hebench/frontend, not
// available offline).  It dlopen()s a backend, binds the API Bridge C ABI with dlsym and drives every
// subscribed benchmark in HEBench order:
//   describe -> createBenchmark -> initBenchmark -> encode -> encrypt -> load -> operate (warm-up, timed) -> store -> decrypt -> decode
// then validates the decoded values against a cleartext ground truth with a relative tolerance and prints
// the line the reference's CI greps ("[ Info    ] Failed: 0", R/.github/workflows/validate_testharness_output.sh:7).
//
//   mini_harness --backend_lib_path libhebench_seal_backend.so [--random_seed 1234] [--filter TEXT] [--list]
//                [--samples A,B] [--batch N] [--iterations K] [--n N] [--dims R,C0,C1] [--poly N] [--depth D] [--csv FILE]
//                [--sub V0,B0,V1,B1]    element-wise / dot product: operate() on that index range of the loaded samples (ParameterIndexer)
//                [--expect-operate-error]   the run passes when operate() rejects the request
//                [--json FILE]          one JSON object per benchmark: wall time of every phase (encode ... decode), every timed operate()
//                [--extra-operate K]    K more untimed operate() calls after the timed ones (the backend profiles those, HEB_B200_PROFILE_*)
#include <dlfcn.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <random>
#include <sstream>
#include <string>
#include <vector>

#include "hebench/api_bridge/api.h"

using namespace hebench::APIBridge;

#define BIND(name) decltype(&hebench::APIBridge::name) p_##name = reinterpret_cast<decltype(&hebench::APIBridge::name)>(dlsym(lib, #name)); \
    if (!p_##name) { fprintf(stderr, "missing symbol %s\n", #name); return 2; }

struct Options {
    std::string lib, filter, csv, json;
    unsigned seed = 1234;
    bool list = false;
    uint64_t samples[2] = { 2, 3 }, batch = 8, iterations = 2, n = 0, dims[3] = { 0, 0, 0 }, poly = 0, depth = 0, extra_operate = 0;
    uint64_t sub[4] = { 0, 0, 0, 0 };   // --sub v0,b0,v1,b1: operate() on the index range [v0, v0+b0) x [v1, v1+b1) of the loaded samples
    bool have_sub = false, expect_operate_error = false;
};

static const char *workloadName(Workload w)
{
    switch (w) {
    case EltwiseAdd: return "EltwiseAdd";
    case EltwiseMultiply: return "EltwiseMultiply";
    case DotProduct: return "DotProduct";
    case MatrixMultiply: return "MatrixMultiply";
    case LogisticRegression_PolyD3: return "LogisticRegression_PolyD3";
    case LogisticRegression_PolyD5: return "LogisticRegression_PolyD5";
    case LogisticRegression_PolyD7: return "LogisticRegression_PolyD7";
    default: return "Other";
    }
}

static bool is_logreg(Workload w) { return w == LogisticRegression_PolyD3 || w == LogisticRegression_PolyD5 || w == LogisticRegression_PolyD7; }
// ground truth of the sigmoid approximations (HEBench's logistic-regression workloads, degree 3 / 5 / 7)
static double sigmoid_poly(Workload w, double x)
{
    const double x2 = x * x;
    if (w == LogisticRegression_PolyD5) return 0.5 + x * (0.19131 + x2 * (-0.0045963 + x2 * 0.0000412332));
    if (w == LogisticRegression_PolyD7) return 0.5 + x * (0.21687 + x2 * (-0.0081918 + x2 * (0.000165838 + x2 * -0.00000119581)));
    return 0.5 + 0.15012 * x - 0.0015930078125 * x * x * x;
}

struct NativeData {   // owns the buffers of a DataPackCollection
    std::vector<std::vector<std::vector<unsigned char>>> bytes;   // [pack][buffer]
    std::vector<std::vector<DataBuffer>> buffers;
    std::vector<DataPack> packs;
    DataPackCollection coll;
    void build()
    {
        buffers.resize(bytes.size());
        packs.resize(bytes.size());
        for (size_t p = 0; p < bytes.size(); ++p) {
            buffers[p].resize(bytes[p].size());
            for (size_t b = 0; b < bytes[p].size(); ++b) buffers[p][b] = DataBuffer{ bytes[p][b].data(), bytes[p][b].size(), 0 };
            packs[p] = DataPack{ buffers[p].data(), buffers[p].size(), p };
        }
        coll = DataPackCollection{ packs.data(), packs.size() };
    }
};

template <class T> static T *as(std::vector<unsigned char> &v) { return reinterpret_cast<T *>(v.data()); }

static bool almostEqual(double a, double b)
{
    const double tol = 0.01;   // relative; values flushed to 0 below 5e-5 by the backend are compared absolutely
    const double m   = std::max(std::fabs(a), std::fabs(b));
    return std::fabs(a - b) <= tol * m || std::fabs(a - b) < 1e-3;
}

int main(int argc, char **argv)
{
    Options o;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        auto next = [&]() { return std::string(i + 1 < argc ? argv[++i] : ""); };
        if (a == "--backend_lib_path") o.lib = next();
        else if (a == "--random_seed") o.seed = (unsigned)atol(next().c_str());
        else if (a == "--filter") o.filter = next();
        else if (a == "--csv") o.csv = next();
        else if (a == "--json") o.json = next();
        else if (a == "--extra-operate") o.extra_operate = strtoull(next().c_str(), nullptr, 10);
        else if (a == "--list") o.list = true;
        else if (a == "--samples") sscanf(next().c_str(), "%lu,%lu", &o.samples[0], &o.samples[1]);
        else if (a == "--sub") { sscanf(next().c_str(), "%lu,%lu,%lu,%lu", &o.sub[0], &o.sub[1], &o.sub[2], &o.sub[3]); o.have_sub = true; }
        else if (a == "--expect-operate-error") o.expect_operate_error = true;
        else if (a == "--batch") o.batch = strtoull(next().c_str(), nullptr, 10);
        else if (a == "--iterations") o.iterations = strtoull(next().c_str(), nullptr, 10);
        else if (a == "--n") o.n = strtoull(next().c_str(), nullptr, 10);
        else if (a == "--dims") sscanf(next().c_str(), "%lu,%lu,%lu", &o.dims[0], &o.dims[1], &o.dims[2]);
        else if (a == "--poly") o.poly = strtoull(next().c_str(), nullptr, 10);
        else if (a == "--depth") o.depth = strtoull(next().c_str(), nullptr, 10);
        else { fprintf(stderr, "unknown option %s\n", a.c_str()); return 2; }
    }
    if (o.lib.empty()) { fprintf(stderr, "usage: mini_harness --backend_lib_path <lib.so> [options]\n"); return 2; }
    void *lib = dlopen(o.lib.c_str(), RTLD_NOW | RTLD_LOCAL);
    if (!lib) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 2; }
    BIND(initEngine) BIND(destroyHandle) BIND(subscribeBenchmarksCount) BIND(subscribeBenchmarks) BIND(getWorkloadParamsDetails)
    BIND(describeBenchmark) BIND(createBenchmark) BIND(initBenchmark) BIND(encode) BIND(decode) BIND(encrypt) BIND(decrypt) BIND(load)
    BIND(store) BIND(operate) BIND(getSchemeName) BIND(getSchemeSecurityName) BIND(getBenchmarkDescriptionEx) BIND(getErrorDescription)
    BIND(getLastErrorDescription)

    Handle engine{};
    if (p_initEngine(&engine, nullptr, 0)) { fprintf(stderr, "initEngine failed\n"); return 2; }
    auto lastError = [&]() {
        char buf[1024] = { 0 };
        p_getLastErrorDescription(engine, buf, sizeof buf);
        return std::string(buf);
    };
    uint64_t count = 0;
    p_subscribeBenchmarksCount(engine, &count);
    std::vector<Handle> descs(count);
    p_subscribeBenchmarks(engine, descs.data(), count);
    printf("[ Info    ] backend %s: %lu benchmarks subscribed\n", o.lib.c_str(), count);

    std::ofstream csv;
    if (!o.csv.empty()) {
        csv.open(o.csv);
        csv << "index,workload,scheme,category,other,params,results_per_operate,operate_ms,samples_per_s,load_ms,store_ms,validated,failed_values\n";
    }
    std::ofstream json;
    if (!o.json.empty()) json.open(o.json);
    std::mt19937_64 rng(o.seed);
    uint64_t failed = 0, ran = 0;
    for (uint64_t bi = 0; bi < count; ++bi) {
        uint64_t n_params = 0, n_defaults = 0;
        p_getWorkloadParamsDetails(engine, descs[bi], &n_params, &n_defaults);
        std::vector<WorkloadParam> wp(n_params);
        WorkloadParams wps{ wp.data(), n_params };
        BenchmarkDescriptor bd;
        p_describeBenchmark(engine, descs[bi], &bd, &wps, 1);
        char scheme[64] = { 0 }, sec[64] = { 0 };
        p_getSchemeName(engine, bd.scheme, scheme, sizeof scheme);
        p_getSchemeSecurityName(engine, bd.scheme, bd.security, sec, sizeof sec);
        std::ostringstream title;
        title << workloadName(bd.workload) << " " << scheme << " " << (bd.category == Latency ? "Latency" : "Offline") << " other=" << bd.other;
        if (o.list) {
            printf("%2lu: %s | %s |", bi, title.str().c_str(), sec);
            for (auto &p : wp) printf(" %s=%lu", p.name, p.u_param);
            printf("\n");
            continue;
        }
        if (!o.filter.empty() && title.str().find(o.filter) == std::string::npos) continue;
        // parameter overrides
        const bool is_mat = bd.workload == MatrixMultiply;
        if (is_mat) { for (int k = 0; k < 3; ++k) if (o.dims[k]) wp[k].u_param = o.dims[k]; }
        else if (o.n) wp[0].u_param = o.n;
        if (o.poly) wp[is_mat ? 3 : 1].u_param = o.poly;
        if (o.depth) wp[is_mat ? 4 : 2].u_param = o.depth;   // MultiplicativeDepth follows PolyModulusDegree
        const bool f64 = bd.data_type == Float64;
        // concrete sample counts
        uint64_t s0 = 1, s1 = 1, batch = 1;
        if (bd.category == Offline) {
            if (is_logreg(bd.workload)) { batch = o.batch; bd.cat_params.offline.data_count[2] = batch; }
            else { s0 = o.samples[0]; s1 = o.samples[1]; bd.cat_params.offline.data_count[0] = s0; bd.cat_params.offline.data_count[1] = s1; }
        }
        printf("[ Info    ] %2lu: %s:", bi, title.str().c_str());
        for (auto &p : wp) printf(" %s=%lu", p.name, p.u_param);
        printf("\n");
        fflush(stdout);
        ++ran;
        Handle bench{};
        if (p_createBenchmark(engine, descs[bi], &wps, &bench) || p_initBenchmark(bench, &bd)) {
            printf("[ Error   ] createBenchmark: %s\n", lastError().c_str());
            ++failed;
            continue;
        }
        // ---- inputs + ground truth
        std::uniform_real_distribution<double> ud(is_logreg(bd.workload) ? -0.5 : -1.0, is_logreg(bd.workload) ? 0.5 : 1.0);
        std::uniform_int_distribution<int64_t> id(-10, 10);
        auto fill = [&](std::vector<unsigned char> &buf, size_t n) {
            buf.resize(n * 8);
            for (size_t i = 0; i < n; ++i) {
                if (f64) as<double>(buf)[i] = ud(rng);
                else as<int64_t>(buf)[i] = id(rng);
            }
        };
        auto val = [&](std::vector<unsigned char> &buf, size_t i) -> double { return f64 ? as<double>(buf)[i] : (double)as<int64_t>(buf)[i]; };
        NativeData in, out;
        std::vector<std::vector<double>> truth;   // [result buffer][value]
        std::vector<ParameterIndexer> idx;
        uint64_t results = 1;
        if (bd.workload == EltwiseAdd || bd.workload == EltwiseMultiply || bd.workload == DotProduct) {
            const uint64_t n = wp[0].u_param;
            in.bytes.resize(2);
            in.bytes[0].resize(s0);
            in.bytes[1].resize(s1);
            for (auto &b : in.bytes[0]) fill(b, n);
            for (auto &b : in.bytes[1]) fill(b, n);
            // ParameterIndexer sub-ranges (R/src/benchmarks/ckks/seal_ckks_element_wise_benchmark.cpp:334-336): results are
            // ordered (i - v0) * b1 + (j - v1)
            const uint64_t v0 = o.have_sub ? o.sub[0] : 0, b0 = o.have_sub ? o.sub[1] : s0, v1 = o.have_sub ? o.sub[2] : 0, b1 = o.have_sub ? o.sub[3] : s1;
            results = b0 * b1;
            for (uint64_t i = v0; i < v0 + b0 && i < s0; ++i)
                for (uint64_t j = v1; j < v1 + b1 && j < s1; ++j) {
                    std::vector<double> t;
                    if (bd.workload == DotProduct) {
                        double acc = 0;
                        for (uint64_t k = 0; k < n; ++k) acc += val(in.bytes[0][i], k) * val(in.bytes[1][j], k);
                        t.push_back(acc);
                    } else
                        for (uint64_t k = 0; k < n; ++k)
                            t.push_back(bd.workload == EltwiseAdd ? val(in.bytes[0][i], k) + val(in.bytes[1][j], k) : val(in.bytes[0][i], k) * val(in.bytes[1][j], k));
                    truth.push_back(t);
                }
            idx = { { v0, b0 }, { v1, b1 } };
        } else if (is_mat) {
            const uint64_t r0 = wp[0].u_param, c0 = wp[1].u_param, c1 = wp[2].u_param;
            in.bytes.resize(2);
            in.bytes[0].resize(1);
            in.bytes[1].resize(1);
            fill(in.bytes[0][0], r0 * c0);
            fill(in.bytes[1][0], c0 * c1);
            std::vector<double> t(r0 * c1, 0.0);
            for (uint64_t i = 0; i < r0; ++i)
                for (uint64_t j = 0; j < c1; ++j)
                    for (uint64_t k = 0; k < c0; ++k) t[i * c1 + j] += val(in.bytes[0][0], i * c0 + k) * val(in.bytes[1][0], k * c1 + j);
            truth.push_back(t);
            results = 1;
            idx = { { 0, 1 }, { 0, 1 } };
            if (o.have_sub) idx = { { o.sub[0], o.sub[1] }, { o.sub[2], o.sub[3] } };   // only to see it rejected (--expect-operate-error)
        } else {   // logistic regression
            const uint64_t n = wp[0].u_param;
            in.bytes.resize(3);
            in.bytes[0].resize(1);
            in.bytes[1].resize(1);
            in.bytes[2].resize(batch);
            fill(in.bytes[0][0], n);
            fill(in.bytes[1][0], 1);
            for (auto &b : in.bytes[2]) fill(b, n);
            for (uint64_t s = 0; s < batch; ++s) {
                double x = val(in.bytes[1][0], 0);
                for (uint64_t k = 0; k < n; ++k) x += val(in.bytes[0][0], k) * val(in.bytes[2][s], k);
                truth.push_back({ sigmoid_poly(bd.workload, x) });
            }
            results = batch;
            idx = { { 0, 1 }, { 0, 1 }, { 0, batch } };
            if (o.have_sub) idx[2] = { o.sub[0], o.sub[1] };   // only to see it rejected (--expect-operate-error)
        }
        in.build();
        out.bytes.resize(1);
        out.bytes[0].resize(truth.size());
        for (size_t r = 0; r < truth.size(); ++r) out.bytes[0][r].assign(truth[r].size() * 8, 0);
        out.build();

        // ---- HEBench flow
        auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
            return std::chrono::duration<double, std::milli>(b - a).count();
        };
        Handle h_enc{}, h_cipher{}, h_remote{}, h_result{}, h_local{}, h_plain{};
        bool ok = true, rejected = false;
        double t_load = 0, t_op = 0, t_store = 0, t_encode = 0, t_encrypt = 0, t_decrypt = 0, t_decode = 0, t_warm = 0;
        std::vector<double> t_ops;
        do {
            auto e0 = std::chrono::steady_clock::now();
            if (p_encode(bench, &in.coll, &h_enc)) { ok = false; break; }
            auto e1 = std::chrono::steady_clock::now();
            if (p_encrypt(bench, h_enc, &h_cipher)) { ok = false; break; }
            auto t0 = std::chrono::steady_clock::now();
            t_encode  = ms(e0, e1);
            t_encrypt = ms(e1, t0);
            if (p_load(bench, &h_cipher, 1, &h_remote)) { ok = false; break; }
            auto t1 = std::chrono::steady_clock::now();
            t_load  = ms(t0, t1);
            if (const int rc = p_operate(bench, h_remote, idx.data(), idx.size(), &h_result)) {   // warm-up
                if (o.expect_operate_error) {
                    printf("[ Info    ]     operate() rejected the indexers as expected: code %d, %s\n", rc, lastError().c_str());
                    rejected = true;
                }
                ok = false;
                break;
            }
            t_warm = ms(t1, std::chrono::steady_clock::now());
            for (uint64_t it = 0; it < o.iterations; ++it) {
                p_destroyHandle(h_result);
                auto a = std::chrono::steady_clock::now();
                if (p_operate(bench, h_remote, idx.data(), idx.size(), &h_result)) { ok = false; break; }
                t_ops.push_back(ms(a, std::chrono::steady_clock::now()));
                t_op += t_ops.back();
            }
            for (uint64_t it = 0; ok && it < o.extra_operate; ++it) {   // untimed: the backend's profiled calls
                p_destroyHandle(h_result);
                if (p_operate(bench, h_remote, idx.data(), idx.size(), &h_result)) ok = false;
            }
            if (!ok) break;
            t_op /= (double)std::max<uint64_t>(1, o.iterations);
            auto t2 = std::chrono::steady_clock::now();
            if (p_store(bench, h_result, &h_local, 1)) { ok = false; break; }
            auto t3 = std::chrono::steady_clock::now();
            t_store = ms(t2, t3);
            if (p_decrypt(bench, h_local, &h_plain)) { ok = false; break; }
            auto t4 = std::chrono::steady_clock::now();
            if (p_decode(bench, h_plain, &out.coll)) { ok = false; break; }
            t_decrypt = ms(t3, t4);
            t_decode  = ms(t4, std::chrono::steady_clock::now());
        } while (false);
        uint64_t bad = 0;
        if (o.expect_operate_error) {
            if (!rejected) {
                printf("[ Error   ] operate() accepted indexers it had to reject\n");
                bad = 1;
            }
        } else if (!ok) {
            printf("[ Error   ] %s\n", lastError().c_str());
            bad = 1;
        } else {
            for (size_t r = 0; r < truth.size(); ++r)
                for (size_t k = 0; k < truth[r].size(); ++k) {
                    const double got = val(out.bytes[0][r], k);
                    if (!almostEqual(got, truth[r][k])) {
                        if (bad < 5) printf("[ Warning ] result %zu value %zu: got %.9g expected %.9g\n", r, k, got, truth[r][k]);
                        ++bad;
                    }
                }
        }
        const double sps = t_op > 0 ? results / (t_op / 1e3) : 0;
        printf("[ Info    ]     load %.2f ms, operate %.3f ms (%lu results, %.1f samples/s), store %.2f ms -> %s\n", t_load, t_op, results, sps, t_store,
               bad ? "FAILED" : "ok");
        if (csv.is_open()) {
            csv << bi << ',' << workloadName(bd.workload) << ',' << scheme << ',' << (bd.category == Latency ? "Latency" : "Offline") << ',' << bd.other << ',';
            for (auto &p : wp) csv << p.name << '=' << p.u_param << ' ';
            csv << ',' << results << ',' << t_op << ',' << sps << ',' << t_load << ',' << t_store << ',' << (ok ? 1 : 0) << ',' << bad << '\n';
        }
        if (json.is_open()) {
            json << "{\"index\": " << bi << ", \"title\": \"" << title.str() << "\", \"params\": {";
            for (size_t k = 0; k < wp.size(); ++k) json << (k ? ", " : "") << "\"" << wp[k].name << "\": " << wp[k].u_param;
            json << "}, \"results_per_operate\": " << results << ", \"validated\": " << (bad ? "false" : "true") << ", \"failed_values\": " << bad
                 << ", \"encode_ms\": " << t_encode << ", \"encrypt_ms\": " << t_encrypt << ", \"load_ms\": " << t_load << ", \"warmup_operate_ms\": " << t_warm
                 << ", \"operate_ms\": " << t_op << ", \"operate_ms_all\": [";
            for (size_t k = 0; k < t_ops.size(); ++k) json << (k ? ", " : "") << t_ops[k];
            json << "], \"store_ms\": " << t_store << ", \"decrypt_ms\": " << t_decrypt << ", \"decode_ms\": " << t_decode << "}\n";
            json.flush();
        }
        if (bad) ++failed;
        for (Handle h : { h_plain, h_local, h_result, h_remote, h_cipher, h_enc, bench })
            if (h.p) p_destroyHandle(h);
    }
    for (Handle h : descs) p_destroyHandle(h);
    p_destroyHandle(engine);
    if (!o.list) {
        printf("[ Info    ] Total: %lu\n", ran);
        printf("[ Info    ] Failed: %lu\n", failed);
    }
    return failed ? 1 : 0;
}
