// dvg_prof.cu -- measurement support: per-kernel CUDA-event timing (dvg_profile_*) and the
// FP32 / FP64 FMA peak probes bench.py uses as roofline denominators (MEASURED_PEAKS.json only
// carries the HBM and bf16 tensor peaks, and this path is bound by the CUDA-core pipes).
#include "dvg_internal.h"

#include <map>
#include <string>
#include <vector>

namespace dvg {

bool g_profile_on = false;

namespace {
struct Span { const char *name; cudaEvent_t a, b; };
std::vector<Span> g_spans;
std::vector<cudaEvent_t> g_pool;
cudaEvent_t g_open = nullptr;
const char *g_open_name = nullptr;

cudaEvent_t get_event() {
    if (!g_pool.empty()) { cudaEvent_t e = g_pool.back(); g_pool.pop_back(); return e; }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
}
}  // namespace

void prof_begin(const char *name, cudaStream_t st) {
    g_open = get_event();
    g_open_name = name;
    cudaEventRecord(g_open, st);
}
void prof_end(cudaStream_t st) {
    cudaEvent_t b = get_event();
    cudaEventRecord(b, st);
    g_spans.push_back(Span{g_open_name, g_open, b});
}

// Synchronises, aggregates by kernel name and clears.  Text: "name,launches,total_ms\n" per line.
int prof_report(char *buf, long long cap) {
    std::map<std::string, std::pair<int, double>> agg;
    for (const Span &s : g_spans) {
        cudaEventSynchronize(s.b);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, s.a, s.b);
        auto &e = agg[s.name];
        e.first++; e.second += ms;
        g_pool.push_back(s.a); g_pool.push_back(s.b);
    }
    g_spans.clear();
    std::string out;
    for (auto &kv : agg) {
        char line[256];
        snprintf(line, sizeof line, "%s,%d,%.6f\n", kv.first.c_str(), kv.second.first, kv.second.second);
        out += line;
    }
    if ((long long)out.size() + 1 > cap) return -1;
    memcpy(buf, out.c_str(), out.size() + 1);
    return (int)out.size();
}

// Peak probes: 8 independent FMA chains per thread, fully unrolled; 2 flops per FMA.
template <typename T>
__global__ void __launch_bounds__(256) k_peak_probe(T *out, int iters, T seed) {
    T a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const T m = (T)0.999, c = (T)0.001;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 16; k++) {
            a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
            a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
        }
    }
    T s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (s == (T)-12345) out[0] = s;
}

// which: 0 = FP32, 1 = FP64.  Returns flops executed through *flops.
double peak_probe_flops(int iters, int blocks) { return (double)blocks * 256 * (double)iters * 16 * 8 * 2; }

void launch_peak_probe(int which, float *out, int iters, cudaStream_t st) {
    const int blocks = 148 * 8;
    if (which == 0) k_peak_probe<float><<<blocks, 256, 0, st>>>(out, iters, 1.0f);
    else k_peak_probe<double><<<blocks, 256, 0, st>>>((double *)out, iters, 1.0);
}

}  // namespace dvg
