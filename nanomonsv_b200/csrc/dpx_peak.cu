// dpx_peak.cu -- measures the sustained issue rate of the integer/DPX instructions the
// Smith-Waterman kernel is built from (register-only loops, no memory traffic).
// The result is the denominator of the integer roofline quoted in DESIGN.md / bench.py.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o dpx_peak dpx_peak.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

constexpr int CHAINS = 8;
constexpr int ITERS = 4096;

enum Op { VIADDMAX_S16X2 = 0, VIADDMAX_S16X2_RELU, VIMAX3_S16X2_RELU, VIMAX_S16X2, VIBMAX_S16X2,
          VIADDMAX_S32, VIMAX3_S32_RELU, IADD, IMAD, SW_CELL, SW_CELL_IMAD, NOPS };
static const char *op_names[NOPS] = {"viaddmax_s16x2", "viaddmax_s16x2_relu", "vimax3_s16x2_relu", "vimax_s16x2",
  "vibmax_s16x2", "viaddmax_s32", "vimax3_s32_relu", "iadd", "imad", "sw_cell(6 dpx/2 cells)", "sw_cell(5 dpx+1 iadd)"};
// lane-instructions issued per loop trip and chain
static const int op_insts[NOPS] = {1, 1, 1, 1, 1, 1, 1, 1, 1, 6, 6};

template <int OP>
__global__ void __launch_bounds__(256) peak_kernel(unsigned *out, unsigned seed, unsigned k1, unsigned k2)
{
    unsigned a[CHAINS], b[CHAINS], c[CHAINS], d[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) {
        a[i] = seed + threadIdx.x * 3 + i; b[i] = seed ^ (i * 77u); c[i] = threadIdx.x + i; d[i] = i;
    }
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) {
            if (OP == VIADDMAX_S16X2)       a[i] = __viaddmax_s16x2(a[i], k1, b[i]);
            else if (OP == VIADDMAX_S16X2_RELU) a[i] = __viaddmax_s16x2_relu(a[i], k1, b[i]);
            else if (OP == VIMAX3_S16X2_RELU)   a[i] = __vimax3_s16x2_relu(a[i], b[i], k1);
            else if (OP == VIMAX_S16X2)     a[i] = __vimax_s16x2_relu(a[i], b[i] ^ a[i]);
            else if (OP == VIBMAX_S16X2)  { bool p, q; a[i] = __vibmax_s16x2(a[i], b[i] + a[i], &p, &q); }
            else if (OP == VIADDMAX_S32)    a[i] = __viaddmax_s32((int)a[i], (int)k1, (int)b[i]);
            else if (OP == VIMAX3_S32_RELU) a[i] = __vimax3_s32_relu((int)a[i], (int)b[i], (int)k1);
            else if (OP == IADD)            a[i] = a[i] + (b[i] ^ k1);   // LOP3 + IADD3 may fuse; see SASS
            else if (OP == IMAD)            a[i] = a[i] * k1 + b[i];
            else if (OP == SW_CELL) {
                // a = H(diag), b = E', c = F', d = best ; k1 = -open packed, k2 = -ext packed
                unsigned t  = __viaddmax_s16x2(a[i], seed, 0x80008000u);      // Hdiag + S
                unsigned x  = __viaddmax_s16x2(b[i], k1, t);                // max(E'-open, t)
                unsigned h  = __viaddmax_s16x2_relu(c[i], k1, x);           // max(F'-open, x, 0)
                b[i] = __viaddmax_s16x2(b[i], k2, h);
                c[i] = __viaddmax_s16x2(c[i], k2, h);
                d[i] = __vimax_s16x2_relu(d[i], h);
                a[i] = h;
            } else if (OP == SW_CELL_IMAD) {
                unsigned t  = a[i] + seed;                                   // plain 32-bit add (may go to IMAD/IADD3)
                unsigned x  = __viaddmax_s16x2(b[i], k1, t);
                unsigned h  = __viaddmax_s16x2_relu(c[i], k1, x);
                b[i] = __viaddmax_s16x2(b[i], k2, h);
                c[i] = __viaddmax_s16x2(c[i], k2, h);
                d[i] = __vimax_s16x2_relu(d[i], h);
                a[i] = h;
            }
        }
    }
    unsigned r = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) r ^= a[i] ^ b[i] ^ c[i] ^ d[i];
    if (r == 0x12345678u) out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int OP>
static int run(int sms, double clock_ghz_nominal, unsigned *d_out, FILE *json, bool first)
{
    const int blocks = sms * 8, threads = 256;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int w = 0; w < 3; ++w) peak_kernel<OP><<<blocks, threads>>>(d_out, 12345u, 0xfffdfffdu, 0xffffffffu);
    CK(cudaDeviceSynchronize());
    float best_ms = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        CK(cudaEventRecord(e0));
        peak_kernel<OP><<<blocks, threads>>>(d_out, 12345u, 0xfffdfffdu, 0xffffffffu);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best_ms) best_ms = ms;
    }
    double lane_ops = (double)blocks * threads * ITERS * CHAINS * op_insts[OP];
    double tops = lane_ops / (best_ms * 1e-3) / 1e12;
    double per_clk_sm = lane_ops / (best_ms * 1e-3) / (clock_ghz_nominal * 1e9) / sms;
    printf("%-26s %8.3f ms  %8.3f T lane-inst/s  %7.2f lane-inst/clk/SM (at %.3f GHz nominal)\n",
           op_names[OP], best_ms, tops, per_clk_sm, clock_ghz_nominal);
    fprintf(json, "%s\"%s\": {\"ms\": %.4f, \"tera_lane_inst_per_s\": %.4f, \"lane_inst_per_clk_per_sm_at_max_clock\": %.3f}",
            first ? "" : ", ", op_names[OP], best_ms, tops, per_clk_sm);
    return 0;
}

int main(int argc, char **argv)
{
    const char *json_path = argc > 1 ? argv[1] : "dpx_peak.json";
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    int sms = prop.multiProcessorCount;
    int clk_khz = 0; CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
    double ghz = clk_khz * 1e-6;
    printf("device %s, %d SMs, max SM clock %.3f GHz\n", prop.name, sms, ghz);
    unsigned *d_out; CK(cudaMalloc(&d_out, (size_t)sms * 8 * 256 * sizeof(unsigned)));
    FILE *json = fopen(json_path, "w");
    if (!json) { fprintf(stderr, "cannot open %s\n", json_path); return 1; }
    fprintf(json, "{\"device\": \"%s\", \"sms\": %d, \"sm_max_ghz\": %.4f, \"ops\": {", prop.name, sms, ghz);
    int rc = 0;
    rc |= run<VIADDMAX_S16X2>(sms, ghz, d_out, json, true);
    rc |= run<VIADDMAX_S16X2_RELU>(sms, ghz, d_out, json, false);
    rc |= run<VIMAX3_S16X2_RELU>(sms, ghz, d_out, json, false);
    rc |= run<VIMAX_S16X2>(sms, ghz, d_out, json, false);
    rc |= run<VIBMAX_S16X2>(sms, ghz, d_out, json, false);
    rc |= run<VIADDMAX_S32>(sms, ghz, d_out, json, false);
    rc |= run<VIMAX3_S32_RELU>(sms, ghz, d_out, json, false);
    rc |= run<IADD>(sms, ghz, d_out, json, false);
    rc |= run<IMAD>(sms, ghz, d_out, json, false);
    rc |= run<SW_CELL>(sms, ghz, d_out, json, false);
    rc |= run<SW_CELL_IMAD>(sms, ghz, d_out, json, false);
    fprintf(json, "}}\n");
    fclose(json);
    cudaFree(d_out);
    return rc;
}
