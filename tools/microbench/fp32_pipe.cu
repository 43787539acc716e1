// Issue-rate microbenchmark for the fp32 pipe on sm_100a: scalar FFMA / FADD / FMUL against the
// packed FFMA2 / FADD2 / FMUL2 forms.  Prints warp-instructions per clock per SM and lane-ops per clock per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp32_pipe fp32_pipe.cu && ./fp32_pipe
#include <cuda_runtime.h>
#include <cstdio>

constexpr int kChains = 16;   // independent dependency chains per thread
constexpr int kIters = 4096;

template <int MODE>
__global__ void __launch_bounds__(512, 1) bench(float* out, float x, float y, long long* cycles) {
    float2 a[kChains], c[kChains];
    const float2 b = make_float2(x, y), d = make_float2(y, x);
#pragma unroll
    for (int i = 0; i < kChains; ++i) { a[i] = make_float2(x + i, y - i); c[i] = make_float2(y + i, x - i); }
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
        for (int i = 0; i < kChains; ++i) {
            if (MODE == 0) { a[i].x = fmaf(a[i].x, b.x, c[i].x); a[i].y = fmaf(a[i].y, b.y, c[i].y); }       // 2 FFMA
            if (MODE == 1) { a[i].x = a[i].x + c[i].x; a[i].y = a[i].y + c[i].y; }                              // 2 FADD
            if (MODE == 2) { a[i].x = a[i].x * b.x; a[i].y = a[i].y * d.y; }                                    // 2 FMUL
            if (MODE == 3) { a[i] = __ffma2_rn(a[i], b, c[i]); }                                                // 1 FFMA2
            if (MODE == 4) { a[i] = __fadd2_rn(a[i], c[i]); }                                                   // 1 FADD2
            if (MODE == 5) { a[i] = __fmul2_rn(a[i], b); }                                                      // 1 FMUL2
            if (MODE == 6) { a[i].x = fmaf(a[i].x, b.x, c[i].x); a[i].y = a[i].y + c[i].y; }                   // FFMA + FADD
            if (MODE == 7) { a[i] = __ffma2_rn(a[i], b, c[i]); c[i].x = c[i].x + d.x; }                         // FFMA2 + FADD
            if (MODE == 8) { a[i] = __ffma2_rn(a[i], b, c[i]); c[i] = __fadd2_rn(c[i], d); }                    // FFMA2 + FADD2
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kChains; ++i) s += a[i].x + a[i].y + c[i].x + c[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int instr_per_slot, int laneops_per_slot, int threads) {
    float* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    bench<MODE><<<148, threads>>>(out, 1.0001f, 0.9999f, cyc);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    bench<MODE><<<148, threads>>>(out, 1.0001f, 0.9999f, cyc);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
    const double warps = threads / 32.0;
    const double winstr = warps * kIters * (double)kChains * instr_per_slot;
    printf("%-16s threads %4d  %8.0f clk  %6.3f warp-instr/clk/SM  %7.2f lane-ops/clk/SM  (%.3f ms)\n", name, threads, c,
           winstr / c, warps * 32 * kIters * (double)kChains * laneops_per_slot / c, ms);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    for (int threads : {128, 256, 512}) {
        run<0>("FFMA x2", 2, 2, threads);
        run<1>("FADD x2", 2, 2, threads);
        run<2>("FMUL x2", 2, 2, threads);
        run<3>("FFMA2", 1, 2, threads);
        run<4>("FADD2", 1, 2, threads);
        run<5>("FMUL2", 1, 2, threads);
        run<6>("FFMA+FADD", 2, 2, threads);
        run<7>("FFMA2+FADD", 2, 3, threads);
        run<8>("FFMA2+FADD2", 2, 4, threads);
    }
    return 0;
}
