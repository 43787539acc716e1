// L2 -> SM read bandwidth on sm_100a: every SM streams an L2-resident buffer with (a) LDG.128 and (b) TMA bulk copies
// into shared memory (the tile mixer's access pattern: 8 KB slices).  Prints GB/s for several buffer sizes; the sizes
// below the 126 MB L2 give the L2 ceiling the render kernels are measured against, the large one the HBM rate.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l2_bw l2_bw.cu && ./l2_bw
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

__global__ void __launch_bounds__(256) read_ldg(const float4* __restrict__ p, size_t n4, int passes, float* sink) {
    float acc = 0.f;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int it = 0; it < passes; ++it) {
        for (size_t base = (size_t)blockIdx.x * blockDim.x + threadIdx.x; base < n4; base += 4 * stride) {
            float4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const size_t i = base + u * stride;
                v[u] = i < n4 ? __ldcg(p + i) : make_float4(0.f, 0.f, 0.f, 0.f);   // L2 only: no L1 reuse
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) acc += v[u].x + v[u].y + v[u].z + v[u].w;
        }
    }
    if (acc == 123.456f) *sink = acc;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// one warp per CTA, lane l drives ring slot l: kStages 8 KB TMA bulk copies in flight per CTA, three CTAs per SM;
// nothing reads the data (pure transport)
constexpr int kStages = 8, kSlice = 8192;
__global__ void __launch_bounds__(32) read_tma(const char* __restrict__ p, size_t bytes, int passes) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + kStages * kSlice);
    const int lane = threadIdx.x;
    if (lane < kStages) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar + lane)));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    const size_t n_slices = bytes / kSlice;
    // lane l owns ring slot l: issue a copy, wait for it, next slice - kStages copies in flight per CTA
    if (lane < kStages) {
        uint32_t phase = 0;
        for (int it = 0; it < passes; ++it) {
            for (size_t s = (size_t)blockIdx.x * kStages + lane; s < n_slices; s += (size_t)gridDim.x * kStages) {
                asm volatile("{ .reg .b64 t; mbarrier.arrive.expect_tx.shared::cta.b64 t, [%0], %1; }" ::"r"(smem_u32(bar + lane)),
                             "r"(kSlice)
                             : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                 smem_u32(smem + lane * kSlice)),
                             "l"(p + s * kSlice), "r"(kSlice), "r"(smem_u32(bar + lane))
                             : "memory");
                asm volatile(
                    "{\n\t.reg .pred q;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%0], %1;\n\t@q bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(
                        smem_u32(bar + lane)),
                    "r"(phase)
                    : "memory");
                phase ^= 1u;
            }
        }
    }
}

// The tile mixer's pattern: slices of arbitrary length (16-byte granules) at arbitrary 16-byte offsets, `in_flight`
// copies per CTA (ring slots), from a list.  desc[i] = {offset_bytes, bytes}.
__global__ void __launch_bounds__(32) read_tma_list(const char* __restrict__ p, const uint2* __restrict__ desc, int n_desc,
                                                    int passes, int in_flight) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + kStages * kSlice);
    const int lane = threadIdx.x;
    if (lane < kStages) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar + lane)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    if (lane < in_flight) {
        uint32_t phase = 0;
        for (int it = 0; it < passes; ++it) {
            for (int s = blockIdx.x * in_flight + lane; s < n_desc; s += gridDim.x * in_flight) {
                const uint2 d = __ldg(desc + s);
                asm volatile("{ .reg .b64 t; mbarrier.arrive.expect_tx.shared::cta.b64 t, [%0], %1; }" ::"r"(smem_u32(bar + lane)),
                             "r"(d.y)
                             : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                 smem_u32(smem + lane * kSlice)),
                             "l"(p + d.x), "r"(d.y), "r"(smem_u32(bar + lane))
                             : "memory");
                asm volatile(
                    "{\n\t.reg .pred q;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%0], %1;\n\t@q bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(
                        smem_u32(bar + lane)),
                    "r"(phase)
                    : "memory");
                phase ^= 1u;
            }
        }
    }
}

static void mixer_pattern(int sms) {
    // 96 MB region, 200 k slices: length uniform in [1 KB, 8 KB] (16-byte granules), offset any multiple of 16 bytes
    const size_t region = 96u << 20;
    const int n = 200000;
    uint2* h = (uint2*)malloc(n * sizeof(uint2));
    unsigned long long st = 88172645463325252ull, total = 0;
    for (int i = 0; i < n; ++i) {
        st ^= st << 13; st ^= st >> 7; st ^= st << 17;
        const unsigned len = (1024 + (unsigned)(st % 7168)) & ~15u;
        st ^= st << 13; st ^= st >> 7; st ^= st << 17;
        const unsigned off = (unsigned)(st % (region - 8192)) & ~15u;
        h[i] = make_uint2(off, len);
        total += len;
    }
    char* buf; uint2* d;
    cudaMalloc(&buf, region); cudaMemset(buf, 1, region);
    cudaMalloc(&d, n * sizeof(uint2)); cudaMemcpy(d, h, n * sizeof(uint2), cudaMemcpyHostToDevice);
    const size_t smem_tma = kStages * kSlice + 128;
    cudaFuncSetAttribute(read_tma_list, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_tma);
    const int configs[][2] = {{3, 8}, {3, 4}, {3, 2}, {2, 8}, {1, 8}, {3, 1}};   // {CTAs per SM, copies in flight per CTA}
    for (auto& c : configs) {
        cudaEvent_t a, b;
        cudaEventCreate(&a); cudaEventCreate(&b);
        const int passes = 4;
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(a);
            read_tma_list<<<sms * c[0], 32, smem_tma>>>(buf, d, n, passes, c[1]);
            cudaEventRecord(b);
            cudaEventSynchronize(b);
        }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, a, b);
        printf("mixer-like slices (1-8 KB, 16 B aligned): %d CTAs/SM x %d in flight = %2d per SM: %8.3f ms  %8.1f GB/s  (%s)\n", c[0],
               c[1], c[0] * c[1], ms, (double)total * passes / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
    }
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float* sink;
    cudaMalloc(&sink, 4);
    const size_t sizes[] = {16u << 20, 32u << 20, 64u << 20, 96u << 20, 2048ull << 20};
    const size_t smem_tma = kStages * kSlice + 128;
    cudaFuncSetAttribute(read_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_tma);
    for (size_t bytes : sizes) {
        char* buf;
        cudaMalloc(&buf, bytes);
        cudaMemset(buf, 1, bytes);
        const int passes = bytes > (512u << 20) ? 2 : 40;
        for (int mode = 0; mode < 2; ++mode) {
            cudaEvent_t a, b;
            cudaEventCreate(&a); cudaEventCreate(&b);
            for (int rep = 0; rep < 2; ++rep) {  // the first repetition warms L2
                cudaEventRecord(a);
                if (mode == 0) read_ldg<<<sms * 8, 256>>>((const float4*)buf, bytes / 16, passes, sink);
                else read_tma<<<sms * 3, 32, smem_tma>>>(buf, bytes, passes);
                cudaEventRecord(b);
                cudaEventSynchronize(b);
            }
            float ms = 0.f;
            cudaEventElapsedTime(&ms, a, b);
            printf("%-8s %6zu MB x %2d passes: %8.3f ms  %8.1f GB/s  (%s)\n", mode == 0 ? "LDG.128" : "TMA 8KB", bytes >> 20, passes, ms,
                   (double)bytes * passes / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
        }
        cudaFree(buf);
    }
    mixer_pattern(sms);
    return 0;
}
