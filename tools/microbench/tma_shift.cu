// Can TMA deliver a 2048-float slice that starts at ANY float offset (not a multiple of 16 bytes) into a 128-byte
// aligned shared-memory buffer?  cp.async.bulk (1-D bulk copy) needs 16-byte aligned source, destination and size, so
// the tile mixer's consumers read their slices with scalar LDS.  A tensor-map copy (cp.async.bulk.tensor) takes
// ELEMENT coordinates: this probe checks, for every shift 0..7,
//   (a) a 3-D map over the flat bank with overlapping strides {4 B, 16 B, 1024 B}, box {256, 1, 8}: one copy per slice
//   (b) a 1-D map, box {256}: eight copies per slice
// against the source, and measures both against the 1-D bulk copy on the mixer's access pattern (random offsets in
// an L2-resident region, 8 KB slices, `in_flight` slices per warp).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_shift tma_shift.cu -lcuda && ./tma_shift
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int n) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n));
}
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) {
    asm volatile("{ .reg .b64 t; mbarrier.arrive.expect_tx.shared::cta.b64 t, [%0], %1; }" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t ph) {
    asm volatile("{\n\t.reg .pred q;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%0], %1;\n\t@q bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(
                     smem_u32(b)), "r"(ph) : "memory");
}
__device__ __forceinline__ void tma3(void* dst, const CUtensorMap* m, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                     smem_u32(dst)), "l"(m), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma1(void* dst, const CUtensorMap* m, int c0, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.1d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2}], [%3];" ::"r"(
                     smem_u32(dst)), "l"(m), "r"(c0), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

constexpr int kSlice = 2048;          // floats
constexpr int kStagesMax = 8;

// mode 0: bulk copy (offset rounded down to 4 floats), 1: 3-D overlapping map, 2: 1-D map x 8
// One warp per CTA; lane l < in_flight drives ring slot l.  check != 0: compare every element with the source.
__global__ void __launch_bounds__(32) probe(const float* __restrict__ src, const __grid_constant__ CUtensorMap m3,
                                            const __grid_constant__ CUtensorMap m1, const unsigned* __restrict__ offs, int n,
                                            int passes, int in_flight, int mode, int check, unsigned long long* bad) {
    extern __shared__ __align__(128) unsigned char smem[];
    float* buf = reinterpret_cast<float*>(smem);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + kStagesMax * kSlice * 4);
    const int lane = threadIdx.x;
    if (lane < kStagesMax) mbar_init(bar + lane, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    if (lane < in_flight) {
        uint32_t phase = 0;
        float* dst = buf + lane * kSlice;
        for (int it = 0; it < passes; ++it) {
            for (int s = blockIdx.x * in_flight + lane; s < n; s += gridDim.x * in_flight) {
                unsigned off = __ldg(offs + s);
                mbar_expect(bar + lane, kSlice * 4);
                if (mode == 0) {
                    off &= ~3u;
                    bulk(dst, src + off, kSlice * 4, bar + lane);
                } else if (mode == 1) {
                    tma3(dst, &m3, (int)(off & 3u), (int)((off >> 2) & 63u), (int)(off >> 8), bar + lane);
                } else {
#pragma unroll
                    for (int k = 0; k < 8; ++k) tma1(dst + 256 * k, &m1, (int)off + 256 * k, bar + lane);
                }
                mbar_wait(bar + lane, phase);
                phase ^= 1u;
                if (check) {
                    unsigned long long nb = 0;
                    for (int i = 0; i < kSlice; ++i) nb += dst[i] != src[off + i];
                    if (nb) atomicAdd(bad, nb);
                }
            }
        }
    }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// usage: tma_shift <variant>   (one process per variant: a faulting TMA poisons the context)
//   0: 3-D overlapping {259, 64, N/256} strides {16, 1024}      1: the same with dim0 = 260      2: dim0 = 512
//   3: 3-D overlapping, ALIGNED offsets only (multiples of 4)    4: 1-D map x 8 (strides pointer non-null)
//   5: 1-D map x 8, aligned offsets only                         6: 2-D non-overlapping {256, N/256} box {256, 8}, aligned to 256
int main(int argc, char** argv) {
    const int variant = argc > 1 ? atoi(argv[1]) : 0;
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const size_t n_floats = 24u << 20;   // 96 MB: L2-resident
    float* buf;
    cudaMalloc(&buf, n_floats * 4);
    std::vector<float> h(n_floats);
    for (size_t i = 0; i < n_floats; ++i) h[i] = (float)(i % 1000003u) + 0.25f;
    cudaMemcpy(buf, h.data(), n_floats * 4, cudaMemcpyHostToDevice);

    EncodeFn encode = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qr) != cudaSuccess || !encode) {
        printf("cuTensorMapEncodeTiled not available\n");
        return 1;
    }
    CUtensorMap m3, m1;
    memset(&m3, 0, sizeof(m3)); memset(&m1, 0, sizeof(m1));
    int mode = 1;
    unsigned align_mask = ~0u;
    CUresult r = CUDA_SUCCESS;
    if (variant <= 3) {
        const cuuint64_t d0 = variant == 1 ? 260 : variant == 2 ? 512 : 259;
        const cuuint64_t dims[3] = {d0, 64, n_floats / 256};
        const cuuint64_t strides[2] = {16, 1024};
        const cuuint32_t box[3] = {256, 1, 8}, es[3] = {1, 1, 1};
        r = encode(&m3, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (variant == 3) align_mask = ~3u;
    } else if (variant <= 5) {
        const cuuint64_t dims[1] = {n_floats};
        const cuuint64_t strides[1] = {0};
        const cuuint32_t box[1] = {256}, es[1] = {1};
        r = encode(&m1, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 1, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        mode = 2;
        if (variant == 5) align_mask = ~3u;
    } else {
        const cuuint64_t dims[3] = {256, 1, n_floats / 256};   // as a 3-D map so that the same kernel path serves
        const cuuint64_t strides[2] = {1024, 1024};
        const cuuint32_t box[3] = {256, 1, 8}, es[3] = {1, 1, 1};
        r = encode(&m3, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        align_mask = ~255u;
    }
    printf("variant %d: encode CUresult %d\n", variant, (int)r);
    if (r != CUDA_SUCCESS) return 0;
    const size_t smem = kStagesMax * kSlice * 4 + 128;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    unsigned long long* bad;
    cudaMalloc(&bad, 8);
    {
        std::vector<unsigned> offs;
        for (unsigned base : {0u, 4096u, 1000000u - 64u, 12345600u})
            for (unsigned s = 0; s < 8; ++s) offs.push_back((base + s) & align_mask);
        for (unsigned s = 0; s < 8; ++s) offs.push_back(((unsigned)n_floats - kSlice - s) & align_mask);
        unsigned* d;
        cudaMalloc(&d, offs.size() * 4);
        cudaMemcpy(d, offs.data(), offs.size() * 4, cudaMemcpyHostToDevice);
        cudaMemset(bad, 0, 8);
        probe<<<4, 32, smem>>>(buf, m3, m1, d, (int)offs.size(), 1, 4, mode, 1, bad);
        const cudaError_t e = cudaDeviceSynchronize();
        unsigned long long nb = 0;
        cudaMemcpy(&nb, bad, 8, cudaMemcpyDeviceToHost);
        printf("variant %d correctness: %llu mismatching floats over %zu slices (%s)\n", variant, nb, offs.size(),
               cudaGetErrorString(e));
        if (e != cudaSuccess) return 0;
        cudaFree(d);
    }
    // ---- throughput on the mixer's pattern, against the 16-byte aligned bulk copy
    const int n = 200000;
    std::vector<unsigned> offs(n);
    unsigned long long st = 88172645463325252ull;
    for (int i = 0; i < n; ++i) {
        st ^= st << 13; st ^= st >> 7; st ^= st << 17;
        offs[i] = (unsigned)(st % (n_floats - kSlice)) & align_mask;
    }
    unsigned* d;
    cudaMalloc(&d, n * 4);
    cudaMemcpy(d, offs.data(), n * 4, cudaMemcpyHostToDevice);
    const int configs[][2] = {{3, 8}, {3, 4}, {2, 8}, {1, 8}};
    for (int md : {0, mode})
        for (auto& c : configs) {
            cudaEvent_t a, b;
            cudaEventCreate(&a); cudaEventCreate(&b);
            const int passes = 4;
            for (int rep = 0; rep < 2; ++rep) {
                cudaEventRecord(a);
                probe<<<sms * c[0], 32, smem>>>(buf, m3, m1, d, n, passes, c[1], md, 0, bad);
                cudaEventRecord(b);
                cudaEventSynchronize(b);
            }
            float ms = 0.f;
            cudaEventElapsedTime(&ms, a, b);
            printf("variant %d %-12s %d CTAs/SM x %d in flight: %8.3f ms  %8.1f GB/s  (%s)\n", variant,
                   md == 0 ? "bulk aligned" : "tensor map", c[0], c[1], ms, (double)n * kSlice * 4 * passes / ms / 1e6,
                   cudaGetErrorString(cudaGetLastError()));
        }
    return 0;
}
