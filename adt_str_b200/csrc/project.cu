// project_to_mel on the 5th-generation tensor cores (sm_100a): tcgen05.mma, accumulators in tensor memory.
//
// Replaces the first layer behind the log-mel (reference model.py:224-226, 249):
//     self.project_to_mel = nn.Linear(n_mels, d_query * nhead);   src_emb = self.project_to_mel(src_emb)
// which runs under bf16 autocast (configs/train/setting-1.yaml: bf16): the float32 log-mel (rows = B * T, 128) and the
// weight are cast to bf16, the product is accumulated in float32, the bias (bf16) is added and the result is bf16.
// It is the one dense contraction next to the path: (rows, 128) x (128, 768).  With K = 128 the
// GEMM is bound by HBM, not by the tensor cores - 512 B read and 1536 B written per row against 197 kflop - so the
// kernel is built around the bytes: the cast of the log-mel is fused in (the library path runs a separate cast kernel,
// another 768 B per row), the whole weight matrix stays in shared memory for the life of a CTA, and the 768-wide
// output row leaves through the epilogue once.
//
// One persistent CTA of 8 warps per SM.  Per tile of 128 rows:
//   A      a warp instruction fetches one row (512 contiguous bytes), one tile AHEAD, behind the epilogue of the current
//          one; each lane converts its four floats to bf16 and stores them into the K-major no-swizzle canonical
//          layout: 8-row x 16-byte core matrices, rows of a k-group contiguous
//          (offset(r, k) = (k / 8) * 2064 + r * 16 + (k % 8) * 2: LBO = 2064 B - 16 B of padding per k-group make the
//          warp's stores conflict-free - SBO = 128 B)
//   MMA    one thread issues tcgen05.mma.cta_group::1.kind::f16 (M = 128, N = 256, K = 16) eight times per 256-column
//          slab of the output, operands by shared-memory descriptors, D in tensor memory; three slabs for 768
//          columns alternate between two 256-column TMEM stages; tcgen05.commit arrives on the stage's mbarrier
//   D      all eight warps drain a stage: warp w reads the TMEM lanes of its quarter (w % 4) and its half of the
//          columns (w / 4) with tcgen05.ld.32x32b.x32 - one output row per thread, 32 columns per instruction - adds
//          the bias, rounds to bf16 and writes 64 contiguous bytes per instruction.  The next slab's MMAs run meanwhile.
#include <cuda_bf16.h>

#include <string.h>

#include <algorithm>
#include <vector>

#include "common.cuh"

struct adtfe_linear {
    int device = 0;
    int sm_count = 148;
    int32_t n_in = 128, n_out = 768;
    void* w_image = nullptr;   // bf16, canonical K-major layout of the whole (n_out, 128) weight: n_out * 256 bytes
    float* bias = nullptr;     // n_out floats, already rounded to bf16 (autocast casts the bias too)
    size_t smem_bytes = 0;
};

namespace adtfe {

constexpr int kPM = 128;          // rows per tile = UMMA M
constexpr int kPK = 128;          // n_mels
constexpr int kPN = 256;          // columns per MMA slab = UMMA N (the last slab may be narrower)
constexpr int kPThreads = 256;
constexpr int kALbo = kPM * 16 + 16;         // bytes between the k-groups of the A tile: 2048 + 16, so that the 32 lanes of
                                             // a warp (one row: 16 k-groups x 2 halves) store to 32 different banks
constexpr int kATileBytes = 16 * kALbo;      // 32.25 KB

__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    // cute::UMMA::SmemDescriptor: start address, leading / stride byte offsets (4 LSB dropped), version 1 (Blackwell),
    // base offset 0, layout type 0 = no swizzle
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3ffffu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

__host__ __device__ constexpr uint32_t umma_idesc_bf16(int m, int n) {
    // cute::UMMA::InstrDescriptor: D = F32 (1 << 4), A = B = BF16 (1 << 7, 1 << 10), both K-major, N >> 3 at bit 17,
    // M >> 4 at bit 24
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    const __nv_bfloat162 p = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&p);
}

__global__ void __launch_bounds__(kPThreads, 1) project_kernel(const float* __restrict__ x, int64_t n_rows,
                                                               const void* __restrict__ w_image,
                                                               const float* __restrict__ bias, int n_out,
                                                               __nv_bfloat16* __restrict__ out) {
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* s_w = smem;                                     // n_out * 256 bytes
    unsigned char* s_a = smem + (size_t)n_out * 256;               // kATileBytes
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_a + kATileBytes);   // [0]: weight copy, [1], [2]: TMEM stages
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_bar + 3);
    __nv_bfloat16* s_bias = reinterpret_cast<__nv_bfloat16*>(s_bar + 4);   // n_out bf16 (the bias is bf16 under autocast)

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        mbar_init(s_bar + 0, 1);
        mbar_init(s_bar + 1, 1);
        mbar_init(s_bar + 2, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < n_out; i += kPThreads) s_bias[i] = __float2bfloat16_rn(bias[i]);   // exact: rounded on the host
    if (warp == 0) {   // 512 columns of tensor memory: two stages of 256 float32 accumulator columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(512)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *s_tmem;

    if (tid == 0) {   // the whole weight image, once per CTA: one k-group (n_out rows x 16 bytes) per bulk copy
        const uint32_t group_bytes = (uint32_t)n_out * 16u;
        mbar_expect_tx(s_bar + 0, group_bytes * 16u);
        for (int g = 0; g < 16; ++g)
            bulk_g2s(s_w + (size_t)g * group_bytes, (const unsigned char*)w_image + (size_t)g * group_bytes, group_bytes,
                     s_bar + 0);
    }

    const int n_slabs = (n_out + kPN - 1) / kPN;
    const uint32_t a_addr = smem_u32(s_a), w_addr = smem_u32(s_w);
    const uint32_t w_lbo = (uint32_t)n_out * 16u;
    uint32_t phase[2] = {0u, 0u};
    bool w_ready = false;
    const int64_t n_tiles = (n_rows + kPM - 1) / kPM;

    auto issue_slab = [&](int j) {   // one thread: the eight K = 16 steps of output columns [256 j, 256 j + cols)
        const int cols = min(kPN, n_out - j * kPN);
        const uint32_t idesc = umma_idesc_bf16(kPM, cols);
        const uint32_t d = tmem_base + (uint32_t)(j & 1) * kPN;
#pragma unroll
        for (int ks = 0; ks < kPK / 16; ++ks) {
            const uint64_t adesc = umma_desc(a_addr + (uint32_t)(2 * ks) * kALbo, kALbo, 128);
            const uint64_t bdesc = umma_desc(w_addr + (uint32_t)(2 * ks) * w_lbo + (uint32_t)j * (kPN * 16), w_lbo, 128);
            umma_bf16(d, adesc, bdesc, idesc, ks > 0 ? 1u : 0u);
        }
        umma_commit(s_bar + 1 + (j & 1));
    };

    // A tile in registers: iteration `it` of warp w is row it * 8 + w of the tile, lane l its float4 l - a whole row
    // (512 contiguous bytes) per warp instruction.  Fetched one tile ahead, behind the epilogue of the current one.
    float4 v[16];
    auto fetch = [&](int64_t tile) {
#pragma unroll
        for (int it = 0; it < 16; ++it) {
            const int64_t row = tile * kPM + it * 8 + warp;
            v[it] = row < n_rows ? __ldg(reinterpret_cast<const float4*>(x + row * kPK) + lane)
                                 : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    if ((int64_t)blockIdx.x < n_tiles) fetch(blockIdx.x);

    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        // ---- A: float32 -> bf16 into the canonical layout: k-group lane / 2, half lane % 2 of row it * 8 + warp
#pragma unroll
        for (int it = 0; it < 16; ++it) {
            uint2 q;
            q.x = pack_bf16(v[it].x, v[it].y);
            q.y = pack_bf16(v[it].z, v[it].w);
            *reinterpret_cast<uint2*>(s_a + (size_t)(lane >> 1) * kALbo + (it * 8 + warp) * 16 + (lane & 1) * 8) = q;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic stores -> visible to the tensor core
        __syncthreads();
        if (!w_ready) {   // first tile: the weights have to be there (every thread waits: it also orders the async writes)
            mbar_wait(s_bar + 0, 0u);
            w_ready = true;
        }
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            issue_slab(0);
            if (n_slabs > 1) issue_slab(1);
        }
        if (tile + gridDim.x < n_tiles) fetch(tile + gridDim.x);   // the next tile's rows arrive behind the epilogue
        for (int j = 0; j < n_slabs; ++j) {
            const int stage = j & 1;
            mbar_wait(s_bar + 1 + stage, phase[stage]);
            phase[stage] ^= 1u;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            // ---- D: this thread's row, its half of the slab's columns
            {
                const int cols = min(kPN, n_out - j * kPN);
                const int q = warp & 3, chalf = warp >> 2;
                const int64_t row = tile * kPM + q * 32 + lane;
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)stage * kPN;
                // four chunks of 32 columns, software-pipelined: chunk k + 1 is on its way out of tensor memory while
                // chunk k gets its bias, is rounded and stored
                uint32_t acc[2][32];
                const int c_lo = chalf * 128;
                if (c_lo < cols) tmem_ld32(taddr + (uint32_t)c_lo, acc[0]);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int c0 = c_lo + 32 * k;
                    if (c0 >= cols) break;                    // warp-uniform
                    tmem_ld_wait();
                    if (k < 3 && c0 + 32 < cols) tmem_ld32(taddr + (uint32_t)(c0 + 32), acc[(k + 1) & 1]);
                    const uint32_t(&a)[32] = acc[k & 1];
                    if (row < n_rows) {
                        const uint4* b8 = reinterpret_cast<const uint4*>(s_bias + j * kPN + c0);   // 8 bf16 per load, broadcast
                        uint4* dst = reinterpret_cast<uint4*>(out + row * n_out + j * kPN + c0);
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const uint4 b = b8[i];   // bf16 -> float32: the bits, shifted up
                            uint4 o;
                            o.x = pack_bf16(__uint_as_float(a[8 * i + 0]) + __uint_as_float(b.x << 16),
                                            __uint_as_float(a[8 * i + 1]) + __uint_as_float(b.x & 0xffff0000u));
                            o.y = pack_bf16(__uint_as_float(a[8 * i + 2]) + __uint_as_float(b.y << 16),
                                            __uint_as_float(a[8 * i + 3]) + __uint_as_float(b.y & 0xffff0000u));
                            o.z = pack_bf16(__uint_as_float(a[8 * i + 4]) + __uint_as_float(b.z << 16),
                                            __uint_as_float(a[8 * i + 5]) + __uint_as_float(b.z & 0xffff0000u));
                            o.w = pack_bf16(__uint_as_float(a[8 * i + 6]) + __uint_as_float(b.w << 16),
                                            __uint_as_float(a[8 * i + 7]) + __uint_as_float(b.w & 0xffff0000u));
                            dst[i] = o;
                        }
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            if (j + 2 < n_slabs) {   // the stage just drained takes slab j + 2
                __syncthreads();
                if (tid == 0) {
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    issue_slab(j + 2);
                }
            }
        }
        __syncthreads();   // every MMA of the tile has completed (their commits were waited for): A and TMEM are free
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        if (!w_ready) mbar_wait(s_bar + 0, 0u);   // a CTA without tiles still has the weight copy in flight
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

}  // namespace adtfe

using namespace adtfe;

static uint16_t bf16_bits(float f) {   // round to nearest even, like torch's .to(torch.bfloat16)
    uint32_t u;
    memcpy(&u, &f, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);   // NaN stays NaN
    u += 0x7fffu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}

extern "C" int adtfe_linear_destroy(adtfe_linear* lin) {
    if (!lin) return ADTFE_OK;
    cudaSetDevice(lin->device);
    cudaFree(lin->w_image);
    cudaFree(lin->bias);
    delete lin;
    return ADTFE_OK;
}

extern "C" int adtfe_linear_create(int32_t n_in, int32_t n_out, const float* weight_host, const float* bias_host,
                                   int device, adtfe_linear** out) {
    ADTFE_REQUIRE(out && weight_host, ADTFE_ERR_BAD_ARG, "adtfe_linear_create: null pointer");
    *out = nullptr;
    ADTFE_REQUIRE(n_in == kPK, ADTFE_ERR_UNSUPPORTED, "adtfe_linear_create: n_in %d unsupported (the log-mel has 128 bands)", n_in);
    ADTFE_REQUIRE(n_out >= 32 && n_out <= 768 && n_out % 32 == 0, ADTFE_ERR_UNSUPPORTED,
                  "adtfe_linear_create: n_out %d unsupported (a multiple of 32 up to 768: the weight stays in shared memory)",
                  n_out);
    int rc = adtfe_device_ok(device);
    if (rc != ADTFE_OK) return rc;
    ADTFE_CUDA(cudaSetDevice(device));
    // canonical K-major image: offset(n, k) = (k / 8) * (n_out * 16) + n * 16 + (k % 8) * 2 bytes
    std::vector<uint16_t> image((size_t)n_out * kPK);
    for (int n = 0; n < n_out; ++n)
        for (int k = 0; k < kPK; ++k)
            image[((size_t)(k / 8) * n_out + n) * 8 + (k % 8)] = bf16_bits(weight_host[(size_t)n * kPK + k]);
    std::vector<float> bias(n_out, 0.0f);
    for (int n = 0; bias_host && n < n_out; ++n) {
        const uint32_t u = (uint32_t)bf16_bits(bias_host[n]) << 16;
        memcpy(&bias[n], &u, 4);
    }
    adtfe_linear* lin = new adtfe_linear();
    lin->device = device; lin->sm_count = device_sm_count(device); lin->n_in = n_in; lin->n_out = n_out;
    lin->smem_bytes = (size_t)n_out * 256 + kATileBytes + 32 + (size_t)n_out * 2 + 32;
    if (cudaMalloc(&lin->w_image, image.size() * 2) != cudaSuccess ||
        cudaMemcpy(lin->w_image, image.data(), image.size() * 2, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMalloc((void**)&lin->bias, (size_t)n_out * 4) != cudaSuccess ||
        cudaMemcpy(lin->bias, bias.data(), (size_t)n_out * 4, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaFuncSetAttribute(project_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lin->smem_bytes) != cudaSuccess) {
        set_error("adtfe_linear_create: device setup failed: %s", cudaGetErrorString(cudaGetLastError()));
        adtfe_linear_destroy(lin);
        return ADTFE_ERR_CUDA;
    }
    *out = lin;
    return ADTFE_OK;
}

extern "C" int adtfe_linear_forward(const adtfe_linear* lin, const float* x_dev, int64_t n_rows, void* out_bf16_dev,
                                    void* stream) {
    ADTFE_REQUIRE(lin && n_rows >= 0, ADTFE_ERR_BAD_ARG, "adtfe_linear_forward: bad argument");
    if (n_rows == 0) return ADTFE_OK;
    ADTFE_REQUIRE(x_dev && out_bf16_dev && ((uintptr_t)x_dev & 15) == 0 && ((uintptr_t)out_bf16_dev & 15) == 0,
                  ADTFE_ERR_BAD_ARG, "adtfe_linear_forward: null or misaligned buffer (16 bytes)");
    const int64_t n_tiles = (n_rows + kPM - 1) / kPM;
    const int grid = (int)std::min<int64_t>(n_tiles, lin->sm_count);
    project_kernel<<<grid, kPThreads, lin->smem_bytes, (cudaStream_t)stream>>>(x_dev, n_rows, lin->w_image, lin->bias,
                                                                              lin->n_out, (__nv_bfloat16*)out_bf16_dev);
    ADTFE_CUDA(cudaGetLastError());
    return ADTFE_OK;
}
