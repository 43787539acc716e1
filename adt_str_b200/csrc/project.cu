// project_to_mel on the 5th-generation tensor cores (sm_100a): tcgen05.mma, accumulators in tensor memory.
//
// Replaces the first layer behind the log-mel (reference model.py:224-226, 249):
//     self.project_to_mel = nn.Linear(n_mels, d_query * nhead);   src_emb = self.project_to_mel(src_emb)
// which runs under bf16 autocast (configs/train/setting-1.yaml: bf16): the float32 log-mel (rows = B * T, 128) and the
// weight are cast to bf16, the product is accumulated in float32, the bias (bf16) is added and the result is bf16.
// It is the one dense contraction next to the path: (rows, 128) x (128, 768).  With K = 128 the
// GEMM is bound by HBM, not by the tensor cores - 512 B read and 1536 B written per row against 197 kflop - so the
// kernel is built around the bytes: the cast of the log-mel is fused in (the library path runs a separate cast kernel,
// another 768 B per row), a CTA's part of the weight matrix stays in shared memory for its whole life, and every
// output element leaves once, through a TMA store.
//
// ONE persistent CTA of 16 warps per SM; the phases of a tile of 128 rows belong to different warps and are handed
// over through mbarriers, so that fetching, converting, multiplying, draining and storing all run at once:
//   warps 8..14  loaders: a warp instruction fetches one row (512 contiguous bytes), nineteen rows in flight per warp -
//                a whole tile per SM - then float32 -> bf16 into one of TWO A buffers in the K-major no-swizzle canonical
//                layout: 8-row x 16-byte core matrices, rows of a k-group contiguous
//                (offset(r, k) = (k / 8) * 2064 + r * 16 + (k % 8) * 2: LBO = 2064 B - 16 B of padding per k-group make
//                the warp's stores conflict-free - SBO = 128 B)
//   warp 15      one thread issues tcgen05.mma.cta_group::1.kind::f16 (M = 128, N = 128, K = 16, operands by
//                shared-memory descriptors): the part's columns in chunks of 128, each chunk into one of FOUR 128-column
//                accumulators in tensor memory (a ring: the tensor core runs up to four chunks ahead of the drain, across
//                tile boundaries); tcgen05.commit signals the chunk's accumulator and, behind the tile's last chunk,
//                frees the A buffer
//   warps 0..7   drain: warp w reads the TMEM lanes of quarter w % 4, half w / 4 of the chunk's columns (two
//                tcgen05.ld.32x32b.x32 in flight - one output row per thread), releases the accumulator as soon as the
//                values are in registers, adds the bias, rounds to bf16, parks 32 rows x 128 bytes in its staging rows
//                and hands them to the TMA (one tensor store per warp and chunk)
// The output columns are split over the CTAs of a tile in parts of at most 384 columns (768 -> 2 x 384: 96 KB of weight
// per CTA, resident for the life of the CTA), so a tile's rows are converted twice (the second read comes from L2).
//
// History (profiles/README.md): round 2 first shipped a kernel that ran a tile's phases one behind the other in every
// warp and leaned on a second CTA per SM to overlap them (parts of 192 columns, staged epilogue read back and stored by
// the warp): 2.6 ms on 4 M rows against 1.7 ms for cast + cuBLAS.  ncu on the first warp-specialised version showed
// the drain as the stage everything waits for, and under it the SM's load / store data pipe at 65 %: bias loads, the
// staging round trip and the stores were 3 000 of its 3 800 wavefronts per tile.  Hence the drain below.
#include <cuda.h>
#include <cuda_bf16.h>

#include <string.h>

#include <algorithm>
#include <vector>

#include "common.cuh"

struct adtfe_linear {
    int device = 0;
    int sm_count = 148;
    int32_t n_in = 128, n_out = 768;
    int32_t n_parts = 1;       // CTAs that share a tile of rows, each with its own range of output columns
    void* w_image = nullptr;   // bf16, canonical K-major layout per column part, parts back to back: n_out * 256 bytes
    float* bias = nullptr;     // n_out floats, already rounded to bf16 (autocast casts the bias too)
    size_t smem_bytes = 0;
};

namespace adtfe {

constexpr int kPM = 128;          // rows per tile = UMMA M
constexpr int kPK = 128;          // n_mels
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

constexpr int kLoaders = 7;                   // loader warps: rows lw, lw + 7, ... of the tile, 19 in flight per warp
constexpr int kLoadRows = (kPM + kLoaders - 1) / kLoaders;
constexpr int kPThreads = (8 + kLoaders + 1) * 32;   // 16 warps: 128 registers per thread (a 17th warp costs 32 of them -
                                               // warps are allocated in fours - and the drain then walks
                                               // load -> use -> load -> use for want of registers)
constexpr int kPartCols = 384;
constexpr int kChunk = 128;                   // columns per accumulator: 4 x 128 = the SM's 512 TMEM columns
constexpr int kStageBytes = 32 * 128;         // per drain warp: 32 rows x 64 bf16 columns
constexpr int kBars = 13;                     // weight | a_full[2] | a_empty[2] | acc_full[4] | acc_empty[4]

// Column split: the n_out columns go to n_parts = ceil(n_out / widest) CTAs per tile of rows, in multiples of 32 columns
// (768 -> 2 x 384, 512 -> 2 x 256, 416 -> 224 + 192).  The CTAs of a tile convert the same 128 rows (the second read
// comes from L2).
__host__ __device__ inline int n_parts_of(int n_out, int widest = kPartCols) { return (n_out + widest - 1) / widest; }
__host__ __device__ inline int part_cols(int n_out, int part, int widest = kPartCols) {
    const int u = n_out / 32, np = n_parts_of(n_out, widest);
    return 32 * (u / np + (part < u % np ? 1 : 0));
}
__host__ __device__ inline int part_col0(int n_out, int part, int widest = kPartCols) {
    int c = 0;
    for (int p = 0; p < part; ++p) c += part_cols(n_out, p, widest);
    return c;
}

// Drain of one chunk by one warp: NU = 1 or 2 units of 32 columns.  Every shared-memory access is issued in batches,
// ahead of its use (a warp that walks load -> use -> load -> use pays the latency of the shared-memory pipe once per
// access, and the drain is what the tensor core and the loaders wait for), and there are as few of them as possible -
// the load / store data pipe of the SM is the busiest unit of this kernel:
//   * the bias sits in shared memory as bf16 pairs (it IS bf16 under autocast): 4 broadcast loads per unit
//   * a warp's 32 rows x 64 columns leave through ONE TMA tensor store (cp.async.bulk.tensor, the staging rows in the
//     128-byte swizzle of the tensor map: 16-byte piece ^ (row & 7), which is also what keeps the row-wise STS.128 free
//     of bank conflicts); the TMA clips the rows beyond n_rows.  A warp with a single unit (parts whose width is not a
//     multiple of 64) reads its staging rows back and stores 64 contiguous bytes per row itself.
template <int NU>
__device__ __forceinline__ void drain_units(uint32_t taddr, uint64_t* acc_empty, const uint32_t* bias2, unsigned char* mine,
                                            int lane, const CUtensorMap* tmap, int col, __nv_bfloat16* out, int64_t row_base,
                                            int64_t n_rows, int n_out) {
    uint32_t a[NU][32];
#pragma unroll
    for (int j = 0; j < NU; ++j) tmem_ld32(taddr + (uint32_t)(32 * j), a[j]);
    uint4 bv[4 * NU];   // broadcast loads, in flight together with the tcgen05.ld
#pragma unroll
    for (int i = 0; i < 4 * NU; ++i) bv[i] = reinterpret_cast<const uint4*>(bias2)[i];
    tmem_ld_wait();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    mbar_arrive(acc_empty);   // the accumulator is in registers: the tensor core may overwrite it
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the previous chunk's tensor store has read the staging rows
    __syncwarp();
    const int swz = lane & 7;
#pragma unroll
    for (int j = 0; j < NU; ++j) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {   // 8 columns: one 16-byte piece
            const uint4 b = bv[4 * j + i];
            const uint32_t bw[4] = {b.x, b.y, b.z, b.w};
            uint32_t o[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float2 sum = __fadd2_rn(make_float2(__uint_as_float(a[j][8 * i + 2 * k]), __uint_as_float(a[j][8 * i + 2 * k + 1])),
                                              make_float2(__uint_as_float(bw[k] << 16), __uint_as_float(bw[k] & 0xffff0000u)));
                o[k] = pack_bf16(sum.x, sum.y);
            }
            *reinterpret_cast<uint4*>(mine + lane * 128 + (((4 * j + i) ^ swz) << 4)) = make_uint4(o[0], o[1], o[2], o[3]);
        }
    }
    if (NU == 2) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic stores -> visible to the TMA
        __syncwarp();
        if (lane == 0) {
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(tmap), "r"(col),
                         "r"((int)row_base), "r"(smem_u32(mine))
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    } else {
        __syncwarp();
        // consecutive lanes consecutive 16-byte pieces of a row: 4 pieces (64 contiguous bytes) per row, 8 rows per store
        const int r_lane = lane >> 2, piece = lane & 3;
        uint4 val[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int r = k * 8 + r_lane;
            val[k] = *reinterpret_cast<const uint4*>(mine + r * 128 + ((piece ^ (r & 7)) << 4));
        }
        __nv_bfloat16* dst = out + (row_base + r_lane) * n_out + col + piece * 8;
        const int64_t rows_left = n_rows - row_base - r_lane;   // rows of this lane's stride that exist
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (k * 8 < rows_left) *reinterpret_cast<uint4*>(dst + (size_t)k * 8 * n_out) = val[k];
        __syncwarp();   // the staging rows are rewritten by the next chunk
    }
}

__global__ void __launch_bounds__(kPThreads, 1) project_kernel(const float* __restrict__ x, int64_t n_rows,
                                                                      const void* __restrict__ w_image,
                                                                      const float* __restrict__ bias, int n_out,
                                                                      int n_parts, __nv_bfloat16* __restrict__ out,
                                                                      const __grid_constant__ CUtensorMap out_map) {
    extern __shared__ unsigned char smem_raw[];
    // the staging rows carry the tensor map's 128-byte swizzle, a function of the ADDRESS: 1024-byte aligned
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int part = (int)blockIdx.x % n_parts;
    const int nh = part_cols(n_out, part), col0 = part_col0(n_out, part);
    unsigned char* s_w = smem;                                                   // this part's weight image: nh * 256 bytes
    unsigned char* s_stage = smem + (size_t)part_cols(n_out, 0) * 256;   // 8 drain warps x kStageBytes
    unsigned char* s_a = s_stage + 8 * kStageBytes;                             // two A buffers
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_a + 2 * kATileBytes);
    uint64_t *w_full = s_bar, *a_full = s_bar + 1, *a_empty = s_bar + 3, *acc_full = s_bar + 5, *acc_empty = s_bar + 9;
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_bar + kBars);
    uint32_t* s_bias = reinterpret_cast<uint32_t*>(s_bar + kBars + 1);   // bf16 pairs: columns (2 k, 2 k + 1) of the part

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        mbar_init(w_full, 1);
        for (int b = 0; b < 2; ++b) {
            mbar_init(a_full + b, kLoaders * 32);   // every loader thread arrives behind its own stores and proxy fence
            mbar_init(a_empty + b, 1);     // tcgen05.commit
        }
        for (int k = 0; k < 4; ++k) {
            mbar_init(acc_full + k, 1);    // tcgen05.commit
            mbar_init(acc_empty + k, 256); // every drain thread arrives once its tcgen05.ld have completed
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < nh / 2; i += kPThreads)   // the bias has been rounded to bf16 on the host: the low halves are zero
        s_bias[i] = (__float_as_uint(bias[col0 + 2 * i]) >> 16) | (__float_as_uint(bias[col0 + 2 * i + 1]) & 0xffff0000u);
    if (warp == 0) {   // all 512 columns of tensor memory: one CTA per SM
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(512)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *s_tmem;

    const int n_tiles = (int)((n_rows + kPM - 1) / kPM);   // the host refuses 2^31 rows and more
    const int tile0 = (int)blockIdx.x / n_parts, tile_step = (int)gridDim.x / n_parts;
    const int n_chunks = (nh + kChunk - 1) / kChunk;

    if (warp >= 8 && warp < 8 + kLoaders) {
        // ---- loaders
        const int lw = warp - 8;
        uint32_t it = 0;
        for (int tile = tile0; tile < n_tiles; tile += tile_step, ++it) {
            const uint32_t b = it & 1u, use = it >> 1;
            float4 v[kLoadRows];
            const float4* src = reinterpret_cast<const float4*>(x + ((int64_t)tile * kPM + lw) * kPK) + lane;
            const int rows_left = (int)min((int64_t)kPM, n_rows - (int64_t)tile * kPM) - lw;   // of this warp's stride
#pragma unroll
            for (int j = 0; j < kLoadRows; ++j)
                v[j] = j * kLoaders < rows_left ? __ldg(src + (size_t)j * kLoaders * (kPK / 4))
                                                 : make_float4(0.f, 0.f, 0.f, 0.f);
            mbar_wait(a_empty + b, (use & 1u) ^ 1u);   // the MMAs that read this buffer two tiles ago have completed
            unsigned char* a = s_a + b * kATileBytes;
#pragma unroll
            for (int j = 0; j < kLoadRows; ++j) {
                if (j * kLoaders + lw >= kPM) break;   // the last stride of the tile is short
                uint2 qv;
                qv.x = pack_bf16(v[j].x, v[j].y);
                qv.y = pack_bf16(v[j].z, v[j].w);
                *reinterpret_cast<uint2*>(a + (size_t)(lane >> 1) * kALbo + (j * kLoaders + lw) * 16 + (lane & 1) * 8) = qv;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic stores -> visible to the tensor core
            mbar_arrive(a_full + b);
        }
    } else if (warp == 8 + kLoaders) {
        // ---- MMA issue
        if (lane == 0) {   // this part's weight image, once per CTA: one k-group (nh rows x 16 bytes) per bulk copy
            const uint32_t group_bytes = (uint32_t)nh * 16u;
            const unsigned char* src = (const unsigned char*)w_image + (size_t)col0 * 256;
            mbar_expect_tx(w_full, group_bytes * 16u);
            for (int g = 0; g < 16; ++g) bulk_g2s(s_w + (size_t)g * group_bytes, src + (size_t)g * group_bytes, group_bytes, w_full);
        }
        mbar_wait(w_full, 0u);
        const uint32_t a_addr = smem_u32(s_a), w_addr = smem_u32(s_w), w_lbo = (uint32_t)nh * 16u;
        uint32_t it = 0, acc = 0;
        for (int tile = tile0; tile < n_tiles; tile += tile_step, ++it) {
            const uint32_t b = it & 1u, use = it >> 1;
            mbar_wait(a_full + b, use & 1u);
            for (int c = 0; c < n_chunks; ++c, ++acc) {
                const uint32_t slot = acc & 3u;
                mbar_wait(acc_empty + slot, ((acc >> 2) & 1u) ^ 1u);   // drained four chunks ago
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (lane == 0) {
                    const int cw = min(kChunk, nh - c * kChunk);
                    const uint32_t idesc = umma_idesc_bf16(kPM, cw);
#pragma unroll
                    for (int ks = 0; ks < kPK / 16; ++ks) {
                        const uint64_t adesc = umma_desc(a_addr + b * kATileBytes + (uint32_t)(2 * ks) * kALbo, kALbo, 128);
                        const uint64_t bdesc = umma_desc(w_addr + (uint32_t)(2 * ks) * w_lbo + (uint32_t)(c * kChunk) * 16u, w_lbo, 128);
                        umma_bf16(tmem_base + slot * kChunk, adesc, bdesc, idesc, ks > 0 ? 1u : 0u);
                    }
                    umma_commit(acc_full + slot);
                    if (c == n_chunks - 1) umma_commit(a_empty + b);   // all MMAs of the tile: the A buffer is free
                }
                __syncwarp();
            }
        }
    } else {
        // ---- drain
        const int q = warp & 3, h = warp >> 2;
        unsigned char* mine = s_stage + warp * kStageBytes;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
        uint32_t acc = 0;
        for (int tile = tile0; tile < n_tiles; tile += tile_step) {
            const int64_t row_base = (int64_t)tile * kPM + q * 32;
            for (int c = 0; c < n_chunks; ++c, ++acc) {
                const uint32_t slot = acc & 3u;
                const int units = min(kChunk, nh - c * kChunk) / 32;
                const int u_lo = h == 0 ? 0 : (units + 1) / 2, u_hi = h == 0 ? (units + 1) / 2 : units;
                mbar_wait(acc_full + slot, (acc >> 2) & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const int ccol = c * kChunk + 32 * u_lo;   // first column of this warp's share, within the part
                const uint32_t ta = taddr + slot * kChunk + (uint32_t)(32 * u_lo);
                if (u_hi - u_lo == 2)
                    drain_units<2>(ta, acc_empty + slot, s_bias + ccol / 2, mine, lane, &out_map, col0 + ccol, out, row_base, n_rows, n_out);
                else if (u_hi - u_lo == 1)
                    drain_units<1>(ta, acc_empty + slot, s_bias + ccol / 2, mine, lane, &out_map, col0 + ccol, out, row_base, n_rows, n_out);
                else
                    mbar_arrive(acc_empty + slot);   // a chunk of 32 columns has nothing for the second half
            }
        }
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the last tensor store has read its staging rows
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();   // every accumulator has been drained, so every MMA has completed
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
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

// 1 KB of slack to align the staging rows | weight part | staging | two A buffers | mbarriers + TMEM address | bias pairs
static size_t smem_bytes_of(int widest_part) {
    return 1024 + (size_t)widest_part * 256 + 8 * kStageBytes + 2 * kATileBytes + (kBars + 1) * 8 + (size_t)widest_part * 2 + 32;
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
    // canonical K-major image per column part (nh columns from col0 on):
    // offset(n, k) = col0 * 256 + (k / 8) * (nh * 16) + (n - col0) * 16 + (k % 8) * 2 bytes
    const int n_parts = n_parts_of(n_out);
    std::vector<uint16_t> image((size_t)n_out * kPK);
    for (int h = 0; h < n_parts; ++h) {
        const int nh = part_cols(n_out, h), col0 = part_col0(n_out, h);
        for (int n = 0; n < nh; ++n)
            for (int k = 0; k < kPK; ++k)
                image[(size_t)col0 * kPK + ((size_t)(k / 8) * nh + n) * 8 + (k % 8)] =
                    bf16_bits(weight_host[(size_t)(col0 + n) * kPK + k]);
    }
    std::vector<float> bias(n_out, 0.0f);
    for (int n = 0; bias_host && n < n_out; ++n) {
        const uint32_t u = (uint32_t)bf16_bits(bias_host[n]) << 16;
        memcpy(&bias[n], &u, 4);
    }
    adtfe_linear* lin = new adtfe_linear();
    lin->device = device; lin->sm_count = device_sm_count(device); lin->n_in = n_in; lin->n_out = n_out;
    lin->n_parts = n_parts;
    lin->smem_bytes = smem_bytes_of(part_cols(n_out, 0));
    if (cudaMalloc(&lin->w_image, image.size() * 2) != cudaSuccess ||
        cudaMemcpy(lin->w_image, image.data(), image.size() * 2, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMalloc((void**)&lin->bias, (size_t)n_out * 4) != cudaSuccess ||
        cudaMemcpy(lin->bias, bias.data(), (size_t)n_out * 4, cudaMemcpyHostToDevice) != cudaSuccess ||
        // the attribute belongs to the function, not to the handle: the widest part any handle can have (a later,
        // narrower handle must not lower it under an earlier one's launch size)
        cudaFuncSetAttribute(project_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes_of(kPartCols)) != cudaSuccess) {
        set_error("adtfe_linear_create: device setup failed: %s", cudaGetErrorString(cudaGetLastError()));
        adtfe_linear_destroy(lin);
        return ADTFE_ERR_CUDA;
    }
    *out = lin;
    return ADTFE_OK;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {   // the driver's entry point, without linking libcuda
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr) != cudaSuccess ||
        qr != cudaDriverEntryPointSuccess)
        return nullptr;
    return (EncodeTiledFn)fn;
}

extern "C" int adtfe_linear_forward(const adtfe_linear* lin, const float* x_dev, int64_t n_rows, void* out_bf16_dev,
                                    void* stream) {
    ADTFE_REQUIRE(lin && n_rows >= 0 && n_rows < ((int64_t)1 << 31), ADTFE_ERR_BAD_ARG, "adtfe_linear_forward: bad argument");
    if (n_rows == 0) return ADTFE_OK;
    ADTFE_REQUIRE(x_dev && out_bf16_dev && ((uintptr_t)x_dev & 15) == 0 && ((uintptr_t)out_bf16_dev & 15) == 0,
                  ADTFE_ERR_BAD_ARG, "adtfe_linear_forward: null or misaligned buffer (16 bytes)");
    // the output as a 2-D tensor (n_out columns x n_rows rows, bf16) for the drain's TMA stores: boxes of 64 columns x
    // 32 rows, 128-byte swizzle on the shared-memory side; encoded per call (host arithmetic, ~1 us), passed by value
    static const EncodeTiledFn encode = encode_tiled_fn();
    ADTFE_REQUIRE(encode, ADTFE_ERR_CUDA, "adtfe_linear_forward: cuTensorMapEncodeTiled is not available");
    CUtensorMap out_map;
    memset(&out_map, 0, sizeof(out_map));
    const cuuint64_t dims[2] = {(cuuint64_t)lin->n_out, (cuuint64_t)n_rows};
    const cuuint64_t strides[1] = {(cuuint64_t)lin->n_out * 2};
    const cuuint32_t box[2] = {64, 32}, elem[2] = {1, 1};
    const CUresult cr = encode(&out_map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, out_bf16_dev, dims, strides, box, elem,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    ADTFE_REQUIRE(cr == CUDA_SUCCESS, ADTFE_ERR_CUDA, "adtfe_linear_forward: cuTensorMapEncodeTiled failed (%d)", (int)cr);
    const int64_t n_tiles = (n_rows + kPM - 1) / kPM;
    // one CTA per SM: n_parts CTAs per tile of rows
    const int grid = lin->n_parts * (int)std::min<int64_t>(n_tiles, std::max(1, lin->sm_count / lin->n_parts));
    project_kernel<<<grid, kPThreads, lin->smem_bytes, (cudaStream_t)stream>>>(x_dev, n_rows, lin->w_image, lin->bias,
                                                                              lin->n_out, lin->n_parts,
                                                                              (__nv_bfloat16*)out_bf16_dev, out_map);
    ADTFE_CUDA(cudaGetLastError());
    return ADTFE_OK;
}
