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
// Persistent CTAs of 8 warps, TWO per SM; the output columns are split over the CTAs of a tile of 128 rows in parts of at
// most 192 columns (768 -> 4 x 192), so that a CTA's weight part (48 KB), its bf16 copy of the tile (32 KB) and the
// epilogue staging (20 KB) fit twice into an SM - the second CTA is what overlaps the phases of a tile.  Per tile:
//   A      a warp instruction fetches one row (512 contiguous bytes), sixteen in flight per warp; each lane converts
//          its four floats to bf16 and stores them into the K-major no-swizzle canonical layout: 8-row x 16-byte core
//          matrices, rows of a k-group contiguous
//          (offset(r, k) = (k / 8) * 2064 + r * 16 + (k % 8) * 2: LBO = 2064 B - 16 B of padding per k-group make the
//          warp's stores conflict-free - SBO = 128 B)
//   MMA    one thread issues tcgen05.mma.cta_group::1.kind::f16 (M = 128, N = the part's columns, K = 16) eight times,
//          operands by shared-memory descriptors, D in tensor memory (256 columns per CTA); tcgen05.commit arrives on
//          the accumulator's mbarrier
//   D      all eight warps drain: warp w reads the TMEM lanes of its quarter (w % 4) and its half of the part's
//          32-column units (w / 4) with tcgen05.ld.32x32b.x32 - one output row per thread - adds the bias, rounds to
//          bf16 and parks the 32 x 32 block in the warp's staging rows; the warp then writes the block out together,
//          consecutive lanes consecutive 16-byte pieces: whole sectors, 64 contiguous bytes per row (a thread storing
//          its own row touches 32 different lines per instruction: 4x the L1 store transactions).
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
    // the streaming schedule (project_stream_kernel): parts of at most 384 columns, their own weight image
    int32_t n_parts_stream = 1;
    void* w_image_stream = nullptr;
    size_t smem_bytes_stream = 0;
    int32_t schedule = 0;      // adtfe_linear_force_schedule: 0 = by row count, 1 = column split, 2 = streaming
};

namespace adtfe {

constexpr int kPM = 128;          // rows per tile = UMMA M
constexpr int kPK = 128;          // n_mels
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

constexpr int kPartCols = 192;                // output columns per CTA at most: 48 KB of weight, 192 of 256 TMEM columns
constexpr int kStagePitch = 64;               // bytes per staged row: 32 bf16 columns; the 16-byte pieces of row r sit at
                                              // piece ^ ((r >> 1) & 3), which keeps both the row-wise stores (lane = row) and
                                              // the piece-wise loads (lane = quarter row) free of bank conflicts
constexpr int kStageBytes = 32 * kStagePitch; // per warp: its 32 rows x 32 columns

// Column split: the n_out columns go to n_parts = ceil(n_out / 192) CTAs per tile of rows, in multiples of 32 columns
// (768 -> 4 x 192, 512 -> 192 + 160 + 160, 256 -> 2 x 128).  A part's weight (<= 48 KB), its bf16 copy of the tile
// (32 KB) and the epilogue staging (20 KB) leave room for TWO CTAs per SM, which is what overlaps the phases of a tile
// - load and convert, MMA, drain and store - without any hand-off code: while one CTA drains, the other loads.  The
// CTAs of a tile convert the same 128 rows (the re-reads come from L2).
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

__global__ void __launch_bounds__(kPThreads, 2) project_kernel(const float* __restrict__ x, int64_t n_rows,
                                                               const void* __restrict__ w_image,
                                                               const float* __restrict__ bias, int n_out, int n_parts,
                                                               __nv_bfloat16* __restrict__ out) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int part = (int)blockIdx.x % n_parts;
    const int nh = part_cols(n_out, part), col0 = part_col0(n_out, part);
    unsigned char* s_w = smem;                                     // this part's weight image: nh * 256 bytes
    unsigned char* s_a = smem + (size_t)part_cols(n_out, 0) * 256; // kATileBytes (part 0 is the widest)
    unsigned char* s_stage = s_a + kATileBytes;                    // 8 warps x kStageBytes
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_stage + 8 * kStageBytes);   // [0]: weight copy, [1]: accumulator
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_bar + 2);
    float* s_bias = reinterpret_cast<float*>(s_bar + 4);   // nh floats holding bf16 values (the bias is bf16 under autocast)

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        mbar_init(s_bar + 0, 1);
        mbar_init(s_bar + 1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < nh; i += kPThreads) s_bias[i] = bias[col0 + i];   // rounded to bf16 on the host
    if (warp == 0) {   // 256 columns of tensor memory (two CTAs per SM share the 512)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(256)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *s_tmem;

    if (tid == 0) {   // this part's weight image, once per CTA: one k-group (nh rows x 16 bytes) per bulk copy
        const uint32_t group_bytes = (uint32_t)nh * 16u;
        const unsigned char* src = (const unsigned char*)w_image + (size_t)col0 * 256;   // the parts lie back to back
        mbar_expect_tx(s_bar + 0, group_bytes * 16u);
        for (int g = 0; g < 16; ++g) bulk_g2s(s_w + (size_t)g * group_bytes, src + (size_t)g * group_bytes, group_bytes, s_bar + 0);
    }

    const uint32_t a_addr = smem_u32(s_a), w_addr = smem_u32(s_w);
    const uint32_t w_lbo = (uint32_t)nh * 16u;
    const uint32_t idesc = umma_idesc_bf16(kPM, nh);
    uint32_t phase = 0u;
    bool w_ready = false;
    const int64_t n_tiles = (n_rows + kPM - 1) / kPM;
    const int64_t tile0 = (int64_t)blockIdx.x / n_parts, tile_step = (int64_t)gridDim.x / n_parts;
    // the epilogue's share of this warp: TMEM lanes of quarter q, the 32-column units [u_lo, u_hi) of the part
    const int q = warp & 3, units = nh / 32;
    const int u_lo = (warp >> 2) == 0 ? 0 : (units + 1) / 2, u_hi = (warp >> 2) == 0 ? (units + 1) / 2 : units;
    unsigned char* mine = s_stage + warp * kStageBytes;

    // A tile in registers: iteration `it` of warp w is row it * 8 + w of the tile, lane l its float4 l - a whole row
    // (512 contiguous bytes) per warp instruction.  Fetched one tile AHEAD, behind the epilogue of the current one.
    float4 v[16];
    auto fetch = [&](int64_t tile) {
#pragma unroll
        for (int it = 0; it < 16; ++it) {
            const int64_t row = tile * kPM + it * 8 + warp;
            v[it] = row < n_rows ? __ldg(reinterpret_cast<const float4*>(x + row * kPK) + lane)
                                 : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    if (tile0 < n_tiles) fetch(tile0);

    for (int64_t tile = tile0; tile < n_tiles; tile += tile_step) {
        // ---- A: float32 -> bf16 into the canonical layout: k-group lane / 2, half lane % 2 of row it * 8 + warp
#pragma unroll
        for (int it = 0; it < 16; ++it) {
            uint2 qv;
            qv.x = pack_bf16(v[it].x, v[it].y);
            qv.y = pack_bf16(v[it].z, v[it].w);
            *reinterpret_cast<uint2*>(s_a + (size_t)(lane >> 1) * kALbo + (it * 8 + warp) * 16 + (lane & 1) * 8) = qv;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic stores -> visible to the tensor core
        __syncthreads();
        if (!w_ready) {   // first tile: the weights have to be there (every thread waits: it also orders the async writes)
            mbar_wait(s_bar + 0, 0u);
            w_ready = true;
        }
        if (tid == 0) {   // one thread: the eight K = 16 steps of this part's columns
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int ks = 0; ks < kPK / 16; ++ks) {
                const uint64_t adesc = umma_desc(a_addr + (uint32_t)(2 * ks) * kALbo, kALbo, 128);
                const uint64_t bdesc = umma_desc(w_addr + (uint32_t)(2 * ks) * w_lbo, w_lbo, 128);
                umma_bf16(tmem_base, adesc, bdesc, idesc, ks > 0 ? 1u : 0u);
            }
            umma_commit(s_bar + 1);
        }
        if (tile + tile_step < n_tiles) fetch(tile + tile_step);   // the next tile's rows arrive behind the epilogue
        mbar_wait(s_bar + 1, phase);
        phase ^= 1u;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // ---- D: the warp drains its TMEM lanes, 32 columns per tcgen05.ld (one output row per thread), adds the bias,
        // rounds to bf16 and parks the 32 x 32 block in its staging rows; the warp then writes the block out together,
        // consecutive lanes consecutive 16-byte pieces of a row: whole 32-byte sectors, four pieces (64 contiguous
        // bytes) per row and eight rows per store instruction (a thread storing its own row touches 32 different lines
        // per instruction: 4x the L1 store transactions, 3.3 ms against the library's 1.7 ms on 4 M rows)
        {
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
            const int64_t row_base = tile * kPM + q * 32;
            for (int u = u_lo; u < u_hi; ++u) {   // warp-uniform: at most three units (192 columns over two warp groups)
                uint32_t a[32];
                tmem_ld32(taddr + (uint32_t)(32 * u), a);
                tmem_ld_wait();
                const float4* b4 = reinterpret_cast<const float4*>(s_bias + 32 * u);   // broadcast loads
                const int swz = (lane >> 1) & 3;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float4 b0 = b4[2 * i], b1 = b4[2 * i + 1];
                    uint4 o;
                    o.x = pack_bf16(__uint_as_float(a[8 * i + 0]) + b0.x, __uint_as_float(a[8 * i + 1]) + b0.y);
                    o.y = pack_bf16(__uint_as_float(a[8 * i + 2]) + b0.z, __uint_as_float(a[8 * i + 3]) + b0.w);
                    o.z = pack_bf16(__uint_as_float(a[8 * i + 4]) + b1.x, __uint_as_float(a[8 * i + 5]) + b1.y);
                    o.w = pack_bf16(__uint_as_float(a[8 * i + 6]) + b1.z, __uint_as_float(a[8 * i + 7]) + b1.w);
                    *reinterpret_cast<uint4*>(mine + lane * kStagePitch + ((i ^ swz) << 4)) = o;
                }
                __syncwarp();
                __nv_bfloat16* obase = out + (size_t)col0 + 32 * u;
#pragma unroll
                for (int r0 = 0; r0 < 32; r0 += 8) {
                    const int r = r0 + (lane >> 2), piece = lane & 3;
                    const uint4 val = *reinterpret_cast<const uint4*>(mine + r * kStagePitch + ((piece ^ ((r >> 1) & 3)) << 4));
                    if (row_base + r < n_rows) *reinterpret_cast<uint4*>(obase + (row_base + r) * n_out + piece * 8) = val;
                }
                __syncwarp();   // the staging rows are rewritten by the next unit
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();   // the MMAs of the tile have completed (their commit was waited for) and the accumulator is
                           // drained: A and TMEM are free
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        if (!w_ready) mbar_wait(s_bar + 0, 0u);   // a CTA without tiles still has the weight copy in flight
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256) : "memory");
    }
}

// ---- the streaming schedule: many tiles per CTA (a whole step's log-mel at once) ----------------------------------
// The column-split kernel above runs a tile's phases - fetch, convert, MMA, drain, store - one behind the other in
// every warp and leans on a second CTA per SM to overlap them: right for the single training batch the reference
// passes per call (two rounds of tiles), latency-bound for millions of rows.  Here the phases belong to different
// warps of ONE CTA per SM and a tile's hand-overs are mbarriers:
//   warps 8..14  loaders: a warp instruction fetches one row (512 contiguous bytes), nineteen rows in flight per warp -
//                a whole tile per SM - then float32 -> bf16 into one of TWO A buffers (same canonical layout)
//   warp 15      one thread issues the tcgen05.mma: the part's columns in chunks of 128, each chunk into one of FOUR
//                128-column accumulators in tensor memory (a ring: the tensor core runs up to four chunks ahead of
//                the drain, across tile boundaries); tcgen05.commit signals the chunk's accumulator and, behind the
//                tile's last chunk, frees the A buffer
//   warps 0..7   drain: warp w reads the TMEM lanes of quarter w % 4, half w / 4 of the chunk's columns (two
//                tcgen05.ld.32x32b.x32 in flight), releases the accumulator as soon as the values are in registers,
//                adds the bias, rounds to bf16, parks 32 rows x 128 bytes in its staging rows (16-byte pieces
//                XOR-swizzled by the row) and stores whole 128-byte lines
// The part is up to 384 columns (768 -> 2 x 384: 96 KB of weight per CTA), so a tile's rows are converted twice
// instead of four times.
constexpr int kSLoaders = 7;                   // loader warps: rows lw, lw + 7, ... of the tile, 19 in flight per warp
constexpr int kSLoadRows = (kPM + kSLoaders - 1) / kSLoaders;
constexpr int kSThreads = (8 + kSLoaders + 1) * 32;   // 16 warps: 128 registers per thread (a 17th warp costs 32 of them -
                                               // warps are allocated in fours - and the drain then walks
                                               // load -> use -> load -> use for want of registers)
constexpr int kSPartCols = 384;
constexpr int kSChunk = 128;                   // columns per accumulator: 4 x 128 = the SM's 512 TMEM columns
constexpr int kSStageBytes = 32 * 128;         // per drain warp: 32 rows x 64 bf16 columns
constexpr int kSBars = 13;                     // weight | a_full[2] | a_empty[2] | acc_full[4] | acc_empty[4]

template <int NU>   // 32-column units of this warp in the chunk: 1 or 2
__device__ __forceinline__ void drain_units(uint32_t taddr, uint64_t* acc_empty, const float* bias, unsigned char* mine,
                                            int lane, __nv_bfloat16* obase, int64_t row_base, int64_t n_rows, int n_out) {
    // Every shared-memory access of the drain is issued in batches, ahead of its use: a warp that walks load -> use ->
    // load -> use pays the latency of the shared-memory pipe 24 times per chunk, and the drain is what the tensor core
    // and the loaders wait for.
    uint32_t a[NU][32];
#pragma unroll
    for (int j = 0; j < NU; ++j) tmem_ld32(taddr + (uint32_t)(32 * j), a[j]);
    float4 bv[8];   // the bias of unit 0 (broadcast loads), in flight together with the tcgen05.ld
#pragma unroll
    for (int i = 0; i < 8; ++i) bv[i] = reinterpret_cast<const float4*>(bias)[i];
    tmem_ld_wait();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    mbar_arrive(acc_empty);   // the accumulator is in registers: the tensor core may overwrite it
    const int swz = lane & 7;
#pragma unroll
    for (int j = 0; j < NU; ++j) {
        uint4 o[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float4 b0 = bv[2 * i], b1 = bv[2 * i + 1];
            const float2 s0 = __fadd2_rn(make_float2(__uint_as_float(a[j][8 * i + 0]), __uint_as_float(a[j][8 * i + 1])), make_float2(b0.x, b0.y));
            const float2 s1 = __fadd2_rn(make_float2(__uint_as_float(a[j][8 * i + 2]), __uint_as_float(a[j][8 * i + 3])), make_float2(b0.z, b0.w));
            const float2 s2 = __fadd2_rn(make_float2(__uint_as_float(a[j][8 * i + 4]), __uint_as_float(a[j][8 * i + 5])), make_float2(b1.x, b1.y));
            const float2 s3 = __fadd2_rn(make_float2(__uint_as_float(a[j][8 * i + 6]), __uint_as_float(a[j][8 * i + 7])), make_float2(b1.z, b1.w));
            o[i] = make_uint4(pack_bf16(s0.x, s0.y), pack_bf16(s1.x, s1.y), pack_bf16(s2.x, s2.y), pack_bf16(s3.x, s3.y));
        }
        if (j + 1 < NU) {   // the next unit's bias, behind the last use of this one's
#pragma unroll
            for (int i = 0; i < 8; ++i) bv[i] = reinterpret_cast<const float4*>(bias + 32 * (j + 1))[i];
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) *reinterpret_cast<uint4*>(mine + lane * 128 + (((4 * j + i) ^ swz) << 4)) = o[i];
    }
    __syncwarp();
    // consecutive lanes consecutive 16-byte pieces of a row: 4 * NU pieces (64 or 128 contiguous bytes) per row
    constexpr int kPieces = 4 * NU, kRowsPerStore = 32 / kPieces, kStores = 32 / kRowsPerStore;
    const int r_lane = lane / kPieces, piece = lane % kPieces;
    uint4 val[kStores];
#pragma unroll
    for (int k = 0; k < kStores; ++k) {
        const int r = k * kRowsPerStore + r_lane;
        val[k] = *reinterpret_cast<const uint4*>(mine + r * 128 + ((piece ^ (r & 7)) << 4));
    }
    __nv_bfloat16* dst = obase + (row_base + r_lane) * n_out + piece * 8;
    const int64_t rows_left = n_rows - row_base - r_lane;   // rows of this lane's stride that exist
#pragma unroll
    for (int k = 0; k < kStores; ++k)
        if (k * kRowsPerStore < rows_left) *reinterpret_cast<uint4*>(dst + (size_t)k * kRowsPerStore * n_out) = val[k];
    __syncwarp();   // the staging rows are rewritten by the next chunk
}

__global__ void __launch_bounds__(kSThreads, 1) project_stream_kernel(const float* __restrict__ x, int64_t n_rows,
                                                                      const void* __restrict__ w_image,
                                                                      const float* __restrict__ bias, int n_out,
                                                                      int n_parts, __nv_bfloat16* __restrict__ out) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int part = (int)blockIdx.x % n_parts;
    const int nh = part_cols(n_out, part, kSPartCols), col0 = part_col0(n_out, part, kSPartCols);
    unsigned char* s_w = smem;                                                  // this part's weight image: nh * 256 bytes
    unsigned char* s_a = smem + (size_t)part_cols(n_out, 0, kSPartCols) * 256;  // two A buffers
    unsigned char* s_stage = s_a + 2 * kATileBytes;                             // 8 drain warps x kSStageBytes
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_stage + 8 * kSStageBytes);
    uint64_t *w_full = s_bar, *a_full = s_bar + 1, *a_empty = s_bar + 3, *acc_full = s_bar + 5, *acc_empty = s_bar + 9;
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_bar + kSBars);
    float* s_bias = reinterpret_cast<float*>(s_bar + kSBars + 1);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        mbar_init(w_full, 1);
        for (int b = 0; b < 2; ++b) {
            mbar_init(a_full + b, kSLoaders * 32);   // every loader thread arrives behind its own stores and proxy fence
            mbar_init(a_empty + b, 1);     // tcgen05.commit
        }
        for (int k = 0; k < 4; ++k) {
            mbar_init(acc_full + k, 1);    // tcgen05.commit
            mbar_init(acc_empty + k, 256); // every drain thread arrives once its tcgen05.ld have completed
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < nh; i += kSThreads) s_bias[i] = bias[col0 + i];
    if (warp == 0) {   // all 512 columns of tensor memory: one CTA per SM
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(512)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *s_tmem;

    const int n_tiles = (int)((n_rows + kPM - 1) / kPM);   // the host refuses more than 2^31 - 1 tiles
    const int tile0 = (int)blockIdx.x / n_parts, tile_step = (int)gridDim.x / n_parts;
    const int n_chunks = (nh + kSChunk - 1) / kSChunk;

    if (warp >= 8 && warp < 8 + kSLoaders) {
        // ---- loaders
        const int lw = warp - 8;
        uint32_t it = 0;
        for (int tile = tile0; tile < n_tiles; tile += tile_step, ++it) {
            const uint32_t b = it & 1u, use = it >> 1;
            float4 v[kSLoadRows];
            const float4* src = reinterpret_cast<const float4*>(x + ((int64_t)tile * kPM + lw) * kPK) + lane;
            const int rows_left = (int)min((int64_t)kPM, n_rows - (int64_t)tile * kPM) - lw;   // of this warp's stride
#pragma unroll
            for (int j = 0; j < kSLoadRows; ++j)
                v[j] = j * kSLoaders < rows_left ? __ldg(src + (size_t)j * kSLoaders * (kPK / 4))
                                                 : make_float4(0.f, 0.f, 0.f, 0.f);
            mbar_wait(a_empty + b, (use & 1u) ^ 1u);   // the MMAs that read this buffer two tiles ago have completed
            unsigned char* a = s_a + b * kATileBytes;
#pragma unroll
            for (int j = 0; j < kSLoadRows; ++j) {
                if (j * kSLoaders + lw >= kPM) break;   // the last stride of the tile is short
                uint2 qv;
                qv.x = pack_bf16(v[j].x, v[j].y);
                qv.y = pack_bf16(v[j].z, v[j].w);
                *reinterpret_cast<uint2*>(a + (size_t)(lane >> 1) * kALbo + (j * kSLoaders + lw) * 16 + (lane & 1) * 8) = qv;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic stores -> visible to the tensor core
            mbar_arrive(a_full + b);
        }
    } else if (warp == 8 + kSLoaders) {
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
                    const int cw = min(kSChunk, nh - c * kSChunk);
                    const uint32_t idesc = umma_idesc_bf16(kPM, cw);
#pragma unroll
                    for (int ks = 0; ks < kPK / 16; ++ks) {
                        const uint64_t adesc = umma_desc(a_addr + b * kATileBytes + (uint32_t)(2 * ks) * kALbo, kALbo, 128);
                        const uint64_t bdesc = umma_desc(w_addr + (uint32_t)(2 * ks) * w_lbo + (uint32_t)(c * kSChunk) * 16u, w_lbo, 128);
                        umma_bf16(tmem_base + slot * kSChunk, adesc, bdesc, idesc, ks > 0 ? 1u : 0u);
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
        unsigned char* mine = s_stage + warp * kSStageBytes;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
        uint32_t acc = 0;
        for (int tile = tile0; tile < n_tiles; tile += tile_step) {
            const int64_t row_base = (int64_t)tile * kPM + q * 32;
            for (int c = 0; c < n_chunks; ++c, ++acc) {
                const uint32_t slot = acc & 3u;
                const int units = min(kSChunk, nh - c * kSChunk) / 32;
                const int u_lo = h == 0 ? 0 : (units + 1) / 2, u_hi = h == 0 ? (units + 1) / 2 : units;
                mbar_wait(acc_full + slot, (acc >> 2) & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const int ccol = c * kSChunk + 32 * u_lo;   // first column of this warp's share, within the part
                const uint32_t ta = taddr + slot * kSChunk + (uint32_t)(32 * u_lo);
                if (u_hi - u_lo == 2)
                    drain_units<2>(ta, acc_empty + slot, s_bias + ccol, mine, lane, out + col0 + ccol, row_base, n_rows, n_out);
                else if (u_hi - u_lo == 1)
                    drain_units<1>(ta, acc_empty + slot, s_bias + ccol, mine, lane, out + col0 + ccol, row_base, n_rows, n_out);
                else
                    mbar_arrive(acc_empty + slot);   // a chunk of 32 columns has nothing for the second half
            }
        }
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
    cudaFree(lin->w_image_stream);
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
    // canonical K-major image per column part (nh columns from col0 on):
    // offset(n, k) = col0 * 256 + (k / 8) * (nh * 16) + (n - col0) * 16 + (k % 8) * 2 bytes
    auto make_image = [&](int widest) {
        std::vector<uint16_t> image((size_t)n_out * kPK);
        for (int h = 0; h < n_parts_of(n_out, widest); ++h) {
            const int nh = part_cols(n_out, h, widest), col0 = part_col0(n_out, h, widest);
            for (int n = 0; n < nh; ++n)
                for (int k = 0; k < kPK; ++k)
                    image[(size_t)col0 * kPK + ((size_t)(k / 8) * nh + n) * 8 + (k % 8)] =
                        bf16_bits(weight_host[(size_t)(col0 + n) * kPK + k]);
        }
        return image;
    };
    const int n_parts = n_parts_of(n_out);
    const std::vector<uint16_t> image = make_image(kPartCols), image_stream = make_image(kSPartCols);
    std::vector<float> bias(n_out, 0.0f);
    for (int n = 0; bias_host && n < n_out; ++n) {
        const uint32_t u = (uint32_t)bf16_bits(bias_host[n]) << 16;
        memcpy(&bias[n], &u, 4);
    }
    adtfe_linear* lin = new adtfe_linear();
    lin->device = device; lin->sm_count = device_sm_count(device); lin->n_in = n_in; lin->n_out = n_out;
    lin->n_parts = n_parts;
    lin->smem_bytes = (size_t)part_cols(n_out, 0) * 256 + kATileBytes + 8 * kStageBytes + 32 +
                      (size_t)part_cols(n_out, 0) * 4 + 32;
    lin->n_parts_stream = n_parts_of(n_out, kSPartCols);
    lin->smem_bytes_stream = (size_t)part_cols(n_out, 0, kSPartCols) * 256 + 2 * kATileBytes + 8 * kSStageBytes +
                             (kSBars + 1) * 8 + (size_t)part_cols(n_out, 0, kSPartCols) * 4 + 32;
    if (cudaMalloc(&lin->w_image, image.size() * 2) != cudaSuccess ||
        cudaMemcpy(lin->w_image, image.data(), image.size() * 2, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMalloc(&lin->w_image_stream, image_stream.size() * 2) != cudaSuccess ||
        cudaMemcpy(lin->w_image_stream, image_stream.data(), image_stream.size() * 2, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaFuncSetAttribute(project_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lin->smem_bytes_stream) != cudaSuccess ||
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

extern "C" int adtfe_linear_force_schedule(adtfe_linear* lin, int32_t schedule) {
    ADTFE_REQUIRE(lin && schedule >= 0 && schedule <= 2, ADTFE_ERR_BAD_ARG, "adtfe_linear_force_schedule: bad argument");
    lin->schedule = schedule;
    return ADTFE_OK;
}

constexpr int kStreamMinTilesPerSm = 4;

extern "C" int adtfe_linear_forward(const adtfe_linear* lin, const float* x_dev, int64_t n_rows, void* out_bf16_dev,
                                    void* stream) {
    ADTFE_REQUIRE(lin && n_rows >= 0, ADTFE_ERR_BAD_ARG, "adtfe_linear_forward: bad argument");
    if (n_rows == 0) return ADTFE_OK;
    ADTFE_REQUIRE(x_dev && out_bf16_dev && ((uintptr_t)x_dev & 15) == 0 && ((uintptr_t)out_bf16_dev & 15) == 0,
                  ADTFE_ERR_BAD_ARG, "adtfe_linear_forward: null or misaligned buffer (16 bytes)");
    const int64_t n_tiles = (n_rows + kPM - 1) / kPM;
    ADTFE_REQUIRE(n_tiles < ((int64_t)1 << 31), ADTFE_ERR_BAD_ARG, "adtfe_linear_forward: too many rows");
    // Many tiles per SM: the streaming schedule (one CTA per SM, the phases of a tile in different warps).  A call
    // with a few rounds of tiles - the reference's single batch of 64 x 246 rows is 123 tiles - is a matter of
    // latency: the column split spreads it over two CTAs per SM.
    const bool stream_it = lin->schedule == 2 || (lin->schedule == 0 && n_tiles >= (int64_t)kStreamMinTilesPerSm * lin->sm_count);
    if (stream_it) {
        const int np = lin->n_parts_stream;
        const int grid = np * (int)std::min<int64_t>(n_tiles, std::max(1, lin->sm_count / np));
        project_stream_kernel<<<grid, kSThreads, lin->smem_bytes_stream, (cudaStream_t)stream>>>(
            x_dev, n_rows, lin->w_image_stream, lin->bias, lin->n_out, np, (__nv_bfloat16*)out_bf16_dev);
        ADTFE_CUDA(cudaGetLastError());
        return ADTFE_OK;
    }
    // two CTAs per SM: n_parts CTAs per tile of rows
    const int grid = lin->n_parts * (int)std::min<int64_t>(n_tiles, std::max(1, 2 * lin->sm_count / lin->n_parts));
    project_kernel<<<grid, kPThreads, lin->smem_bytes, (cudaStream_t)stream>>>(x_dev, n_rows, lin->w_image, lin->bias,
                                                                              lin->n_out, lin->n_parts,
                                                                              (__nv_bfloat16*)out_bf16_dev);
    ADTFE_CUDA(cudaGetLastError());
    return ADTFE_OK;
}
