// Overlap-add tile mixer (sm_100a): peaks -> tiles -> normalise.
//
// Replaces the audio arithmetic of SynthDrum.__call__ (reference modules/synthetiser.py):
//   drum_rendering 214-239:  o = (1-mixup)*a + mixup*b ; o /= max|o| ; o *= vol ;
//                            track[start : start+len] += o[:len]
//   instrument_mixer 149-156: wav = sum_i track_i * gain_i ; wav / max|wav| * max_volume
// and the zero padding of collate_fn (data_modules/train_dataset.py:53).
//
// The two data-dependent maxima make it three kernels per chunk of segments:
//   1. peak_kernel      one CTA per (segment, instrument, 4096-sample chunk): both one-shots are
//                       read once and max|ca*a + cb*b| is taken for every note of that
//                       instrument at the same time (each note has its own mixup); warps and chunks
//                       meet in an atomicMax on the float bits (order-independent); chunk 0 also
//                       resolves the bank lookups into ResolvedEvent records.
//   2. slice_kernel     per (tile, note) one 32-byte record with everything the mixer needs (copy source and
//                       size, alignment, covered range, coefficients), in tile order.
//      mix_kernel       persistent one-warp CTAs take 2048-sample output tiles from a queue; each warp streams
//                       its tile's one-shot slices through its own shared-memory ring with TMA bulk copies
//                       and adds them in event order (the host-built CSR tile -> events) into registers, so
//                       the result is deterministic and needs no atomics; writes the raw tile and its |max|.
//   3. normalise_kernel row peak = max of the tile maxima; wav / peak * max_volume in place
//                       (an all-zero mix gives NaN, like the reference's 0/0).
#include <algorithm>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace adtfe {

constexpr int kPeakWarps = 8;        // warps per CTA of the peak pass: one (segment, instrument) group per warp
constexpr int kPeakNotes = ADTFE_PEAK_NOTES;   // notes of one group bounded together (they share every block scan)
constexpr int kBlock = ADTFE_PEAK_BLOCK;   // samples per block of the bank's block maxima

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

__device__ __forceinline__ float nan_max(float a, float b) {  // torch.max propagates NaN
    return (a != a || b != b) ? __int_as_float(0x7fc00000) : fmaxf(a, b);
}

// (ca*a.x + cb*b.x, ca*a.y + cb*b.y) for two samples in two packed sm_100a instructions: fl(cb*b + fl(ca*a)).
// torch rounds both products before the sum (mul, mul, add, synthetiser.py:223); this form skips the rounding
// of the second product, so a peak can differ from the reference's in its last bit (<= 6e-8 relative, two
// orders below the 1e-5 waveform tolerance).  ptxas contracts packed mul + add into FFMA2 whatever the source
// says (even explicit .rn PTX under --fmad=false), so the contraction is spelled out rather than left to it.
__device__ __forceinline__ float2 mix2(float ax, float ay, float bx, float by, float ca, float cb) {
    return __ffma2_rn(make_float2(bx, by), make_float2(cb, cb), __fmul2_rn(make_float2(ax, ay), make_float2(ca, ca)));
}
// the same arithmetic on the block maxima: fl(cb*B + fl(ca*A)) >= |fl(cb*b + fl(ca*a))| for every sample of the block,
// because ca, cb >= 0, |a| <= A, |b| <= B and rounding to nearest is monotonic - a rigorous bound, no slack needed
__device__ __forceinline__ float mix_bound(float A, float B, float ca, float cb) {
    return __fmaf_rn(cb, B, __fmul_rn(ca, A));
}

// max |x| over every kBlock-sample block of every one-shot (bank creation, once): one CTA per one-shot, one warp per
// block.  NaN samples are ignored, as fmaxf ignores them in the peak pass itself.
__global__ void __launch_bounds__(256) blockmax_kernel(const float* __restrict__ pcm, const int64_t* __restrict__ offsets,
                                                       const int32_t* __restrict__ lengths,
                                                       const int32_t* __restrict__ bm_off, float* __restrict__ bm) {
    const int id = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float* x = pcm + offsets[id];
    const int len = lengths[id], nblk = (len + kBlock - 1) / kBlock;
    for (int k = warp; k < nblk; k += 8) {
        float m = 0.0f;
        for (int i = k * kBlock + lane; i < min(len, (k + 1) * kBlock); i += 32) m = fmaxf(m, fabsf(x[i]));
        m = warp_max(m);
        if (lane == 0) bm[bm_off[id] + k] = m;
    }
}

// Exact max |ca*a + cb*b| over block k of the mixed one-shot for NC notes at once: 2 float4 per lane and one-shot, per
// pair of samples two packed fp32 instructions and one three-input FMNMX per note; the warp's maximum by one REDUX on
// the float bits (non-negative floats order like unsigned integers), returned to every lane.
template <int NC>
__device__ __forceinline__ void scan_block(const float4* __restrict__ a4, const float4* __restrict__ b4, int la, int lb,
                                           int k, int lane, const float (&ca)[NC], const float (&cb)[NC], float (&best)[NC]) {
    constexpr int kIters = kBlock / 128;
    float4 va[kIters], vb[kIters];
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    // every one-shot is padded to 4 floats, so whole float4s up to the padded length are readable
    const int la4 = (la + 3) >> 2, lb4 = (lb + 3) >> 2;
#pragma unroll
    for (int it = 0; it < kIters; ++it) {
        const int i4 = k * (kBlock / 4) + lane + 32 * it;
        va[it] = i4 < la4 ? __ldg(a4 + i4) : z;
        vb[it] = i4 < lb4 ? __ldg(b4 + i4) : z;
        // the float4 that holds a one-shot's last sample: ignore whatever pads it
        if (i4 == (la >> 2) && (la & 3)) {
            if ((la & 3) < 2) va[it].y = 0.f;
            if ((la & 3) < 3) va[it].z = 0.f;
            va[it].w = 0.f;
        }
        if (i4 == (lb >> 2) && (lb & 3)) {
            if ((lb & 3) < 2) vb[it].y = 0.f;
            if ((lb & 3) < 3) vb[it].z = 0.f;
            vb[it].w = 0.f;
        }
    }
#pragma unroll
    for (int i = 0; i < NC; ++i) {
        float m = 0.0f;
#pragma unroll
        for (int it = 0; it < kIters; ++it) {
            const float2 lo = mix2(va[it].x, va[it].y, vb[it].x, vb[it].y, ca[i], cb[i]);
            const float2 hi = mix2(va[it].z, va[it].w, vb[it].z, vb[it].w, ca[i], cb[i]);
            m = fmaxf(fmaxf(m, fabsf(lo.x)), fabsf(lo.y));
            m = fmaxf(fmaxf(m, fabsf(hi.x)), fabsf(hi.y));
        }
        best[i] = fmaxf(best[i], __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(m))));
    }
}

// Peaks of NC notes of one group by branch and bound over the blocks of the mixed one-shot.  The bank holds max |x| of
// every kBlock-sample block of every one-shot; the bound of block k for a note is mix_bound(A_k, B_k) - the note's own
// arithmetic on the block maxima.  The block with the largest bound (of the first note) is scanned exactly, which
// gives every note a lower bound of its peak; then only blocks whose bound still exceeds a note's lower bound are
// scanned (in ascending order: a drum one-shot decays, so the survivors are the first few blocks).  The maximum found is
// the maximum over all samples - bit-identical to a full scan - while a decaying one-shot costs a few KB of reads
// instead of its whole length (measured on the bench workload: 7.7 GB -> 0.x GB of DRAM reads per step).
template <int NC>
__device__ __forceinline__ void peak_notes(const float4* __restrict__ a4, const float4* __restrict__ b4, int la, int lb,
                                           const float* __restrict__ bma, const float* __restrict__ bmb,
                                           const adtfe_event* __restrict__ ev, int* __restrict__ peak_bits, int lane) {
    float ca[NC], cb[NC], best[NC];
#pragma unroll
    for (int i = 0; i < NC; ++i) {   // the same address in every lane: broadcast loads
        const float2 c = __ldg(reinterpret_cast<const float2*>(&ev[i].ca));
        ca[i] = c.x; cb[i] = c.y; best[i] = 0.0f;
    }
    const int nba = (la + kBlock - 1) / kBlock, nbb = (lb + kBlock - 1) / kBlock, nblk = max(nba, nbb);
    // the block with the largest bound for the first note
    float top = -1.0f;
    int top_k = 0;
    for (int k = lane; k < nblk; k += 32) {
        const float ub = mix_bound(k < nba ? __ldg(bma + k) : 0.0f, k < nbb ? __ldg(bmb + k) : 0.0f, ca[0], cb[0]);
        if (ub > top) { top = ub; top_k = k; }
    }
    const unsigned top_bits = __reduce_max_sync(0xffffffffu, __float_as_uint(fmaxf(top, 0.0f)));
    const unsigned holders = __ballot_sync(0xffffffffu, top >= 0.0f && __float_as_uint(top) == top_bits);
    const int k0 = holders ? __shfl_sync(0xffffffffu, top_k, __ffs((int)holders) - 1) : 0;
    if (nblk > 0) scan_block<NC>(a4, b4, la, lb, k0, lane, ca, cb, best);
    for (int j0 = 0; j0 < nblk; j0 += 32) {
        const int k = j0 + lane;
        const float A = k < nba ? __ldg(bma + k) : 0.0f, B = k < nbb ? __ldg(bmb + k) : 0.0f;
        bool cand = false;
#pragma unroll
        for (int i = 0; i < NC; ++i) cand |= mix_bound(A, B, ca[i], cb[i]) > best[i];
        unsigned todo = __ballot_sync(0xffffffffu, cand && k < nblk && k != k0);
        while (todo) {
            const int l = __ffs((int)todo) - 1;
            todo &= todo - 1u;
            const float Al = __shfl_sync(0xffffffffu, A, l), Bl = __shfl_sync(0xffffffffu, B, l);
            bool still = false;   // the lower bounds have risen since the ballot
#pragma unroll
            for (int i = 0; i < NC; ++i) still |= mix_bound(Al, Bl, ca[i], cb[i]) > best[i];
            if (still) scan_block<NC>(a4, b4, la, lb, j0 + l, lane, ca, cb, best);
        }
    }
#pragma unroll
    for (int i = 0; i < NC; ++i)
        if (lane == i) peak_bits[i] = (int)__float_as_uint(best[i]);   // an all-zero one-shot keeps 0: gain / 0 in the mixer
}

// One warp per group: the notes of one instrument in one segment (same two one-shots, one mixup each).  Work items with
// chunk != 0 (the per-span items of ABI <= 4 planners) are skipped.  The warp also resolves the bank lookups of its
// notes into ResolvedEvent records for the tile mixer.
__global__ void __launch_bounds__(kPeakWarps * 32) peak_kernel(
    const float* __restrict__ pcm, const float* __restrict__ bm, const int32_t* __restrict__ bm_off,
    const adtfe_event* __restrict__ events, const adtfe_peak_item* __restrict__ work, int n_items,
    ResolvedEvent* __restrict__ resolved, int* __restrict__ peak_bits) {
    const int w = blockIdx.x * kPeakWarps + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (w >= n_items) return;
    const adtfe_peak_item item = work[w];
    const int e0 = item.first_event, e1 = e0 + item.n_events;
    if (item.chunk != 0 || e0 >= e1) return;
    const int64_t a_off = item.a_off, b_off = item.b_off;
    const int la = item.la, lb = item.lb;
    const int2 ids = __ldg(reinterpret_cast<const int2*>(&events[e0].main_id));
    const float* bma = bm + __ldg(bm_off + ids.x);
    const float* bmb = bm + __ldg(bm_off + ids.y);
    for (int e = e0 + lane; e < e1; e += 32) {   // resolve the bank lookups once per note for the tile mixer
        const adtfe_event ev = events[e];
        ResolvedEvent r;
        r.a_off = a_off; r.b_off = b_off;
        r.la = min(la, ev.len); r.lb = min(lb, ev.len);
        r.start = ev.start; r.len = ev.len; r.ca = ev.ca; r.cb = ev.cb; r.gain = ev.gain; r.pad = 0;
        resolved[e] = r;
    }
    const float4* a4 = reinterpret_cast<const float4*>(pcm + a_off);
    const float4* b4 = reinterpret_cast<const float4*>(pcm + b_off);
    for (int c0 = e0; c0 < e1; c0 += kPeakNotes) {
        switch (min(kPeakNotes, e1 - c0)) {
            case 1: peak_notes<1>(a4, b4, la, lb, bma, bmb, events + c0, peak_bits + c0, lane); break;
            case 2: peak_notes<2>(a4, b4, la, lb, bma, bmb, events + c0, peak_bits + c0, lane); break;
            case 3: peak_notes<3>(a4, b4, la, lb, bma, bmb, events + c0, peak_bits + c0, lane); break;
            case 4: peak_notes<4>(a4, b4, la, lb, bma, bmb, events + c0, peak_bits + c0, lane); break;
            case 5: peak_notes<5>(a4, b4, la, lb, bma, bmb, events + c0, peak_bits + c0, lane); break;
            case 6: peak_notes<6>(a4, b4, la, lb, bma, bmb, events + c0, peak_bits + c0, lane); break;
            case 7: peak_notes<7>(a4, b4, la, lb, bma, bmb, events + c0, peak_bits + c0, lane); break;
            default: peak_notes<8>(a4, b4, la, lb, bma, bmb, events + c0, peak_bits + c0, lane); break;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Tile mixer, in two kernels.
//
// slice_kernel   one warp per output tile: lane i turns the tile's i-th note (host CSR tile -> events, the resolved
//                bank lookups and the peak of the peak pass) into a 32-byte TileSlice record - which 16-byte
//                granules of the two one-shots reach the tile, where they land relative to the tile, the covered
//                sample range and the two coefficients c * gain / peak.  Records are stored in tile order, so the
//                mixer fetches a tile's notes with ONE coalesced load instead of a three-deep dependent chain.
// mix_kernel     persistent ONE-WARP CTAs, nothing shared between warps.  A warp owns a tile of 2048 output samples
//                (64 per lane, sample lane + 32 j, in registers) and streams the tile's slices through its own ring
//                of kDepth shared-memory slots with TMA bulk copies (cp.async.bulk + one mbarrier per slot).  There
//                is no producer warp and no hand-off: the lane that owns a note issues that note's copies itself
//                (it holds the record in registers), the warp waits on the slot's mbarrier, adds the slice
//                `coef * x` into the accumulators - fixed event order, main then sub, no atomics - and the lane that
//                owns stream position c + kDepth refills the freed slot.  The slice stream runs ACROSS tiles: the
//                records of the next tile are fetched two tiles ahead (queue head -> tile_ptr -> records, one level
//                per tile), so while a tile's last slices are consumed the next tile's first ones are already in
//                flight and the ring never drains.
// Round 1's mixer (a producer warp feeding consumer warps through full/empty mbarriers) spent 407 warp instructions
// per slice, 96 of them the LDS + FFMA2 that do the work (ncu: 16 M of 91 M instructions mbarrier polling, 9 M
// branches, the rest descriptor traffic); this one spends ~130.  TMA cannot realign: a tensor-map copy with an
// element offset that is not a multiple of 16 bytes faults (tools/microbench/tma_shift.cu, profiles/r02_tma_shift.txt),
// so a slice lands with its source alignment and the lanes read it with conflict-free scalar LDS (lane-contiguous).
#ifndef ADTFE_MIX_DEPTH
#define ADTFE_MIX_DEPTH 2
#endif
#ifndef ADTFE_MIX_CTAS
#define ADTFE_MIX_CTAS 12
#endif
#ifndef ADTFE_MIX_WIDTH
#define ADTFE_MIX_WIDTH 2048
#endif
constexpr int kDepth = ADTFE_MIX_DEPTH;                  // ring slots (slices in flight or being consumed) per warp
constexpr int kMixCtasPerSm = ADTFE_MIX_CTAS;            // one-warp CTAs per SM (bounded by shared memory)
constexpr int kWidth = ADTFE_MIX_WIDTH;                  // output samples a warp owns at a time: a host tile or a part of it
constexpr int kSub = ADTFE_TILE / kWidth;                // warp tiles per host tile
constexpr int kStageFloats = kWidth + 32;                // >= kWidth + 3 alignment slack, bytes a multiple of 128
constexpr int kAcc = kWidth / 32;                        // samples per lane
static_assert(kSub * kWidth == ADTFE_TILE && (kAcc == 64 || kAcc == 32 || kAcc == 16), "warp tile width");

struct __align__(16) TileSlice {   // one (warp tile, note): 32 bytes
    uint32_t a_g4, b_g4;           // first 16-byte granule to copy, as float offset / 4 into the bank
    uint16_t na4, nb4;             // granules to copy (0: the source does not reach this tile)
    int16_t base;                  // tile sample i sits at slot index i + base
    uint16_t vlo;                  // tile samples vlo <= i < vhi_* are covered
    uint16_t vhi_a, vhi_b;
    uint32_t pad;
    float coef_a, coef_b;          // ca * gain / peak, cb * gain / peak
};
static_assert(sizeof(TileSlice) == 32, "TileSlice layout");

__global__ void __launch_bounds__(256) slice_kernel(const ResolvedEvent* __restrict__ resolved,
                                                    const int* __restrict__ peak_bits,
                                                    const int32_t* __restrict__ tile_ptr,
                                                    const int32_t* __restrict__ tile_events,
                                                    TileSlice* __restrict__ slices, int tiles_per_seg, int n_tiles) {
    const int tile = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (tile >= n_tiles) return;
    const int p0 = __ldg(tile_ptr + tile), p1 = __ldg(tile_ptr + tile + 1);
    const int lo = (tile % tiles_per_seg) * ADTFE_TILE;
    for (int p = p0 + lane; p < p1; p += 32) {
        const int e = __ldg(tile_events + p);
        const ResolvedEvent ev = resolved[e];
        // an all-zero one-shot has peak 0: gain/0 = inf (or NaN), and 0 * inf = NaN over the whole note - the
        // reference's o / 0 (synthetiser.py:225)
        const float scale = ev.gain / __int_as_float(peak_bits[e]);
        const int rel = lo - ev.start;                       // tile sample i is source sample i + rel
        const int r_lo = max(0, rel);
        const int r_hi_a = min(ev.la, rel + ADTFE_TILE), r_hi_b = min(ev.lb, rel + ADTFE_TILE);
        const int g_lo = r_lo & ~3;                          // storage is padded to 4 floats
        const int base = rel - g_lo, vlo = r_lo - rel;       // host tile: sample i sits at copy index i + base, i >= vlo
        const int vhi_a = max(r_hi_a - rel, 0), vhi_b = max(r_hi_b - rel, 0);
        const float coef_a = ev.ca * scale, coef_b = ev.cb * scale;
#pragma unroll
        for (int h = 0; h < kSub; ++h) {                     // the part of the note on warp tile h of the host tile
            const int H = h * kWidth;
            const int lo_h = max(vlo, H), hi_a = min(vhi_a, H + kWidth), hi_b = min(vhi_b, H + kWidth);
            const int gs = max(lo_h + base, 0) >> 2;         // first granule of the copy that the part needs
            TileSlice t;
            t.a_g4 = (uint32_t)((ev.a_off + g_lo) >> 2) + (uint32_t)gs;
            t.b_g4 = (uint32_t)((ev.b_off + g_lo) >> 2) + (uint32_t)gs;
            t.na4 = hi_a > lo_h ? (uint16_t)(((hi_a + base + 3) >> 2) - gs) : 0;
            t.nb4 = hi_b > lo_h ? (uint16_t)(((hi_b + base + 3) >> 2) - gs) : 0;
            t.base = (int16_t)(H + base - 4 * gs);           // warp-tile sample i' = i - H sits at slot index i' + base'
            t.vlo = (uint16_t)max(lo_h - H, 0);
            t.vhi_a = (uint16_t)max(hi_a - H, 0);
            t.vhi_b = (uint16_t)max(hi_b - H, 0);
            t.pad = 0;
            t.coef_a = coef_a;
            t.coef_b = coef_b;
            slices[(size_t)p * kSub + h] = t;
        }
    }
}

struct MixArgs {
    const float* pcm;
    const TileSlice* slices;
    const int32_t* tile_ptr;
    float* wav;
    float* tile_max;
    int* tile_counter;
    int64_t ld_wav;
    int32_t tiles_per_seg, n_tiles;
};

// The notes of (up to 32 of) one tile, one per lane, as the mixer keeps them in registers.
struct NoteRegs {
    uint32_t a_g4, b_g4;
    uint32_t n4;        // na4 | nb4 << 16
    uint32_t base_vlo;  // (uint16)base | vlo << 16
    uint32_t vhi;       // vhi_a | vhi_b << 16
    float coef_a, coef_b;
    int pos_a, pos_b;   // position of the lane's slices in the batch's stream (-1: not live)
};
struct Batch {
    NoteRegs r;
    unsigned m_a, m_b;  // lanes whose main / sub slice is live
    int n;              // slices in the batch
    int tile;           // < 0: no more work
    int p_next, p_end;  // records of the tile not yet in a batch: [p_next, p_end)
    int pre;            // slices of this batch already issued (behind the previous batch's)
};

__device__ __forceinline__ void load_records(const TileSlice* __restrict__ slices, int p0, int p1, int h, int lane,
                                             uint4& q0, uint4& q1) {
    q0 = make_uint4(0u, 0u, 0u, 0u);
    q1 = make_uint4(0u, 0u, 0u, 0u);
    if (p0 + lane < p1) {
        const uint4* q = reinterpret_cast<const uint4*>(slices + (size_t)(p0 + lane) * kSub + h);
        q0 = __ldg(q);
        q1 = __ldg(q + 1);
    }
}

__device__ __forceinline__ void decode_batch(Batch& b, const uint4& q0, const uint4& q1, int lane) {
    b.r.a_g4 = q0.x; b.r.b_g4 = q0.y; b.r.n4 = q0.z; b.r.base_vlo = q0.w;
    b.r.vhi = q1.x; b.r.coef_a = __uint_as_float(q1.z); b.r.coef_b = __uint_as_float(q1.w);
    const bool live_a = (q0.z & 0xffffu) != 0u, live_b = (q0.z >> 16) != 0u;
    b.m_a = __ballot_sync(0xffffffffu, live_a);
    b.m_b = __ballot_sync(0xffffffffu, live_b);
    const unsigned below = (1u << lane) - 1u;
    const int before = __popc(b.m_a & below) + __popc(b.m_b & below);
    b.r.pos_a = live_a ? before : -1;
    b.r.pos_b = live_b ? before + (live_a ? 1 : 0) : -1;
    b.n = __popc(b.m_a) + __popc(b.m_b);
    b.pre = 0;
}

// the lane that owns stream position `pos` of batch `b` (global sequence number seq) starts the copy into its slot
__device__ __forceinline__ void issue_slice(const MixArgs& a, const NoteRegs& r, bool sub, unsigned seq, float* ring,
                                            uint64_t* bars) {
    const unsigned slot = seq % (unsigned)kDepth;
    const uint32_t bytes = (sub ? (r.n4 >> 16) : (r.n4 & 0xffffu)) * 16u;
    const float* src = a.pcm + (size_t)(sub ? r.b_g4 : r.a_g4) * 4u;
    // The slot's previous contents were read by loads whose values have been consumed (the FMAs ran) before the
    // __syncwarp that precedes this call, so the copy cannot overtake them; no proxy fence is needed for that
    // read-then-overwrite order (it compiles to MEMBAR.ALL.CTA, which would also wait for the tile's global stores).
    mbar_expect_tx(bars + slot, bytes);
    bulk_g2s(ring + slot * kStageFloats, src, bytes, bars + slot);
}

// issue the stream positions [from, to) of batch b (b's position 0 has global sequence number seq0)
__device__ __forceinline__ void issue_range(const MixArgs& a, const Batch& b, int from, int to, unsigned seq0, float* ring,
                                            uint64_t* bars) {
    if (b.r.pos_a >= from && b.r.pos_a < to) issue_slice(a, b.r, false, seq0 + (unsigned)b.r.pos_a, ring, bars);
    if (b.r.pos_b >= from && b.r.pos_b < to) issue_slice(a, b.r, true, seq0 + (unsigned)b.r.pos_b, ring, bars);
}

__global__ void __launch_bounds__(32, kMixCtasPerSm) mix_kernel(const MixArgs a) {
    extern __shared__ __align__(128) unsigned char mix_smem[];
    float* ring = reinterpret_cast<float*>(mix_smem);                                  // kDepth * kStageFloats
    uint64_t* bars = reinterpret_cast<uint64_t*>(ring + kDepth * kStageFloats);
    const int lane = threadIdx.x;
    if (lane < kDepth) mbar_init(bars + lane, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();

    // ---- metadata pipeline, one level per tile: queue head -> tile_ptr -> records -> decoded batch.  Every level's
    // loads are issued one tile before their results are used, so none of the four round trips is ever waited for.
    int ticket = 0;                         // lane 0: the queue head taken for the tile after id_c (not yet broadcast)
    if (lane == 0) ticket = atomicAdd(a.tile_counter, 1);
    int id_c = -1, pc0 = 0, pc1 = 0;        // tile whose tile_ptr entries are requested
    int id_b = -1, pb0 = 0, pb1 = 0;        // tile whose records are requested
    bool drained = false;                   // the queue is empty: no more tickets are taken
    uint4 q0 = make_uint4(0u, 0u, 0u, 0u), q1 = q0;
    Batch cur, nxt;
    nxt.tile = -1; nxt.n = 0; nxt.pre = 0; nxt.p_next = nxt.p_end = 0; nxt.m_a = nxt.m_b = 0u;
    nxt.r.pos_a = nxt.r.pos_b = -1;
    // advance the pipeline by one tile: nxt <- decoded records of id_b; records of id_c requested; pointers of the
    // tile behind the ticket requested; a new ticket taken
    auto advance = [&]() {
        nxt.tile = id_b;
        nxt.p_next = min(pb0 + 32, pb1);
        nxt.p_end = pb1;
        decode_batch(nxt, q0, q1, lane);
        id_b = id_c; pb0 = pc0; pb1 = pc1;
        if (id_b >= 0) load_records(a.slices, pb0, min(pb0 + 32, pb1), id_b % kSub, lane, q0, q1);
        else { q0 = make_uint4(0u, 0u, 0u, 0u); q1 = q0; }
        id_c = -1;
        if (!drained) {
            const int t = __shfl_sync(0xffffffffu, ticket, 0);
            if (t < a.n_tiles * kSub) {
                id_c = t;                   // warp tile t: part t % kSub of host tile t / kSub
                pc0 = __ldg(a.tile_ptr + t / kSub);
                pc1 = __ldg(a.tile_ptr + t / kSub + 1);
                if (lane == 0) ticket = atomicAdd(a.tile_counter, 1);
            } else {
                drained = true;
            }
        }
    };
    advance();   // id_c <- first tile, its pointers requested
    advance();   // id_b <- first tile, its records requested
    advance();   // nxt  <- first tile
    unsigned seq = 0;   // slices consumed so far by this warp: slot = seq % kDepth, parity = (seq / kDepth) & 1
    if (nxt.tile >= 0) {
        nxt.pre = min(nxt.n, kDepth);
        issue_range(a, nxt, 0, nxt.pre, seq, ring, bars);
    }

    float2 acc2[kAcc / 2];
#pragma unroll
    for (int k = 0; k < kAcc / 2; ++k) acc2[k] = make_float2(0.0f, 0.0f);

    for (;;) {
        cur = nxt;
        if (cur.tile < 0) break;
        const bool last_batch = cur.p_next >= cur.p_end;   // the tile's notes fit this batch (<= 32: the usual case)
        if (last_batch) {
            advance();                                      // nxt: the next tile (its copies start behind cur's)
        } else {                                            // more than 32 notes on the tile: the rest follows, fetched now
            uint4 r0, r1;
            load_records(a.slices, cur.p_next, min(cur.p_next + 32, cur.p_end), cur.tile % kSub, lane, r0, r1);
            Batch more;
            more.tile = cur.tile; more.p_next = min(cur.p_next + 32, cur.p_end); more.p_end = cur.p_end;
            decode_batch(more, r0, r1, lane);
            // the prefetched next tile keeps waiting in (id_b, q0, q1): park it by not advancing
            nxt = more;
        }
        // stream: cur's slices 0 .. cur.n-1 (the first cur.pre already issued), then nxt's
        const int n_x = cur.n, n_y = nxt.tile >= 0 ? nxt.n : 0;
        const unsigned seq0 = seq;
        {   // fill the ring: positions cur.pre .. kDepth-1 of the stream
            issue_range(a, cur, cur.pre, min(n_x, kDepth), seq0, ring, bars);
            if (n_x < kDepth) issue_range(a, nxt, 0, min(n_y, kDepth - n_x), seq0 + (unsigned)n_x, ring, bars);
        }
        int c = 0;   // stream position being consumed
        unsigned todo = cur.m_a | cur.m_b;
        while (todo) {
            const int e = __ffs((int)todo) - 1;
            todo &= todo - 1u;
            const uint32_t base_vlo = __shfl_sync(0xffffffffu, cur.r.base_vlo, e);
            const uint32_t vhi2 = __shfl_sync(0xffffffffu, cur.r.vhi, e);
            const float coef_a = __shfl_sync(0xffffffffu, cur.r.coef_a, e);
            const float coef_b = __shfl_sync(0xffffffffu, cur.r.coef_b, e);
            const int base = (int)(int16_t)(base_vlo & 0xffffu), vlo = (int)(base_vlo >> 16);
#pragma unroll 1
            for (int sub = 0; sub < 2; ++sub) {
                if (!(((sub ? cur.m_b : cur.m_a) >> e) & 1u)) continue;
                const float coef = sub ? coef_b : coef_a;
                const int vhi = (int)(sub ? (vhi2 >> 16) : (vhi2 & 0xffffu));
                const unsigned slot = seq % (unsigned)kDepth;
                mbar_wait(bars + slot, (seq / (unsigned)kDepth) & 1u);   // (a suspend-time hint measures the same)
                const float* src = ring + slot * kStageFloats + base + lane;
                if (vlo == 0 && vhi == kWidth) {
                    const float2 c2 = make_float2(coef, coef);
#pragma unroll
                    for (int k = 0; k < kAcc / 2; ++k)
                        acc2[k] = __ffma2_rn(make_float2(src[64 * k], src[64 * k + 32]), c2, acc2[k]);
                } else {
                    // A note starts or ends inside the tile: this lane's samples lane + 32 j lie inside [vlo, vhi) for
                    // jlo <= j < jhi.  One bit mask per half instead of two compares per sample; samples outside the
                    // note stay untouched even when coef is inf / NaN.
                    const int jlo = (max(vlo - lane, 0) + 31) >> 5, jhi = min((max(vhi - lane, 0) + 31) >> 5, kAcc);
                    auto below = [](int j) { return j >= 32 ? 0xffffffffu : (j <= 0 ? 0u : (1u << j) - 1u); };
                    const unsigned mask_lo = below(jhi) & ~below(jlo);
                    const unsigned mask_hi = below(jhi - 32) & ~below(jlo - 32);
#pragma unroll
                    for (int j = 0; j < kAcc; ++j) {
                        if (((j < 32 ? mask_lo : mask_hi) >> (j & 31)) & 1u) {
                            if (j & 1) acc2[j >> 1].y = fmaf(src[32 * j], coef, acc2[j >> 1].y);
                            else acc2[j >> 1].x = fmaf(src[32 * j], coef, acc2[j >> 1].x);
                        }
                    }
                }
                __syncwarp();   // every lane has read the slot: it can be refilled
                const int t = c + kDepth;   // the stream position that takes the freed slot
                if (t < n_x) issue_range(a, cur, t, t + 1, seq0, ring, bars);
                else if (t - n_x < n_y) issue_range(a, nxt, t - n_x, t - n_x + 1, seq0 + (unsigned)n_x, ring, bars);
                ++c;
                ++seq;
            }
        }
        nxt.pre = min(n_y, kDepth);
        if (!last_batch) continue;   // the tile's remaining notes come with the next batch

        // ---- the tile is complete: write it (not yet normalised) and publish its |max|
        {
            const int tile = cur.tile;   // warp tile: part tile % kSub of host tile tile / kSub
            const int seg = tile / (a.tiles_per_seg * kSub), lo = (tile - seg * a.tiles_per_seg * kSub) * kWidth;
            float* row = a.wav + (int64_t)seg * a.ld_wav + lo + lane;
            // |max| on the float bits: non-negative floats order like unsigned integers and a NaN's bits lie above
            // infinity's, so the unsigned maximum propagates NaN like torch.max; one REDUX per tile
            unsigned mb = 0u;
            if (lo + kWidth <= a.ld_wav) {
#pragma unroll
                for (int k = 0; k < kAcc / 2; ++k) {
                    row[64 * k] = acc2[k].x;
                    row[64 * k + 32] = acc2[k].y;
                    mb = max(mb, max(__float_as_uint(fabsf(acc2[k].x)), __float_as_uint(fabsf(acc2[k].y))));
                }
            } else {
                const int room = (int)(a.ld_wav - lo) - lane;   // samples of this lane's column that exist
#pragma unroll
                for (int k = 0; k < kAcc / 2; ++k) {
                    if (64 * k < room) row[64 * k] = acc2[k].x;
                    if (64 * k + 32 < room) row[64 * k + 32] = acc2[k].y;
                    mb = max(mb, max(__float_as_uint(fabsf(acc2[k].x)), __float_as_uint(fabsf(acc2[k].y))));
                }
            }
            mb = __reduce_max_sync(0xffffffffu, mb);
            if (lane == 0) a.tile_max[tile] = __uint_as_float(mb);
#pragma unroll
            for (int k = 0; k < kAcc / 2; ++k) acc2[k] = make_float2(0.0f, 0.0f);
        }
    }
}


// Row normalisation (wav / peak * max_volume, synthetiser.py:142-144,156): one CTA per tile; the
// segment peak is the max of its tile maxima.  An all-zero mix gives NaN (the reference's 0/0),
// samples beyond the segment's length stay exact zeros (collate_fn pads with 0.0).
// Measured alternatives (both bit-identical, both slower on B200): the warp that completes a row (per-row ticket)
// normalising it inside the tile mixer while it is still in L2 - 7.8 ms render against 7.0 ms, the row pass is a
// latency-bound loop behind the mixer's own L2 traffic (~150 us per row and warp), and with small chunks the kernel
// ends on a tail of single warps; round 1 measured the same for its CTA-wide form.
#ifndef ADTFE_NORM_THREADS
#define ADTFE_NORM_THREADS 256
#endif
constexpr int kNormThreads = ADTFE_NORM_THREADS;
// `segments`, `tile_max` and `wav` cover the whole plan.  Per chunk (fx_rows == nullptr) the rows are seg0 + blockIdx.x /
// tiles_per_seg, minus the rows marked in seg_fx: those still have the FX chain ahead of them and are normalised at
// the end of the render, by a launch over the plan's FX records (fx_rows != nullptr).
__global__ void fx_mark_kernel(const adtfe_fx* __restrict__ fx_rows, int n_rows, int* __restrict__ seg_fx) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n_rows) seg_fx[fx_rows[r].seg] = 1;
}

__global__ void __launch_bounds__(kNormThreads) normalise_kernel(const adtfe_segment* __restrict__ segments,
                                                                 const float* __restrict__ tile_max, int tiles_per_seg,
                                                                 int64_t ld_wav, float* __restrict__ wav, int seg0,
                                                                 const int* __restrict__ seg_fx,
                                                                 const adtfe_fx* __restrict__ fx_rows) {
    const int tid = threadIdx.x;
    const int row_id = blockIdx.x / tiles_per_seg, lo = (blockIdx.x - row_id * tiles_per_seg) * ADTFE_TILE;
    const int seg = fx_rows ? fx_rows[row_id].seg : seg0 + row_id;
    float* row = wav + (int64_t)seg * ld_wav + lo;  // ld_wav is a multiple of 4 and the base is 16-byte aligned
    float4* row4 = reinterpret_cast<float4*>(row);
    // The tile's samples are requested FIRST, whatever the row turns out to be: the segment record and the tile maxima
    // are a chain of two dependent loads, and a CTA that waits for them before it asks for its 8 KB has nothing in
    // flight for half of its life.  (Reads within the row's pitch are always inside the matrix.)
    constexpr int kPer = ADTFE_TILE / 4 / kNormThreads;
    static_assert(kPer * kNormThreads * 4 == ADTFE_TILE, "normalise_kernel: threads per tile");
    const int n4_row = (int)min((int64_t)ADTFE_TILE, ld_wav - lo) >> 2;
    float4 v[kPer];
#pragma unroll
    for (int k = 0; k < kPer; ++k) {
        const int i = tid + k * kNormThreads;
        v[k] = i < n4_row ? row4[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (!fx_rows && seg_fx[seg]) return;
    const adtfe_segment sg = segments[seg];
    if ((sg.flags & 1) == 0 || lo >= sg.len) return;   // 0: empty row; ADTFE_SEG_RAW: the caller wants the raw mix
    float peak = 0.0f;
    // one maximum per warp tile of the mixer (kSub per host tile)
    for (int t = tid & 31; t < tiles_per_seg * kSub; t += 32)
        peak = nan_max(peak, __ldg(tile_max + (size_t)seg * tiles_per_seg * kSub + t));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) peak = nan_max(peak, __shfl_xor_sync(0xffffffffu, peak, o));
    const float vol = sg.max_volume;
    const int n = min(ADTFE_TILE, sg.len - lo);
    const int n4 = n >> 2;
    // v / peak * vol with the division by an invariant done the way the hardware sequence does it: q = v*r,
    // one residual correction (correctly rounded for normal operands, i.e. everything a mix can hold), r = 1/peak
    // rounded to nearest once per thread.  peak = 0 (all-zero mix) or NaN still gives NaN everywhere.
    const float r = __frcp_rn(peak);
    auto norm = [&](float x) {
        const float q = __fmul_rn(x, r);
        const float rem = __fmaf_rn(-q, peak, x);
        return __fmul_rn(__fmaf_rn(rem, r, q), vol);
    };
#pragma unroll
    for (int k = 0; k < kPer; ++k) {
        const int i = tid + k * kNormThreads;
        if (i < n4) row4[i] = make_float4(norm(v[k].x), norm(v[k].y), norm(v[k].z), norm(v[k].w));
    }
    for (int i = 4 * n4 + tid; i < n; i += kNormThreads) row[i] = norm(row[i]);
}

constexpr int kFxStreams = 4;   // streams the chunks' reverb launches rotate over (bank streams 3 .. 6)
static_assert(kBankStreams >= 3 + kFxStreams, "bank streams");

static size_t mix_smem_bytes() { return (size_t)kDepth * kStageFloats * 4 + kDepth * 8 + 64; }

}  // namespace adtfe

using namespace adtfe;

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

// Called by adtfe_bank_create on the bank's device (cudaFuncSetAttribute is per device and idempotent): the kernels'
// dynamic shared memory is opted in once per handle, not through process-wide state on the render path.
// Block maxima of the bank for the peak pass (called once by adtfe_bank_create, on the bank's device).
int adtfe::bank_build_blockmax(adtfe_bank* b, const int32_t* lengths_host) {
    std::vector<int32_t> off((size_t)b->n + 1, 0);
    for (int32_t i = 0; i < b->n; ++i) off[i + 1] = off[i] + (lengths_host[i] + kBlock - 1) / kBlock;
    ADTFE_CUDA(cudaMalloc((void**)&b->bm_off, off.size() * 4));
    ADTFE_CUDA(cudaMemcpy(b->bm_off, off.data(), off.size() * 4, cudaMemcpyHostToDevice));
    ADTFE_CUDA(cudaMalloc((void**)&b->blockmax, (size_t)std::max(off.back(), 1) * 4));
    if (b->n > 0) {
        blockmax_kernel<<<b->n, 256>>>(b->pcm, b->offsets, b->lengths, b->bm_off, b->blockmax);
        ADTFE_CUDA(cudaGetLastError());
        ADTFE_CUDA(cudaDeviceSynchronize());
    }
    return ADTFE_OK;
}

int adtfe::mixer_prepare_device() {
    ADTFE_CUDA(cudaFuncSetAttribute(mix_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mix_smem_bytes()));
    return ADTFE_OK;
}

extern "C" size_t adtfe_render_workspace_bytes(int32_t n_events, int32_t n_seg, int32_t tiles_per_seg,
                                               int32_t n_tile_events) {
    if (n_events < 0 || n_seg < 0 || tiles_per_seg < 0 || n_tile_events < 0) return 0;
    // resolved events | peak bits, queue heads, FX row marks (one zeroed block) | tile maxima | per-tile slice records
    return align256((size_t)n_events * sizeof(ResolvedEvent)) + align256((size_t)n_events * 4 + (size_t)n_seg * 8 + 4) +
           align256((size_t)n_seg * tiles_per_seg * kSub * 4) +
           align256((size_t)n_tile_events * kSub * sizeof(TileSlice)) + 256;
}

extern "C" int adtfe_render(const adtfe_bank* bank, const adtfe_plan* plan, float* wav_out_dev, void* workspace_dev,
                            size_t workspace_bytes, void* stream) {
    return render_impl(bank, plan, wav_out_dev, workspace_dev, workspace_bytes, stream);
}

int adtfe::render_impl(const adtfe_bank* bank, const adtfe_plan* plan, float* wav_out_dev, void* workspace_dev,
                       size_t workspace_bytes, void* stream, const ChunkHook* hook, const int** seg_fx_out) {
    ADTFE_REQUIRE(bank && plan, ADTFE_ERR_BAD_ARG, "adtfe_render: null bank or plan");
    ADTFE_REQUIRE(plan->n_seg >= 0 && plan->n_events >= 0 && plan->n_peak_work >= 0 && plan->tiles_per_seg >= 0 &&
                      plan->n_tile_events >= 0 && plan->n_fx >= 0,
                  ADTFE_ERR_BAD_ARG, "adtfe_render: negative count");
    ADTFE_REQUIRE(plan->n_fx == 0 || plan->fx_dev, ADTFE_ERR_BAD_ARG, "adtfe_render: n_fx > 0 without fx_dev");
    if (plan->n_seg == 0 || plan->tiles_per_seg == 0) return ADTFE_OK;
    ADTFE_REQUIRE(plan->ld_wav > 0 && plan->ld_wav % 4 == 0 &&
                      plan->ld_wav <= (int64_t)plan->tiles_per_seg * ADTFE_TILE,
                  ADTFE_ERR_BAD_ARG, "adtfe_render: ld_wav %lld must be a positive multiple of 4 within %d tiles",
                  (long long)plan->ld_wav, plan->tiles_per_seg);
    ADTFE_REQUIRE(wav_out_dev && ((uintptr_t)wav_out_dev & 15) == 0 && plan->segments_dev && plan->tile_ptr_dev,
                  ADTFE_ERR_BAD_ARG, "adtfe_render: null or misaligned buffer (wav_out must be 16-byte aligned)");
    ADTFE_REQUIRE(plan->n_events == 0 || (plan->events_dev && plan->tile_events_dev && plan->peak_work_dev &&
                                          bank->pcm && plan->n_peak_work > 0),
                  ADTFE_ERR_BAD_ARG, "adtfe_render: null event buffers");
    ADTFE_REQUIRE(bank->total / 4 < (1ll << 32), ADTFE_ERR_UNSUPPORTED, "adtfe_render: bank larger than 64 GB");
    const size_t need = adtfe_render_workspace_bytes(plan->n_events, plan->n_seg, plan->tiles_per_seg,
                                                     plan->n_tile_events);
    ADTFE_REQUIRE(workspace_dev && workspace_bytes >= need, ADTFE_ERR_WORKSPACE,
                  "adtfe_render: workspace %zu B < %zu B", workspace_bytes, need);
    // chunk boundaries: the plan's, or one chunk covering everything
    const adtfe_chunk whole[2] = {{0, 0, 0, 0}, {plan->n_seg, plan->n_events, plan->n_peak_work, plan->n_fx}};
    const adtfe_chunk* ch = whole;
    int n_chunks = 1;
    if (plan->chunks_host && plan->n_chunks > 0) {
        ch = plan->chunks_host;
        n_chunks = plan->n_chunks;
        ADTFE_REQUIRE(n_chunks <= plan->n_seg && ch[0].seg == 0 && ch[0].event == 0 && ch[0].peak_work == 0 &&
                          ch[0].fx_row == 0 && ch[n_chunks].seg == plan->n_seg && ch[n_chunks].event == plan->n_events &&
                          ch[n_chunks].peak_work == plan->n_peak_work && ch[n_chunks].fx_row == plan->n_fx,
                      ADTFE_ERR_BAD_ARG, "adtfe_render: chunk boundaries do not cover the plan");
        for (int c = 0; c < n_chunks; ++c)
            ADTFE_REQUIRE(ch[c + 1].seg > ch[c].seg && ch[c + 1].event >= ch[c].event &&
                              ch[c + 1].peak_work >= ch[c].peak_work && ch[c + 1].fx_row >= ch[c].fx_row,
                          ADTFE_ERR_BAD_ARG, "adtfe_render: chunk %d is empty or out of order", c);
    }
    cudaStream_t user = (cudaStream_t)stream;
    char* ws = (char*)(((uintptr_t)workspace_dev + 255) & ~(uintptr_t)255);
    ResolvedEvent* resolved = (ResolvedEvent*)ws;
    int* peak_bits = (int*)(ws + align256((size_t)plan->n_events * sizeof(ResolvedEvent)));
    int* counters = peak_bits + plan->n_events;  // one work-queue head per chunk (n_chunks <= n_seg)
    int* seg_fx = counters + plan->n_seg + 1;    // 1: the row has an FX record (normalised after the FX chain)
    if (seg_fx_out) *seg_fx_out = seg_fx;
    float* tile_max = (float*)((char*)peak_bits + align256((size_t)plan->n_events * 4 + (size_t)plan->n_seg * 8 + 4));
    TileSlice* slices = (TileSlice*)((char*)tile_max + align256((size_t)plan->n_seg * plan->tiles_per_seg * kSub * 4));
    // zero the peaks and the queue heads once, then fork the chunks over the bank's streams
    ADTFE_CUDA(cudaMemsetAsync(peak_bits, 0, ((size_t)plan->n_events + 2 * (size_t)plan->n_seg + 1) * 4, user));
    if (plan->n_fx > 0) {
        fx_mark_kernel<<<(plan->n_fx + 127) / 128, 128, 0, user>>>(plan->fx_dev, plan->n_fx, seg_fx);
        ADTFE_CUDA(cudaGetLastError());
    }
    const bool fork = n_chunks > 1 && bank->n_streams >= 3 + kFxStreams;
    std::unique_lock<std::mutex> lock(bank->mu, std::defer_lock);
    // A chunked plan runs as a three-stage software pipeline over the bank's internal streams - stage 0: peaks and
    // slice records, stage 1: tile mixer, stage 2: row normalisation - chunk c's stage k waiting (event) for its
    // stage k-1, so that the peak pass of chunk c+2 (HBM reads), the mixer of chunk c+1 (L2 -> SM traffic) and the
    // normalisation of chunk c (HBM read + write) can be on the GPU together.  Measured: the same 7.0 ms per step as
    // whole chunks on four streams (round 1) - the kernels are bound by what an SM can keep in flight, and that
    // they share - with one stream and two events fewer per chunk.
    // FX rows take further streams.  Both FX kernels are recursions over a row's samples: a launch takes about the same
    // time whatever its number of rows (reverb 1.3 ms up to 8 rows per SM, dynamics 0.8 ms), so the FX launches of the
    // chunks must not queue behind each other (one stream for all of them cost 25.7 ms instead of 3.3 ms for 64 batches
    // of the stock FX configuration): the reverb of a chunk's FX rows goes behind the chunk's mixer on one of kFxStreams
    // streams in turn, and - once the last chunk's reverb is enqueued - ONE dynamics launch and one normalisation cover
    // all the plan's FX rows.
    cudaStream_t s_peak = user, s_mix = user, s_norm = user, s_fx[kFxStreams];
    for (cudaStream_t& st : s_fx) st = user;
    if (fork) {
        lock.lock();
        s_peak = bank->streams[0]; s_mix = bank->streams[1]; s_norm = bank->streams[2];
        for (int k = 0; k < kFxStreams; ++k) s_fx[k] = bank->streams[3 + k];
        ADTFE_CUDA(cudaEventRecord(bank->fork_event, user));
        for (int k = 0; k < 3 + (plan->n_fx > 0 ? kFxStreams : 0); ++k)
            ADTFE_CUDA(cudaStreamWaitEvent(bank->streams[k], bank->fork_event, 0));
    }
    float* mix_out = wav_out_dev;
    const int tps = plan->tiles_per_seg;
    for (int c = 0; c < n_chunks; ++c) {
        const int s0 = ch[c].seg, n_seg = ch[c + 1].seg - s0;
        const int pw0 = ch[c].peak_work, n_pw = ch[c + 1].peak_work - pw0;
        if (n_pw > 0) {
            trace_open("peak", c, s_peak);
            peak_kernel<<<(n_pw + kPeakWarps - 1) / kPeakWarps, kPeakWarps * 32, 0, s_peak>>>(
                bank->pcm, bank->blockmax, bank->bm_off, plan->events_dev, plan->peak_work_dev + pw0, n_pw, resolved,
                peak_bits);
            trace_close(s_peak);
            ADTFE_CUDA(cudaGetLastError());
        }
        MixArgs a;
        a.pcm = bank->pcm; a.slices = slices;   // indexed like tile_events: the tile_ptr entries are global positions
        a.tile_ptr = plan->tile_ptr_dev + (size_t)s0 * tps;
        a.wav = mix_out + (size_t)s0 * plan->ld_wav;
        a.tile_max = tile_max + (size_t)s0 * tps * kSub; a.tile_counter = counters + c; a.ld_wav = plan->ld_wav;
        a.tiles_per_seg = tps; a.n_tiles = n_seg * tps;
        if (ch[c + 1].event > ch[c].event) {
            trace_open("slice", c, s_peak);
            slice_kernel<<<(a.n_tiles + 7) / 8, 256, 0, s_peak>>>(resolved, peak_bits, a.tile_ptr, plan->tile_events_dev,
                                                                  slices, tps, a.n_tiles);
            trace_close(s_peak);
            ADTFE_CUDA(cudaGetLastError());
        }
        if (fork) {
            cudaEvent_t e = bank->stage_events[0][c % kStageEvents];
            ADTFE_CUDA(cudaEventRecord(e, s_peak));
            ADTFE_CUDA(cudaStreamWaitEvent(s_mix, e, 0));
        }
        const int grid = std::min(a.n_tiles * kSub, kMixCtasPerSm * bank->sm_count);
        trace_open("mix", c, s_mix);
        mix_kernel<<<grid, 32, mix_smem_bytes(), s_mix>>>(a);
        trace_close(s_mix);
        ADTFE_CUDA(cudaGetLastError());
        const int fx0 = ch[c].fx_row, n_fx = ch[c + 1].fx_row - fx0;
        if (fork) {
            cudaEvent_t e = bank->stage_events[1][c % kStageEvents];
            ADTFE_CUDA(cudaEventRecord(e, s_mix));
            ADTFE_CUDA(cudaStreamWaitEvent(s_norm, e, 0));
            if (n_fx > 0) ADTFE_CUDA(cudaStreamWaitEvent(s_fx[c % kFxStreams], e, 0));
        }
        trace_open("normalise", c, s_norm);
        normalise_kernel<<<a.n_tiles, kNormThreads, 0, s_norm>>>(plan->segments_dev, tile_max, tps, plan->ld_wav, mix_out,
                                                                 s0, seg_fx, nullptr);
        trace_close(s_norm);
        ADTFE_CUDA(cudaGetLastError());
        if (hook) {
            const int rc = hook->fn(hook->ctx, c, s0, n_seg, seg_fx, s_norm);
            if (rc != ADTFE_OK) return rc;
        }
        if (n_fx > 0) {   // the chunk's FX rows: raw mix -> reverb now; dynamics, the new row peak and normalise below
            const int rc = fx_reverb_launch(plan, fx0, n_fx, mix_out, s_fx[c % kFxStreams]);
            if (rc != ADTFE_OK) return rc;
        }
    }
    if (plan->n_fx > 0) {
        if (fork) {   // every reverb before the dynamics
            for (int k = 1; k < kFxStreams; ++k) {
                ADTFE_CUDA(cudaEventRecord(bank->join_events[3 + k], s_fx[k]));
                ADTFE_CUDA(cudaStreamWaitEvent(s_fx[0], bank->join_events[3 + k], 0));
            }
        }
        const int rc = fx_dynamics_launch(plan, 0, plan->n_fx, mix_out, tile_max, tps * kSub, s_fx[0]);
        if (rc != ADTFE_OK) return rc;
        trace_open("normalise_fx", 0, s_fx[0]);
        normalise_kernel<<<plan->n_fx * tps, kNormThreads, 0, s_fx[0]>>>(plan->segments_dev, tile_max, tps, plan->ld_wav,
                                                                         mix_out, 0, seg_fx, plan->fx_dev);
        trace_close(s_fx[0]);
        ADTFE_CUDA(cudaGetLastError());
    }
    if (fork) {   // the last stage and the FX chain finish last: two joins
        ADTFE_CUDA(cudaEventRecord(bank->join_events[2], s_norm));
        ADTFE_CUDA(cudaStreamWaitEvent(user, bank->join_events[2], 0));
        if (plan->n_fx > 0) {
            ADTFE_CUDA(cudaEventRecord(bank->join_events[3], s_fx[0]));
            ADTFE_CUDA(cudaStreamWaitEvent(user, bank->join_events[3], 0));
        }
    }
    return ADTFE_OK;
}
