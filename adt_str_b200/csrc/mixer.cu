// Overlap-add tile mixer (sm_100a): peaks -> tiles -> normalise.
//
// Replaces the audio arithmetic of SynthDrum.__call__ (reference modules/synthetiser.py):
//   drum_rendering 214-239:  o = (1-mixup)*a + mixup*b ; o /= max|o| ; o *= vol ;
//                            track[start : start+len] += o[:len]
//   instrument_mixer 149-156: wav = sum_i track_i * gain_i ; wav / max|wav| * max_volume
// and the zero padding of collate_fn (data_modules/train_dataset.py:53).
//
// The two data-dependent maxima make it three short kernels per chunk of segments:
//   1. peak_kernel      one CTA per (segment, instrument, 4096-sample chunk): both one-shots are
//                       read once and max|ca*a + cb*b| is taken for every note of that
//                       instrument at the same time (each note has its own mixup); warps and chunks
//                       meet in an atomicMax on the float bits (order-independent); chunk 0 also
//                       resolves the bank lookups into ResolvedEvent records.
//   2. mix_kernel       persistent CTAs take 2048-sample output tiles from a queue; a producer warp
//                       streams the tile's one-shot slices through a shared-memory ring with TMA bulk
//                       copies, consumer warps add them in event order (the host-built CSR
//                       tile -> events) into registers, so the result is deterministic and needs no
//                       atomics; writes the raw tile and its |max|.
//   3. normalise_kernel row peak = max of the tile maxima; wav / peak * max_volume in place
//                       (an all-zero mix gives NaN, like the reference's 0/0).
#include <algorithm>
#include <mutex>

#include "common.cuh"

namespace adtfe {

#ifndef ADTFE_PEAK_THREADS
#define ADTFE_PEAK_THREADS 128   // 128 x 94 registers: 6 % faster render than 256 x 64; 64 = two CTAs per work item
#endif
constexpr int kPeakThreads = ADTFE_PEAK_THREADS;
constexpr int kPeakChunk = 8;   // notes of one instrument handled per sweep over the one-shots
constexpr int kPeakSpan = ADTFE_PEAK_SPAN;  // samples of the mixed one-shot per peak work item
constexpr int kPeakIters = 8;                                 // float4 per thread per one-shot
constexpr int kPeakCtaSpan = kPeakIters * 4 * kPeakThreads;   // samples of the mixed one-shot scanned by one CTA
constexpr int kPeakSplit = kPeakSpan / kPeakCtaSpan;          // CTAs per work item (1 at 128 threads, 2 at 64)
static_assert(kPeakSplit >= 1 && kPeakSplit * kPeakCtaSpan == kPeakSpan, "peak span must be a multiple of 32 * threads");

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

// max|ca*a + cb*b| over the float4s held in registers, for NC notes at once: per pair of samples two
// packed fp32 instructions and one three-input FMNMX; the warp's maximum by one REDUX on the float bits
// (non-negative floats order like unsigned integers), published with one global atomicMax per warp and note
// (no shared-memory staging, no CTA barrier: barriers were 12 % of the kernel's stalls).
template <int NC>
__device__ __forceinline__ void peak_chunk(const float4 (&va)[kPeakIters], const float4 (&vb)[kPeakIters],
                                           const adtfe_event* __restrict__ ev, int* __restrict__ peak_bits, int tid) {
    float ca[NC], cb[NC], m[NC];
#pragma unroll
    for (int i = 0; i < NC; ++i) {   // the same address in every thread: broadcast loads that hit L1
        const float2 c = __ldg(reinterpret_cast<const float2*>(&ev[i].ca));
        ca[i] = c.x; cb[i] = c.y; m[i] = 0.0f;
    }
#pragma unroll
    for (int it = 0; it < kPeakIters; ++it) {
#pragma unroll
        for (int i = 0; i < NC; ++i) {
            const float2 lo = mix2(va[it].x, va[it].y, vb[it].x, vb[it].y, ca[i], cb[i]);
            const float2 hi = mix2(va[it].z, va[it].w, vb[it].z, vb[it].w, ca[i], cb[i]);
            m[i] = fmaxf(fmaxf(m[i], fabsf(lo.x)), fabsf(lo.y));
            m[i] = fmaxf(fmaxf(m[i], fabsf(hi.x)), fabsf(hi.y));
        }
    }
#pragma unroll
    for (int i = 0; i < NC; ++i) {   // one REDUX per warp and note, then straight to the note's global maximum
        const unsigned w = __reduce_max_sync(0xffffffffu, __float_as_uint(m[i]));
        if ((tid & 31) == 0 && w != 0u) atomicMax(peak_bits + i, (int)w);
    }
}

// One CTA per (group, chunk) work item: a group is the notes of one instrument in one segment
// (same two one-shots, one mixup each); a chunk is kPeakSpan samples of the mixed one-shot, read
// once into registers (all loads in flight together) and reused for every note of the group.
// Peaks are combined with atomicMax on the float bits (non-negative floats order like ints,
// and max is order-independent, so the result is deterministic).
__global__ void __launch_bounds__(kPeakThreads) peak_kernel(
    const float* __restrict__ pcm, const adtfe_event* __restrict__ events, const adtfe_peak_item* __restrict__ work,
    ResolvedEvent* __restrict__ resolved, int* __restrict__ peak_bits) {
    const adtfe_peak_item item = work[blockIdx.x / kPeakSplit];  // one fetch, then the data loads can start
    const int sub = blockIdx.x % kPeakSplit;                     // which part of the item's span this CTA scans
    const int chunk = item.chunk, tid = threadIdx.x;
    const int e0 = item.first_event, e1 = e0 + item.n_events;
    if (e0 >= e1) return;
    const int64_t a_off = item.a_off, b_off = item.b_off;
    const int la = item.la, lb = item.lb, n = item.mix_len;
    // every one-shot is padded to 4 floats, so whole float4s up to the padded length are readable
    const int la4 = (la + 3) >> 2, lb4 = (lb + 3) >> 2;
    const int lo4 = chunk * (kPeakSpan / 4) + sub * (kPeakCtaSpan / 4), hi4 = min((n + 3) >> 2, lo4 + kPeakCtaSpan / 4);
    if (lo4 >= hi4 && !(chunk == 0 && sub == 0)) return;  // this part lies past the end of the mixed one-shot
    const float4* a4 = reinterpret_cast<const float4*>(pcm + a_off);
    const float4* b4 = reinterpret_cast<const float4*>(pcm + b_off);

    float4 va[kPeakIters], vb[kPeakIters];
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    // interior chunks (both one-shots reach past the chunk) need neither load predicates nor tail masks
    const bool interior = lo4 + kPeakCtaSpan / 4 <= (la >> 2) && lo4 + kPeakCtaSpan / 4 <= (lb >> 2);
    if (interior) {
#pragma unroll
        for (int it = 0; it < kPeakIters; ++it) {
            const int i4 = lo4 + tid + it * kPeakThreads;
            va[it] = __ldg(a4 + i4);
            vb[it] = __ldg(b4 + i4);
        }
    } else {
#pragma unroll
        for (int it = 0; it < kPeakIters; ++it) {
            const int i4 = lo4 + tid + it * kPeakThreads;
            va[it] = (i4 < hi4 && i4 < la4) ? __ldg(a4 + i4) : z;
            vb[it] = (i4 < hi4 && i4 < lb4) ? __ldg(b4 + i4) : z;
        }
    }
    if (chunk == 0 && sub == 0) {  // resolve the bank lookups once per note for the tile mixer
        for (int e = e0 + tid; e < e1; e += kPeakThreads) {
            const adtfe_event ev = events[e];
            ResolvedEvent r;
            r.a_off = a_off; r.b_off = b_off;
            r.la = min(la, ev.len); r.lb = min(lb, ev.len);
            r.start = ev.start; r.len = ev.len; r.ca = ev.ca; r.cb = ev.cb; r.gain = ev.gain; r.pad = 0;
            resolved[e] = r;
        }
    }
    if (!interior) {
        // the float4 that holds a one-shot's last sample: ignore whatever pads it (all later ones were not loaded)
        const int qa = la >> 2, ra = la & 3, qb = lb >> 2, rb = lb & 3;
#pragma unroll
        for (int it = 0; it < kPeakIters; ++it) {
            const int i4 = lo4 + tid + it * kPeakThreads;
            if (i4 == qa && ra != 0) {
                if (ra < 2) va[it].y = 0.f;
                if (ra < 3) va[it].z = 0.f;
                va[it].w = 0.f;
            }
            if (i4 == qb && rb != 0) {
                if (rb < 2) vb[it].y = 0.f;
                if (rb < 3) vb[it].z = 0.f;
                vb[it].w = 0.f;
            }
        }
    }
    for (int c0 = e0; c0 < e1; c0 += kPeakChunk) {
        const int nc = min(kPeakChunk, e1 - c0);
        switch (nc) {
            case 1: peak_chunk<1>(va, vb, events + c0, peak_bits + c0, tid); break;
            case 2: peak_chunk<2>(va, vb, events + c0, peak_bits + c0, tid); break;
            case 3: peak_chunk<3>(va, vb, events + c0, peak_bits + c0, tid); break;
            case 4: peak_chunk<4>(va, vb, events + c0, peak_bits + c0, tid); break;
            case 5: peak_chunk<5>(va, vb, events + c0, peak_bits + c0, tid); break;
            case 6: peak_chunk<6>(va, vb, events + c0, peak_bits + c0, tid); break;
            case 7: peak_chunk<7>(va, vb, events + c0, peak_bits + c0, tid); break;
            default: peak_chunk<8>(va, vb, events + c0, peak_bits + c0, tid); break;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Tile mixer.  Persistent CTAs: one producer warp stages the one-shot slices a tile needs through a
// ring of shared-memory buffers with TMA bulk copies (cp.async.bulk + mbarrier), the consumer
// warps accumulate them in arrival (= event) order into registers.  Tiles are handed out by an
// atomic counter, so long and short tiles balance.  Measured on B200 (tools/gpu_variants.sh): the
// kernel is bound by the latency chain of a tile (queue head -> tile_ptr -> events -> records -> TMA),
// so many small CTAs (4 consumer warps x 16 samples per thread, a 2-slot ring, 8 CTAs per SM) beat
// few large ones (8 x 8, 6 slots, 4 per SM) by 12 % on the whole render.  The kernel is issue-bound (ncu: 75 % issue
// utilisation, a fifth of the instructions useful LDS + FFMA), so the per-slice hand-off is paid by as few warps as
// possible: 2 consumer warps x 32 samples per thread, and as many CTAs as shared memory holds (10-11 per SM, launch
// bound 12 = 55 registers) measured 15.04 ms per step against 15.26 ms for 4 x 16 / 8 CTAs.
#ifndef ADTFE_MIX_CONSUMERS
#define ADTFE_MIX_CONSUMERS 2
#endif
#ifndef ADTFE_MIX_STAGES
#define ADTFE_MIX_STAGES 2
#endif
#ifndef ADTFE_MIX_CTAS
#define ADTFE_MIX_CTAS 12
#endif
constexpr int kMixConsumers = ADTFE_MIX_CONSUMERS;       // warps; each thread owns tile / (32 * warps) samples
constexpr int kMixThreads = (kMixConsumers + 1) * 32;    // + the producer warp
constexpr int kPerThread = ADTFE_TILE / (kMixConsumers * 32);
constexpr int kStages = ADTFE_MIX_STAGES;
constexpr int kMixCtasPerSm = ADTFE_MIX_CTAS;
constexpr int kStageFloats = 2080;                       // >= 2048 + 2*3 alignment slack, bytes a multiple of 128
static_assert(kPerThread * kMixConsumers * 32 == ADTFE_TILE && kPerThread <= 32 && kPerThread % 2 == 0,
              "tile / consumer threads");

struct __align__(16) StageDesc {   // written by the producer lane, read (broadcast) by every consumer
    int32_t kind;                  // 0: data, 1: a new tile begins (tile id in `tile`, < 0: no more work)
    int32_t tile;
    int32_t base;                  // sample i of the tile sits at buffer index i + base
    int32_t vlo, vhi;              // samples vlo <= i < vhi of the tile are covered
    float coef;
    int32_t pad0, pad1;
};

struct __align__(16) SliceMsg {    // producer-private list entry: the StageDesc words + what the TMA copy needs
    int4 d0;                       // kind, tile, base, vlo
    int4 d1;                       // vhi, coef bits, -, -
    int4 d2;                       // source float offset (lo, hi), bytes (0: no copy, plain arrive), -
};
constexpr int kListMax = 65;       // a tile marker + two sources of 32 notes
constexpr int kIssue = kStages / 2 > 0 ? kStages / 2 : 1;  // messages issued per producer pass

__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kMixConsumers * 32) : "memory"); }

__host__ __device__ constexpr size_t mix_list_offset() {
    return (((size_t)kStages * kStageFloats * 4 + kStages * 32 + 2 * kStages * 8 + (kMixConsumers + 2) * 4) + 15) & ~(size_t)15;
}

struct MixArgs {
    const float* pcm;
    const ResolvedEvent* resolved;
    const int* peak_bits;
    const int32_t* tile_ptr;
    const int32_t* tile_events;
    const adtfe_segment* segments;
    float* wav;
    float* tile_max;
    int* tile_counter;
    int64_t ld_wav;
    int32_t tiles_per_seg, n_tiles;
};

// Writes the finished (not yet normalised) tile and publishes its |max|.
__device__ __forceinline__ void finish_tile(const MixArgs& a, int tile_id, const float2 (&acc2)[kPerThread / 2], int tid,
                                            float* s_red) {
    const int seg = tile_id / a.tiles_per_seg, lo = (tile_id - seg * a.tiles_per_seg) * ADTFE_TILE;
    // |max| on the float bits: non-negative floats order like unsigned integers and a NaN's bits lie above infinity's,
    // so the unsigned maximum propagates NaN like torch.max - one LOP3 + one integer max per sample, one REDUX per warp
    unsigned mb = 0u;
    float* row = a.wav + (int64_t)seg * a.ld_wav;
#pragma unroll
    for (int j = 0; j < kPerThread; ++j) {
        const int n = lo + tid + j * (kMixConsumers * 32);
        const float v = (j & 1) ? acc2[j >> 1].y : acc2[j >> 1].x;
        if (n < a.ld_wav) row[n] = v;
        mb = max(mb, __float_as_uint(fabsf(v)));
    }
    mb = __reduce_max_sync(0xffffffffu, mb);
    consumer_sync();  // s_red of the previous tile is no longer read
    if ((tid & 31) == 0) s_red[tid >> 5] = __uint_as_float(mb);
    consumer_sync();
    if (tid == 0) {
        unsigned rb = __float_as_uint(s_red[0]);
        for (int i = 1; i < kMixConsumers; ++i) rb = max(rb, __float_as_uint(s_red[i]));
        a.tile_max[tile_id] = __uint_as_float(rb);
    }
}

__global__ void __launch_bounds__(kMixThreads, kMixCtasPerSm) mix_kernel(const MixArgs a) {
    extern __shared__ __align__(128) unsigned char mix_smem[];
    float* s_buf = reinterpret_cast<float*>(mix_smem);                               // kStages * kStageFloats
    StageDesc* s_desc = reinterpret_cast<StageDesc*>(s_buf + kStages * kStageFloats);
    uint64_t* s_full = reinterpret_cast<uint64_t*>(s_desc + kStages);
    uint64_t* s_empty = s_full + kStages;
    float* s_red = reinterpret_cast<float*>(s_empty + kStages);
    SliceMsg* s_list = reinterpret_cast<SliceMsg*>(mix_smem + mix_list_offset());

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(s_full + s, 1); mbar_init(s_empty + s, kMixConsumers); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    int stage = 0;
    uint32_t phase = 0;

    if (warp == kMixConsumers) {
        // ================= producer warp =================
        // Per tile: a marker message, then one message per live (note, source) slice in event order.  The
        // messages of up to 32 notes are first laid out in a warp-private list (ballot + popcount give every
        // lane its slots); then lane j takes the j-th next message: it waits for ring slot stage + j to drain,
        // publishes the descriptor and starts the TMA copy - kStages slots are refilled per pass instead of one.
        SliceMsg* list = s_list;
        for (;;) {
            int tile = 0;
            if (lane == 0) tile = atomicAdd(a.tile_counter, 1);
            tile = __shfl_sync(0xffffffffu, tile, 0);
            const bool done = tile >= a.n_tiles;
            int p0 = 0, p1 = 0, lo = 0;
            if (!done) {
                const int seg = tile / a.tiles_per_seg;
                lo = (tile - seg * a.tiles_per_seg) * ADTFE_TILE;
                p0 = __ldg(a.tile_ptr + tile);
                p1 = __ldg(a.tile_ptr + tile + 1);
            }
            bool first = true;
            int base = p0;
            do {
                const int n = min(32, p1 - base);
                // lane i owns event base + i: both of its sources, clipped to this tile
                int64_t off[2] = {0, 0};
                int r_lo = 0, r_hi[2] = {0, 0}, rel = 0;
                float coef[2] = {0.0f, 0.0f};
                if (lane < n) {
                    const int e = __ldg(a.tile_events + base + lane);
                    const ResolvedEvent ev = a.resolved[e];
                    // an all-zero one-shot has peak 0: gain/0 = inf (or NaN), and 0 * inf = NaN over the
                    // whole note - the reference's o / 0 (synthetiser.py:225)
                    const float scale = ev.gain / __int_as_float(a.peak_bits[e]);
                    rel = lo - ev.start;  // tile sample i is source sample i + rel
                    off[0] = ev.a_off; off[1] = ev.b_off;
                    coef[0] = ev.ca * scale; coef[1] = ev.cb * scale;
                    r_lo = max(0, rel);
                    r_hi[0] = min(ev.la, rel + ADTFE_TILE);
                    r_hi[1] = min(ev.lb, rel + ADTFE_TILE);
                }
                const bool live0 = r_hi[0] > r_lo, live1 = r_hi[1] > r_lo;
                const unsigned m0 = __ballot_sync(0xffffffffu, live0), m1 = __ballot_sync(0xffffffffu, live1);
                const unsigned below = (1u << lane) - 1u;
                int slot = (first ? 1 : 0) + __popc(m0 & below) + __popc(m1 & below);
                const int total = (first ? 1 : 0) + __popc(m0) + __popc(m1);
                if (first && lane == 0) {
                    SliceMsg m;
                    m.d0 = make_int4(1, done ? -1 : tile, 0, 0);
                    m.d1 = make_int4(0, 0, 0, 0);
                    m.d2 = make_int4(0, 0, 0, 0);
                    list[0] = m;
                }
#pragma unroll
                for (int sidx = 0; sidx < 2; ++sidx) {
                    if (sidx == 0 ? live0 : live1) {
                        const int g_lo = r_lo & ~3, g_hi = (r_hi[sidx] + 3) & ~3;  // storage is padded to 4 floats
                        const long long src = off[sidx] + g_lo;
                        SliceMsg m;
                        m.d0 = make_int4(0, tile, rel - g_lo, r_lo - rel);
                        m.d1 = make_int4(r_hi[sidx] - rel, __float_as_int(coef[sidx]), 0, 0);
                        m.d2 = make_int4((int)(unsigned)(src & 0xffffffffll), (int)(src >> 32), (g_hi - g_lo) * 4, 0);
                        list[slot++] = m;
                    }
                }
                __syncwarp();
                for (int k0 = 0; k0 < total; k0 += kIssue) {
                    const int k = k0 + lane;
                    if (lane < kIssue && k < total) {
                        int st = stage + lane;
                        uint32_t ph = phase;
                        if (st >= kStages) { st -= kStages; ph ^= 1u; }
                        const SliceMsg m = list[k];
                        mbar_wait_relaxed(s_empty + st, ph ^ 1u, 2000u);
                        int4* d = reinterpret_cast<int4*>(s_desc + st);
                        d[0] = m.d0;
                        d[1] = m.d1;
                        if (m.d2.z > 0) {
                            const long long src = (long long)(((unsigned long long)(unsigned)m.d2.y << 32) | (unsigned)m.d2.x);
                            mbar_expect_tx(s_full + st, (uint32_t)m.d2.z);
                            bulk_g2s(s_buf + st * kStageFloats, a.pcm + src, (uint32_t)m.d2.z, s_full + st);
                        } else {
                            mbar_arrive(s_full + st);
                        }
                    }
                    stage += min(kIssue, total - k0);
                    if (stage >= kStages) { stage -= kStages; phase ^= 1u; }
                }
                __syncwarp();  // the list is rewritten by the next batch
                first = false;
                base += 32;
            } while (base < p1);
            if (done) break;
        }
        return;
    }

    // ================= consumer warps =================
    // sample j of the thread (tile index tid + T*j) is component j & 1 of acc2[j / 2]: a full slice costs one
    // packed FFMA2 per two samples (the kernel is issue-bound)
    float2 acc2[kPerThread / 2];
#pragma unroll
    for (int k = 0; k < kPerThread / 2; ++k) acc2[k] = make_float2(0.0f, 0.0f);
    int tile = -1;
    for (;;) {
        mbar_wait_relaxed(s_full + stage, phase, 2000u);
        const int4 d0 = *reinterpret_cast<const int4*>(s_desc + stage);
        if (d0.x != 0) {  // a new tile begins
            __syncwarp();
            if (lane == 0) mbar_arrive(s_empty + stage);
            if (++stage == kStages) { stage = 0; phase ^= 1u; }
            if (tile >= 0) finish_tile(a, tile, acc2, tid, s_red);
            tile = d0.y;
            if (tile < 0) break;
#pragma unroll
            for (int k = 0; k < kPerThread / 2; ++k) acc2[k] = make_float2(0.0f, 0.0f);
            continue;
        }
        const int4 d1 = *reinterpret_cast<const int4*>(reinterpret_cast<const char*>(s_desc + stage) + 16);
        const int vlo = d0.w, vhi = d1.x;
        const float coef = __int_as_float(d1.y);
        const float* src = s_buf + stage * kStageFloats + d0.z + tid;
        if (vlo == 0 && vhi == ADTFE_TILE) {
            constexpr int T = kMixConsumers * 32;
            const float2 c2 = make_float2(coef, coef);
#pragma unroll
            for (int k = 0; k < kPerThread / 2; ++k)
                acc2[k] = __ffma2_rn(make_float2(src[(2 * k) * T], src[(2 * k + 1) * T]), c2, acc2[k]);
        } else {
            // A note starts or ends inside the tile: this thread's samples tid + T*j lie inside [vlo, vhi) for
            // jlo <= j < jhi.  One bit mask per slice instead of two compares per sample (the kernel is issue-bound);
            // samples outside the note stay untouched even when coef is inf / NaN.
            constexpr int T = kMixConsumers * 32;
            const int jlo = (max(vlo - tid, 0) + T - 1) / T, jhi = min((max(vhi - tid, 0) + T - 1) / T, kPerThread);
            const unsigned below_hi = jhi >= 32 ? 0xffffffffu : (1u << jhi) - 1u;
            const unsigned below_lo = jlo >= 32 ? 0xffffffffu : (1u << jlo) - 1u;
            const unsigned mask = below_hi & ~below_lo;   // empty when jhi <= jlo
#pragma unroll
            for (int j = 0; j < kPerThread; ++j) {
                if ((mask >> j) & 1u) {
                    if (j & 1) acc2[j >> 1].y = fmaf(src[j * T], coef, acc2[j >> 1].y);
                    else acc2[j >> 1].x = fmaf(src[j * T], coef, acc2[j >> 1].x);
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(s_empty + stage);
        if (++stage == kStages) { stage = 0; phase ^= 1u; }
    }
}


// Row normalisation (wav / peak * max_volume, synthetiser.py:142-144,156): one CTA per tile; the
// segment peak is the max of its tile maxima.  An all-zero mix gives NaN (the reference's 0/0),
// samples beyond the segment's length stay exact zeros (collate_fn pads with 0.0).
#ifndef ADTFE_NORM_THREADS
#define ADTFE_NORM_THREADS 256
#endif
constexpr int kNormThreads = ADTFE_NORM_THREADS;
__global__ void __launch_bounds__(kNormThreads) normalise_kernel(const adtfe_segment* __restrict__ segments,
                                                                 const float* __restrict__ tile_max, int tiles_per_seg,
                                                                 int64_t ld_wav, float* __restrict__ wav) {
    const int tile_id = blockIdx.x, tid = threadIdx.x;
    const int seg = tile_id / tiles_per_seg, lo = (tile_id - seg * tiles_per_seg) * ADTFE_TILE;
    const adtfe_segment sg = segments[seg];
    if (sg.flags == 0 || lo >= sg.len) return;
    float peak = 0.0f;
    for (int t = tid & 31; t < tiles_per_seg; t += 32) peak = nan_max(peak, __ldg(tile_max + seg * tiles_per_seg + t));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) peak = nan_max(peak, __shfl_xor_sync(0xffffffffu, peak, o));
    const float vol = sg.max_volume;
    float* row = wav + (int64_t)seg * ld_wav + lo;  // ld_wav is a multiple of 4 and the base is 16-byte aligned
    const int n = min(ADTFE_TILE, sg.len - lo);
    float4* row4 = reinterpret_cast<float4*>(row);
    const int n4 = n >> 2;
    // v / peak * vol with the division by an invariant done the way the hardware sequence does it: q = v*r,
    // one residual correction (correctly rounded for normal operands, i.e. everything a mix can hold), r = 1/peak
    // rounded to nearest once per thread.  peak = 0 (all-zero mix) or NaN still gives NaN everywhere.
    const float r = __frcp_rn(peak);
    auto norm = [&](float v) {
        const float q = __fmul_rn(v, r);
        const float rem = __fmaf_rn(-q, peak, v);
        return __fmul_rn(__fmaf_rn(rem, r, q), vol);
    };
    for (int i = tid; i < n4; i += kNormThreads) {
        float4 v = row4[i];
        v.x = norm(v.x); v.y = norm(v.y); v.z = norm(v.z); v.w = norm(v.w);
        row4[i] = v;
    }
    for (int i = 4 * n4 + tid; i < n; i += kNormThreads) row[i] = norm(row[i]);
}

static size_t mix_smem_bytes() { return mix_list_offset() + kListMax * sizeof(SliceMsg); }

}  // namespace adtfe

using namespace adtfe;

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

// Called by adtfe_bank_create on the bank's device (cudaFuncSetAttribute is per device and idempotent): the kernels'
// dynamic shared memory is opted in once per handle, not through process-wide state on the render path.
int adtfe::mixer_prepare_device() {
    ADTFE_CUDA(cudaFuncSetAttribute(mix_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mix_smem_bytes()));
    return ADTFE_OK;
}

extern "C" size_t adtfe_render_workspace_bytes(int32_t n_events, int32_t n_seg, int32_t tiles_per_seg) {
    if (n_events < 0 || n_seg < 0 || tiles_per_seg < 0) return 0;
    // resolved events | peak bits, queue heads (one zeroed block) | tile maxima
    return align256((size_t)n_events * sizeof(ResolvedEvent)) + align256((size_t)n_events * 4 + (size_t)n_seg * 4 + 4) +
           align256((size_t)n_seg * tiles_per_seg * 4) + 256;
}

extern "C" int adtfe_render(const adtfe_bank* bank, const adtfe_plan* plan, float* wav_out_dev, void* workspace_dev,
                            size_t workspace_bytes, void* stream) {
    return render_impl(bank, plan, wav_out_dev, workspace_dev, workspace_bytes, stream);
}

int adtfe::render_impl(const adtfe_bank* bank, const adtfe_plan* plan, float* wav_out_dev, void* workspace_dev,
                       size_t workspace_bytes, void* stream) {
    ADTFE_REQUIRE(bank && plan, ADTFE_ERR_BAD_ARG, "adtfe_render: null bank or plan");
    ADTFE_REQUIRE(plan->n_seg >= 0 && plan->n_events >= 0 && plan->n_peak_work >= 0 && plan->tiles_per_seg >= 0,
                  ADTFE_ERR_BAD_ARG, "adtfe_render: negative count");
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
    const size_t need = adtfe_render_workspace_bytes(plan->n_events, plan->n_seg, plan->tiles_per_seg);
    ADTFE_REQUIRE(workspace_dev && workspace_bytes >= need, ADTFE_ERR_WORKSPACE,
                  "adtfe_render: workspace %zu B < %zu B", workspace_bytes, need);
    // chunk boundaries: the plan's, or one chunk covering everything
    const adtfe_chunk whole[2] = {{0, 0, 0}, {plan->n_seg, plan->n_events, plan->n_peak_work}};
    const adtfe_chunk* ch = whole;
    int n_chunks = 1;
    if (plan->chunks_host && plan->n_chunks > 0) {
        ch = plan->chunks_host;
        n_chunks = plan->n_chunks;
        ADTFE_REQUIRE(n_chunks <= plan->n_seg && ch[0].seg == 0 && ch[0].event == 0 && ch[0].peak_work == 0 &&
                          ch[n_chunks].seg == plan->n_seg && ch[n_chunks].event == plan->n_events &&
                          ch[n_chunks].peak_work == plan->n_peak_work,
                      ADTFE_ERR_BAD_ARG, "adtfe_render: chunk boundaries do not cover the plan");
        for (int c = 0; c < n_chunks; ++c)
            ADTFE_REQUIRE(ch[c + 1].seg > ch[c].seg && ch[c + 1].event >= ch[c].event &&
                              ch[c + 1].peak_work >= ch[c].peak_work,
                          ADTFE_ERR_BAD_ARG, "adtfe_render: chunk %d is empty or out of order", c);
    }
    cudaStream_t user = (cudaStream_t)stream;
    char* ws = (char*)(((uintptr_t)workspace_dev + 255) & ~(uintptr_t)255);
    ResolvedEvent* resolved = (ResolvedEvent*)ws;
    int* peak_bits = (int*)(ws + align256((size_t)plan->n_events * sizeof(ResolvedEvent)));
    int* counters = peak_bits + plan->n_events;  // one work-queue head per chunk (n_chunks <= n_seg)
    float* tile_max = (float*)((char*)peak_bits + align256((size_t)plan->n_events * 4 + (size_t)plan->n_seg * 4 + 4));
    // zero the peaks and the queue heads once, then fork the chunks over the bank's streams
    ADTFE_CUDA(cudaMemsetAsync(peak_bits, 0, ((size_t)plan->n_events + (size_t)plan->n_seg + 1) * 4, user));
    const bool fork = n_chunks > 1 && bank->n_streams > 0;
    std::unique_lock<std::mutex> lock(bank->mu, std::defer_lock);
    if (fork) {
        lock.lock();
        ADTFE_CUDA(cudaEventRecord(bank->fork_event, user));
        for (int k = 0; k < bank->n_streams && k < n_chunks; ++k)
            ADTFE_CUDA(cudaStreamWaitEvent(bank->streams[k], bank->fork_event, 0));
    }
    float* mix_out = wav_out_dev;
    const int tps = plan->tiles_per_seg;
    for (int c = 0; c < n_chunks; ++c) {
        cudaStream_t st = fork ? bank->streams[c % bank->n_streams] : user;
        const int s0 = ch[c].seg, n_seg = ch[c + 1].seg - s0;
        const int pw0 = ch[c].peak_work, n_pw = ch[c + 1].peak_work - pw0;
        if (n_pw > 0) {
            trace_open("peak", c, st);
            peak_kernel<<<n_pw * kPeakSplit, kPeakThreads, 0, st>>>(bank->pcm, plan->events_dev, plan->peak_work_dev + pw0, resolved,
                                                      peak_bits);
            trace_close(st);
            ADTFE_CUDA(cudaGetLastError());
        }
        MixArgs a;
        a.pcm = bank->pcm; a.resolved = resolved; a.peak_bits = peak_bits;
        a.tile_ptr = plan->tile_ptr_dev + (size_t)s0 * tps; a.tile_events = plan->tile_events_dev;
        a.segments = plan->segments_dev + s0; a.wav = mix_out + (size_t)s0 * plan->ld_wav;
        a.tile_max = tile_max + (size_t)s0 * tps; a.tile_counter = counters + c; a.ld_wav = plan->ld_wav;
        a.tiles_per_seg = tps; a.n_tiles = n_seg * tps;
        const int grid = std::min(a.n_tiles, kMixCtasPerSm * bank->sm_count);
        trace_open("mix", c, st);
        mix_kernel<<<grid, kMixThreads, mix_smem_bytes(), st>>>(a);
        trace_close(st);
        ADTFE_CUDA(cudaGetLastError());
        trace_open("normalise", c, st);
        normalise_kernel<<<a.n_tiles, kNormThreads, 0, st>>>(a.segments, a.tile_max, tps, plan->ld_wav, a.wav);
        trace_close(st);
        ADTFE_CUDA(cudaGetLastError());
    }
    if (fork) {
        for (int k = 0; k < bank->n_streams && k < n_chunks; ++k) {
            ADTFE_CUDA(cudaEventRecord(bank->join_events[k], bank->streams[k]));
            ADTFE_CUDA(cudaStreamWaitEvent(user, bank->join_events[k], 0));
        }
    }
    return ADTFE_OK;
}
