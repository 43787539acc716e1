// Fused STFT -> |X|^2 -> mel -> log kernel (sm_100a).
//
// Replaces ComputeMelSpectrogram.forward (reference model.py:81-97) and the torchaudio
// MelSpectrogram it calls (model.py:71-78,89): frame (n_fft 2048, centred), periodic Hann,
// real FFT, power, (n_fft/2+1 x n_mels) triangular filterbank, log(x + 1e-10),
// clamp[-23, 12], (x + 23) / 35, keep frames wpi .. T-wpi-2.  Only kept frames are
// computed; their support never reaches the reflect padding.
//
// Persistent kernel, one CTA of 8 warps per SM, up to 255 registers per thread.  A *round* is up
// to 32 consecutive kept frames of one segment (a segment's frames are split into equal rounds);
// nothing but the (frames x n_mels) result goes to HBM.
//
// The fp32 pipe of sm_100a retires one scalar FFMA/FADD/FMUL per clock per scheduler, and a packed
// FFMA2/FADD2/FMUL2 (two results per lane) every two clocks - same arithmetic rate, half the issue
// slots (tools/microbench/fp32_pipe.cu).  The scalar kernel was issue-bound, so every warp now
// transforms TWO frames at once: all values are float2 = (frame f, frame f+1), every butterfly,
// twiddle and power is one packed instruction, and window / twiddle loads and all address
// arithmetic are shared by the two frames.
//
//  span    the round's samples - (frames-1)*hop + 2048 floats, every sample once although it
//          feeds 8.5 frames - are brought into shared memory by one TMA bulk copy
//          (cp.async.bulk + mbarrier) issued a round ahead, so global-load latency never sits
//          on the critical path and no registers are spent on prefetching.  Rows that are not
//          16-byte aligned take a cooperative copy instead.
//  FFT     one warp owns one frame pair at a time (two pairs per round):
//   pass 1  lane n2 holds x[32*n1 + n2] * hann, n1 = 0..63, and runs a 64-point real DFT
//           over n1 in registers (generated straight-line code, tools/gen_fft.py)
//   twiddle Y[k1][n2] *= W_2048^(k1*n2)
//   exchange through the pair's own two columns of the power matrix P[bin][frame] (not yet
//           written): element (k1, n2) sits in row 33*k1 + n2, so the writer (lane n2, one k1 per
//           STS.64) and the reader (lane k1, one n2 per LDS.64) are both conflict-free with
//           immediate offsets only; real parts, then imaginary parts
//   pass 2  lane k1 (0..31) runs a 32-point complex DFT over n2 -> X[k1 + 64*k2];
//           bins above 1024 are the mirror images of bins 64-k1 + 64*(31-k2)
//   column k1 = 32 (bins 32 + 64*k2) is a 32-point DFT across lanes with shuffles
//   |X|^2 goes to P[bin][f .. f+1] (row pitch 34: STS.64 of consecutive bins and LDS.32 of
//           consecutive frames are conflict-free)
//  mel     lanes are the frames.  A triangular filterbank has at most two adjacent filters per
//          bin, so bins are walked once: the bins between two filter centres feed the falling
//          edge of the lower filter and the rising edge of the upper one - one packed FFMA2 per
//          loaded power value on the (up, down) weight pair, weights by broadcast LDS.128.  Each
//          warp owns a contiguous, cost-balanced range of such intervals.  Filterbanks without
//          that structure take a plain per-filter loop over the same P layout.
//  output  log / clamp / affine on the staged (filter x frame) tile, 128-byte coalesced rows.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.cuh"
#include "fft_gen.cuh"

namespace adtfe {

constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;
constexpr int kRound = 32;                   // frames per round (lanes of the mel phase)
constexpr int kPairsPerWarp = kRound / 2 / kWarps;
constexpr int kSpanFloats = 9504;            // >= 31*240 + 2048, bytes a multiple of 128
constexpr int kPPitch = 34;                  // floats per bin row of P[bin][frame]: 32 frames + 2, pitch/2 odd
constexpr int kXPitch = 33;                  // exchange element (k1, n2) lives in row 33*k1 + n2 of the pair's columns
constexpr int kPRows = 1056;                 // 1025 bins + the zero pad row 1025; exchange rows up to 33*31 + 31
constexpr int kWinPitch = 68;                // window, transposed: lane n2 reads n1 = 4q..4q+3 with one LDS.128
constexpr int kMaxMels = 128;
constexpr int kSPitch = 33;
constexpr int kDummyRow = kMaxMels;          // S row kMaxMels: sink for sums nobody reads
constexpr int kSRows = kDummyRow + 1;
constexpr int kSFloats = 4288;               // >= kSRows * kSPitch, bytes a multiple of 128
constexpr int kGroupBins = 4;                // bins per mel group (fast path): two float4 of (up, down) weights
constexpr int kGroupsMax = 360;              // groups kept in shared memory (fast path)
constexpr int kW4Max = 2 * kGroupsMax;
static_assert(kSRows * kSPitch <= kSFloats, "S tile");
static_assert((kSpanFloats * 4) % 128 == 0 && (kPPitch * 4 * kPRows) % 128 == 0 && (kSFloats * 4) % 128 == 0,
              "alignment of the shared carve-up");
static_assert(kXPitch * 31 + 31 < kPRows, "exchange rows");
static_assert(kPairsPerWarp * 2 * kWarps == kRound, "pairs per warp");

// Mel phase tables.  Fast path (triangular filterbank): the bins between the centres of filters j-1
// and j form interval j; it is cut into groups of kGroupBins bins {first bin, S row to flush into or -1}
// with two float4 of weights {up0, down0, up1, down1} each (zero for bins outside the interval), kept in
// global memory and staged into shared memory by every CTA.  Generic path: filter m (first bin, bins,
// float offset of its weights in global memory).
struct MelItem {
    int16_t b0, n;
    int32_t woff;
};
struct MelGroup {
    int32_t b0;   // first of the group's kGroupBins power rows (b0 + kGroupBins - 1 <= n_bins: the pad row)
    int32_t row;  // >= 0: last group of its interval - S[row] = (rising-edge sum carried) + (falling-edge sum)
};
struct MelTables {
    MelItem item[kMaxMels + 1];  // generic path only
    int16_t first[kWarps + 1];   // warp w owns items (generic) / groups (fast) first[w] .. first[w+1]-1
    int32_t fast, n_groups;
};

struct LogmelArgs {
    const float* wav;
    float* out;
    const float* window;
    const float2* twiddle;
    const float2* lane_tw;
    const float* weights;
    const MelGroup* groups;
    const adtfe_mel_row* rows;  // ragged form: per-segment frame count and first output row (else NULL)
    const int* seg_mark;        // row filter of the warp-autonomous kernel (else NULL): rows with (seg_mark[seg] != 0) !=
    int32_t mark_want;          // (mark_want != 0) are left out of this launch (the render's FX rows come later)
    int64_t ld_wav;
    int32_t n_seg, first, count, hop, n_mels;
    int32_t rounds_per_seg, n_rounds;
};

// Round r = part q of segment seg: a segment's `count` frames are split evenly over rounds_per_seg rounds,
// round q has count / rounds_per_seg + (q < count % rounds_per_seg) frames starting at kept frame j0.
struct RoundGeom {
    int seg, j0, nf;
    long long out_row;  // output row of the round's first frame
};
__device__ __forceinline__ RoundGeom round_geom(const LogmelArgs& p, int r) {
    RoundGeom g;
    g.seg = (int)((unsigned)r / (unsigned)p.rounds_per_seg);
    const int q = r - g.seg * p.rounds_per_seg;
    int count = p.count;
    long long base = (long long)g.seg * p.count;
    if (p.rows) {
        const int4 row = __ldg(reinterpret_cast<const int4*>(p.rows + g.seg));  // {out_row lo, hi, count, flags (a silent row is zeros anyway)}
        base = (long long)(((unsigned long long)(unsigned)row.y << 32) | (unsigned)row.x);
        count = row.z;
    }
    const int fb = count / p.rounds_per_seg, rem = count - fb * p.rounds_per_seg;
    g.j0 = q * fb + min(q, rem);
    g.nf = fb + (q < rem ? 1 : 0);
    g.out_row = base + g.j0;
    return g;
}

// Bring the samples of round r into s_span: one TMA bulk copy when source and length are 16-byte
// aligned, a cooperative copy otherwise.  Called by all threads (the choice is CTA-uniform).
__device__ __forceinline__ void issue_span(const LogmelArgs& p, const RoundGeom& g, float* s_span, uint64_t* bar,
                                           int tid) {
    const float* src = p.wav + (long long)g.seg * p.ld_wav + (long long)(p.first + g.j0) * p.hop - 1024;
    const int len = (g.nf - 1) * p.hop + 2048;
    if (g.nf <= 0) {  // a segment with fewer frames than rounds: nothing to fetch
        if (tid == 0) mbar_arrive(bar);
    } else if ((((uintptr_t)src | (uintptr_t)(len * 4)) & 15) == 0) {
        if (tid == 0) {
            mbar_expect_tx(bar, (uint32_t)len * 4u);
            bulk_g2s(s_span, src, (uint32_t)len * 4u, bar);
        }
    } else {
        for (int i = tid; i < len; i += kThreads) s_span[i] = __ldg(src + i);
        if (tid == 0) mbar_arrive(bar);  // the stores become visible at the CTA barrier that follows
    }
}

__device__ __forceinline__ float2 bc2(float a) { return make_float2(a, a); }
__device__ __forceinline__ float2 neg2(float2 a) { return make_float2(-a.x, -a.y); }
__device__ __forceinline__ float2 shfl_xor2(float2 v, int m) {
    return make_float2(__shfl_xor_sync(0xffffffffu, v.x, m), __shfl_xor_sync(0xffffffffu, v.y, m));
}
// (ar + i ai) * (w.x + i w.y), both halves of the pair by the same w: 4 packed instructions
__device__ __forceinline__ void cmul2(float2& ar, float2& ai, const float2 w) {
    const float2 t = __fmul2_rn(ai, bc2(w.y));
    const float2 u = __fmul2_rn(ai, bc2(w.x));
    ai = __ffma2_rn(ar, bc2(w.y), u);
    ar = __ffma2_rn(ar, bc2(w.x), neg2(t));
}

__global__ void __launch_bounds__(kThreads, 1) logmel_kernel(const LogmelArgs p, const __grid_constant__ MelTables tab) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* s_span = reinterpret_cast<float*>(smem_raw);               // kSpanFloats
    float* s_p = s_span + kSpanFloats;                                // kPRows * kPPitch
    float* s_s = s_p + kPRows * kPPitch;                              // kSFloats
    float4* s_w4 = reinterpret_cast<float4*>(s_s + kSFloats);         // kW4Max
    float* s_win = reinterpret_cast<float*>(s_w4 + kW4Max + 2);       // 32 * kWinPitch (two spare weight slots: prefetch)
    float2* s_tw = reinterpret_cast<float2*>(s_win + 32 * kWinPitch); // 32*32
    float2* s_ltw = s_tw + 32 * 32;                                   // 4*32
    int2* s_grp = reinterpret_cast<int2*>(s_ltw + 4 * 32);            // kGroupsMax + 1
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_grp + kGroupsMax + 2);

    const int tid = threadIdx.x, lane = tid & 31;
    // broadcast from lane 0 so the compiler knows the warp index is warp-uniform
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);

    if (tid == 0) mbar_init(s_bar, 1);
    for (int i = tid; i < 2048; i += kThreads) s_win[(i & 31) * kWinPitch + (i >> 5)] = p.window[i];
    for (int i = tid; i < 32 * 32; i += kThreads) s_tw[i] = p.twiddle[i];
    for (int i = tid; i < 4 * 32; i += kThreads) s_ltw[i] = p.lane_tw[i];
    for (int i = tid; i < 2 * tab.n_groups; i += kThreads) s_w4[i] = __ldg(reinterpret_cast<const float4*>(p.weights) + i);
    for (int i = tid; i <= kGroupsMax; i += kThreads)  // the entry after the last group is prefetched, never used
        s_grp[i] = i < tab.n_groups ? __ldg(reinterpret_cast<const int2*>(p.groups) + i) : make_int2(0, -1);
    for (int i = tid; i < kSFloats; i += kThreads) s_s[i] = 0.0f;
    for (int i = tid; i < kPRows * kPPitch; i += kThreads) s_p[i] = 0.0f;
    __syncthreads();
    RoundGeom geo_next = round_geom(p, min((int)blockIdx.x, p.n_rounds - 1));
    if ((int)blockIdx.x < p.n_rounds) issue_span(p, geo_next, s_span, s_bar, tid);
    __syncthreads();

    const int col32_bin = 32 + 64 * (int)(__brev((unsigned)lane) >> 27);
    uint32_t parity = 0;

    for (int round = blockIdx.x; round < p.n_rounds; round += gridDim.x) {
        const RoundGeom geo = geo_next;
        mbar_wait(s_bar, parity);
        parity ^= 1u;

        // ================= FFT phase: frame pairs warp and warp + 8 of the round =================
#pragma unroll 1
        for (int half = 0; half < kPairsPerWarp; ++half) {
            const int f = 2 * (warp + half * kWarps);
            if (f >= geo.nf) break;  // warp-uniform
            // an odd frame count leaves the last pair half empty: its second half repeats the first
            const int hop_b = (f + 1 < geo.nf) ? p.hop : 0;

            // ---- pass 1: window + 64-point real DFT over n1 (stride-32 samples), both frames
            float2 yr[33], yi[33];
            {
                const float* sa = s_span + f * p.hop + lane;
                const float* sb = sa + hop_b;
                const float4* wt = reinterpret_cast<const float4*>(s_win + lane * kWinPitch);
                float2 v[64];
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                    const float4 w = wt[q];
                    v[4 * q + 0] = __fmul2_rn(make_float2(sa[32 * (4 * q + 0)], sb[32 * (4 * q + 0)]), bc2(w.x));
                    v[4 * q + 1] = __fmul2_rn(make_float2(sa[32 * (4 * q + 1)], sb[32 * (4 * q + 1)]), bc2(w.y));
                    v[4 * q + 2] = __fmul2_rn(make_float2(sa[32 * (4 * q + 2)], sb[32 * (4 * q + 2)]), bc2(w.z));
                    v[4 * q + 3] = __fmul2_rn(make_float2(sa[32 * (4 * q + 3)], sb[32 * (4 * q + 3)]), bc2(w.w));
                }
                rdft64(v, yr, yi);
            }
            // ---- twiddle in place (rows 1..31), column 32 is real before its twiddle
#pragma unroll
            for (int k1 = 1; k1 < 32; ++k1) cmul2(yr[k1], yi[k1], s_tw[(k1 - 1) * 32 + lane]);
            float2 cr, ci;
            {
                const float2 w = s_tw[31 * 32 + lane];
                cr = __fmul2_rn(yr[32], bc2(w.x));
                ci = __fmul2_rn(yr[32], bc2(w.y));
            }
            // ---- exchange through columns f, f+1 of P: writer lane n2 -> row 33*k1 + n2, reader lane k1
            float* xw = s_p + lane * kPPitch + f;
            const float* xr = s_p + lane * (kXPitch * kPPitch) + f;
            float2 zr[32], zi[32];
            __syncwarp();
#pragma unroll
            for (int k1 = 0; k1 < 32; ++k1) *reinterpret_cast<float2*>(xw + k1 * (kXPitch * kPPitch)) = yr[k1];
            __syncwarp();
#pragma unroll
            for (int n2 = 0; n2 < 32; ++n2) zr[n2] = *reinterpret_cast<const float2*>(xr + n2 * kPPitch);
            __syncwarp();
            *reinterpret_cast<float2*>(xw) = make_float2(0.0f, 0.0f);  // row 0 is purely real
#pragma unroll
            for (int k1 = 1; k1 < 32; ++k1) *reinterpret_cast<float2*>(xw + k1 * (kXPitch * kPPitch)) = yi[k1];
            __syncwarp();
#pragma unroll
            for (int n2 = 0; n2 < 32; ++n2) zi[n2] = *reinterpret_cast<const float2*>(xr + n2 * kPPitch);
            __syncwarp();

            // ---- 32-point DFT across lanes for column 32 (decimation in frequency, bit-reversed):
            // d = other + sign * own (sign -1 on the upper half), then the per-lane twiddle of the stage
            // (1 on the lower half; W_{2*hb}^(lane mod hb), -i or 1 on the upper half)
#pragma unroll
            for (int s = 0; s < 5; ++s) {
                const int hb = 16 >> s;
                const float2 orr = shfl_xor2(cr, hb), oi = shfl_xor2(ci, hb);
                const float sg = (lane & hb) ? -1.0f : 1.0f;
                cr = __ffma2_rn(cr, bc2(sg), orr);
                ci = __ffma2_rn(ci, bc2(sg), oi);
                if (s < 4) cmul2(cr, ci, s_ltw[s * 32 + lane]);
            }

            // ---- pass 2: 32-point complex DFT over n2 for k1 = lane
            cdft32(zr, zi);

            // ---- power spectrum into P[bin][f .. f+1] (over the dead exchange rows)
            float* pc = s_p + f;
#pragma unroll
            for (int k2 = 0; k2 < 16; ++k2)
                *reinterpret_cast<float2*>(pc + (lane + 64 * k2) * kPPitch) =
                    __ffma2_rn(zi[k2], zi[k2], __fmul2_rn(zr[k2], zr[k2]));
#pragma unroll
            for (int k2 = 16; k2 < 32; ++k2)
                *reinterpret_cast<float2*>(pc + (64 - lane + 64 * (31 - k2)) * kPPitch) =
                    __ffma2_rn(zi[k2], zi[k2], __fmul2_rn(zr[k2], zr[k2]));
            if ((lane & 1) == 0)
                *reinterpret_cast<float2*>(pc + col32_bin * kPPitch) = __ffma2_rn(ci, ci, __fmul2_rn(cr, cr));
            // padded bin pairs read one bin past the spectrum with weight 0: keep it finite
            if (lane == 1) *reinterpret_cast<float2*>(pc + 1025 * kPPitch) = make_float2(0.0f, 0.0f);
        }
        __syncthreads();  // P complete; the span buffer is free

        // the next round's samples arrive while the mel phase runs
        if (round + (int)gridDim.x < p.n_rounds) {
            geo_next = round_geom(p, round + gridDim.x);
            issue_span(p, geo_next, s_span, s_bar, tid);
        }

        // ================= mel phase: lane = frame =================
        {
            const float* pl = s_p + lane;
            const int i0 = tab.first[warp], i1 = tab.first[warp + 1];
            if (tab.fast) {
                // groups i0 .. i1-1 of this warp, software-pipelined two at a time: while one group's four
                // packed FMAs run, the other's record, power values and weights are in flight.  a0 / a1 hold
                // the (rising-edge, falling-edge) sums of the interval's even / odd bins.  A warp's range
                // starts with the interval whose rising edge it needs (that flush goes to the dummy row) and
                // ends with the interval that completes its last filter, so no sum crosses warps.
                float up_prev = 0.0f;
                float2 a0 = make_float2(0.0f, 0.0f), a1 = make_float2(0.0f, 0.0f);
#define MEL_LOAD(G, Q0, Q1, Q2, Q3, W0, W1, idx)                                           \
    G = s_grp[idx];                                                                        \
    {                                                                                      \
        const float* pp_ = pl + G.x * kPPitch;                                             \
        Q0 = pp_[0]; Q1 = pp_[kPPitch]; Q2 = pp_[2 * kPPitch]; Q3 = pp_[3 * kPPitch];      \
        W0 = s_w4[2 * (idx)]; W1 = s_w4[2 * (idx) + 1];                                    \
    }
#define MEL_GROUP(G, Q0, Q1, Q2, Q3, W0, W1)                                               \
    a0 = __ffma2_rn(make_float2(W0.x, W0.y), bc2(Q0), a0);                                 \
    a1 = __ffma2_rn(make_float2(W0.z, W0.w), bc2(Q1), a1);                                 \
    a0 = __ffma2_rn(make_float2(W1.x, W1.y), bc2(Q2), a0);                                 \
    a1 = __ffma2_rn(make_float2(W1.z, W1.w), bc2(Q3), a1);                                 \
    if (G.y >= 0) { /* warp-uniform: the interval ends here */                             \
        const float2 a_ = __fadd2_rn(a0, a1);                                              \
        /* falling edge completes the filter below; a warp's first flush belongs to nobody (row kDummyRow): */ \
        /* not stored, or the warps would race on that row */                                  \
        if (G.y != kDummyRow) s_s[G.y * kSPitch + lane] = up_prev + a_.y;                    \
        up_prev = a_.x;                             /* rising edge waits for the next interval */ \
        a0 = make_float2(0.0f, 0.0f);                                                      \
        a1 = make_float2(0.0f, 0.0f);                                                      \
    }
                int2 ga, gb;
                float pa0, pa1, pa2, pa3, pb0, pb1, pb2, pb3;
                float4 wa0, wa1, wb0, wb1;
                int i = i0;
                MEL_LOAD(ga, pa0, pa1, pa2, pa3, wa0, wa1, i)
#pragma unroll 1
                for (; i + 1 < i1; i += 2) {
                    MEL_LOAD(gb, pb0, pb1, pb2, pb3, wb0, wb1, i + 1)
                    MEL_GROUP(ga, pa0, pa1, pa2, pa3, wa0, wa1)
                    MEL_LOAD(ga, pa0, pa1, pa2, pa3, wa0, wa1, i + 2)   // one past the range at the end: a valid slot
                    MEL_GROUP(gb, pb0, pb1, pb2, pb3, wb0, wb1)
                }
                if (i < i1) { MEL_GROUP(ga, pa0, pa1, pa2, pa3, wa0, wa1) }
#undef MEL_LOAD
#undef MEL_GROUP
            } else {
                for (int m = i0; m < i1; ++m) {
                    const MelItem it = tab.item[m];
                    const float* pp = pl + it.b0 * kPPitch;
                    const float* ww = p.weights + it.woff;
                    float a0 = 0.0f, a1 = 0.0f;
                    int k = 0;
                    for (; k + 1 < it.n; k += 2) {
                        a0 = fmaf(__ldg(ww + k), pp[k * kPPitch], a0);
                        a1 = fmaf(__ldg(ww + k + 1), pp[(k + 1) * kPPitch], a1);
                    }
                    if (k < it.n) a0 = fmaf(__ldg(ww + k), pp[k * kPPitch], a0);
                    s_s[m * kSPitch + lane] = a0 + a1;
                }
            }
        }
        __syncthreads();  // S complete (and a cooperatively copied span visible)

        // ================= output: log / clamp / affine, one 128-filter row per warp store =================
        {
            float mel[kRound / kWarps][4];
#pragma unroll
            for (int h = 0; h < kRound / kWarps; ++h)
#pragma unroll
                for (int i = 0; i < 4; ++i) mel[h][i] = s_s[(lane + 32 * i) * kSPitch + warp + h * kWarps];
#pragma unroll
            for (int h = 0; h < kRound / kWarps; ++h) {
                const int f = warp + h * kWarps;
                float* row = p.out + (geo.out_row + f) * p.n_mels;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float v = __logf(mel[h][i] + 1e-10f);               // |err| ~1e-7 in ln, 35x below the tolerance
                    v = v != v ? v : fminf(fmaxf(v, -23.0f), 12.0f);    // torch.clamp keeps NaN
                    if (f < geo.nf && lane + 32 * i < p.n_mels) row[lane + 32 * i] = (v + 23.0f) * (1.0f / 35.0f);
                }
            }
        }
        // S is rewritten only after the next round's first barrier, which every warp reaches after this point
    }
}


// =============================================================================================
// v6: warp-autonomous kernel.  Same arithmetic for the transform; what changes is the organisation:
// nothing is shared between the warps of a CTA after set-up - no CTA barrier, no common power matrix.
// A warp owns a *unit* of kUnitFrames consecutive kept frames of one segment (packed pairs, one after the other):
//  span    the unit's samples ((kUnitFrames-1)*hop + 2048 floats) arrive in the warp's own buffer by a TMA bulk
//          copy that the warp itself issues for its NEXT unit as soon as the last pass-1 loads of the current one
//          are consumed.  Longer units re-read less of the overlapping frames from L2 (measured: 2 frames per
//          unit 9.05 ms per step, 4 frames 8.27 ms)
//  FFT     as in v5, but the exchange tile and then the pair's power spectrum Q[bin] (float2 = the two
//          frames) live in an 8.4 KB warp-private buffer
//  mel     lanes are BIN RANGES: lane l walks bins 33*l .. 33*l+32 of Q once (conflict-free: stride 33)
//          with two running packed sums - `hi` for the filter whose rising edge the current interval is,
//          `lo` for the filter below (falling edge).  Where the interval changes inside the lane's range
//          (a per-lane bit mask) `lo` is complete for this lane and goes to the lane's next slot of a small
//          partial-sum array M, `lo <- hi`, `hi <- 0`.  A filter gets at most three such partial sums (its
//          two intervals span <= 48 bins), listed per filter in a table
//  output  lanes are filters (4 each): sum the partial sums in lane order, log / clamp / affine, store
// Warps drift apart freely, so the load-heavy and the FMA-heavy stretches of different warps overlap, which
// the lock-step phases of v5 prevented.  More warps do not help (8, 10 and 12 per SM measure the same): the
// kernel is bound by the shared-memory and fp32 pipes, not by latency.  Used when hop % 4 == 0, hop <= 256 and the filterbank is triangular
// with non-empty intervals (both shipped configurations); everything else takes the v5 kernel above.
#ifndef ADTFE_LM6_DIRECT
#define ADTFE_LM6_DIRECT 0    // 1: pass 1 reads the samples straight from global memory (no span buffer, no TMA)
#endif
#ifndef ADTFE_LM6_WARPS
#define ADTFE_LM6_WARPS 8
#endif
#ifndef ADTFE_LM6_UNIT
#define ADTFE_LM6_UNIT 6      // frames per unit (even): the span of a unit is (UNIT-1)*hop + 2048 samples
#endif
constexpr bool kDirect6 = ADTFE_LM6_DIRECT != 0;
constexpr int kWarps6 = ADTFE_LM6_WARPS;
constexpr int kUnitFrames = ADTFE_LM6_UNIT;
// floats of a warp's span buffer for units of `unit` frames: hop <= 256
__host__ __device__ constexpr int w6_span(int unit) { return kDirect6 ? 0 : (unit - 1) * 256 + 2048; }
constexpr int kLaneBins = 33;                 // bins walked by one lane: 32 * 33 = 1056 >= n_bins
constexpr int kQ2 = 32 * kLaneBins;           // float2 entries of a warp's exchange / power buffer
constexpr int kMMax = 200;                    // float2 partial sums per warp

struct Lane6 {        // the mel walk of one lane
    uint32_t mask_lo, mask_hi;   // bit i: the interval changes after the lane's i-th bin
    int32_t base;                // first M slot of the lane
    int32_t n_out;               // slots used (boundaries + 2)
};
struct Comb6 {        // one filter: its partial sums in lane order
    int16_t idx[3];
    int16_t n;
};
struct Logmel6Tables {
    const float2* w;      // [kLaneBins][32]: (rising, falling) weight of lane l's i-th bin
    const Lane6* lane;    // [32]
    const Comb6* comb;    // [4][32]: filter lane + 32*k
};

struct Unit6 {
    int seg, j0, nf, count;
    bool silent;   // the row is all zeros (an empty segment): its frames are exact zeros, no transform needed
    long long out_row;
};
template <int kUnitFrames>
__device__ __forceinline__ Unit6 unit6_geom(const LogmelArgs& p, int u, int units_per_seg) {
    Unit6 g;
    g.seg = (int)((unsigned)u / (unsigned)units_per_seg);
    g.j0 = (u - g.seg * units_per_seg) * kUnitFrames;
    int count = p.count;
    long long base = (long long)g.seg * p.count;
    g.silent = false;
    if (p.rows) {
        const int4 row = __ldg(reinterpret_cast<const int4*>(p.rows + g.seg));  // {out_row lo, hi, count, flags}
        base = (long long)(((unsigned long long)(unsigned)row.y << 32) | (unsigned)row.x);
        count = row.z;
        g.silent = (row.w & ADTFE_MEL_ROW_SILENT) != 0;
    }
    if (p.seg_mark && (__ldg(p.seg_mark + g.seg) != 0) != (p.mark_want != 0)) count = 0;
    g.nf = max(0, min(kUnitFrames, count - g.j0));
    g.count = count;
    g.out_row = base + g.j0;
    return g;
}
__device__ __forceinline__ void issue_span6(const LogmelArgs& p, const Unit6& g, float* span, uint64_t* bar, int lane) {
    if (lane == 0) {
        if (g.nf <= 0 || g.silent) {
            mbar_arrive(bar);
        } else {
            const float* src = p.wav + (long long)g.seg * p.ld_wav + (long long)(p.first + g.j0) * p.hop - 1024;
            const uint32_t bytes = (uint32_t)((g.nf - 1) * p.hop + 2048) * 4u;
            mbar_expect_tx(bar, bytes);
            bulk_g2s(span, src, bytes, bar);
        }
    }
}

// `tma`: every unit's samples start on a 16-byte boundary (aligned base, row pitch a multiple of 4 floats); otherwise the
// warp copies its span itself at the start of the unit (no prefetch) - same arithmetic, so results do not depend on it.
// torch.max propagates NaN (the mixer's tile maxima do too)
__device__ __forceinline__ float nan_max6(float a, float b) {
    return (a != a || b != b) ? __int_as_float(0x7fc00000) : fmaxf(a, b);
}

template <int kWarps6, int kUnitFrames>
__global__ void __launch_bounds__(kWarps6 * 32, 1) logmel6_kernel(const LogmelArgs p, const Logmel6Tables t6, int n_units,
                                                                int units_per_seg, int tma) {
    constexpr int kThreads6 = kWarps6 * 32;
    constexpr int kW6Span = w6_span(kUnitFrames);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* s_win = reinterpret_cast<float*>(smem_raw);                    // 32 * kWinPitch
    float2* s_tw = reinterpret_cast<float2*>(s_win + 32 * kWinPitch);     // 32*32
    float2* s_ltw = s_tw + 32 * 32;                                       // 4*32
    float2* s_w = s_ltw + 4 * 32;                                         // kLaneBins*32
    Lane6* s_lane = reinterpret_cast<Lane6*>(s_w + kLaneBins * 32);       // 32
    Comb6* s_comb = reinterpret_cast<Comb6*>(s_lane + 32);                // 128
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_comb + 128);          // kWarps6 (+ pad to 128 B)
    float* s_warp = reinterpret_cast<float*>(s_bar + 16);                 // per warp: span | Q | M

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    float* span = s_warp + warp * (kW6Span + 2 * kQ2 + 2 * kMMax);
    float2* qb = reinterpret_cast<float2*>(span + kW6Span);
    float2* mb = qb + kQ2;
    uint64_t* bar = s_bar + warp;

    if (lane == 0) mbar_init(bar, 1);
    for (int i = tid; i < 2048; i += kThreads6) s_win[(i & 31) * kWinPitch + (i >> 5)] = p.window[i];
    for (int i = tid; i < 32 * 32; i += kThreads6) s_tw[i] = p.twiddle[i];
    for (int i = tid; i < 4 * 32; i += kThreads6) s_ltw[i] = p.lane_tw[i];
    for (int i = tid; i < kLaneBins * 32; i += kThreads6) s_w[i] = t6.w[i];
    for (int i = tid; i < 32; i += kThreads6) s_lane[i] = t6.lane[i];
    for (int i = tid; i < 128; i += kThreads6) s_comb[i] = t6.comb[i];
    __syncthreads();  // the only CTA-wide barrier

    const int col32_bin = 32 + 64 * (int)(__brev((unsigned)lane) >> 27);
    const int gstride = (int)gridDim.x * kWarps6;
    int u = (int)blockIdx.x * kWarps6 + warp;
    Unit6 g = unit6_geom<kUnitFrames>(p, min(u, n_units - 1), units_per_seg);
    if (kDirect6) tma = 0;
    if (tma && u < n_units) issue_span6(p, g, span, bar, lane);
    uint32_t parity = 0;

#pragma unroll 1
    for (; u < n_units; u += gstride) {
        const bool more = u + gstride < n_units;
        Unit6 gn = g;
        if (more) gn = unit6_geom<kUnitFrames>(p, u + gstride, units_per_seg);
        if (tma) {
            mbar_wait(bar, parity);
            parity ^= 1u;
        } else if (!kDirect6 && g.nf > 0 && !g.silent) {
            const float* src = p.wav + (long long)g.seg * p.ld_wav + (long long)(p.first + g.j0) * p.hop - 1024;
            const int len = (g.nf - 1) * p.hop + 2048;
            __syncwarp();
            for (int i = lane; i < len; i += 32) span[i] = __ldg(src + i);
            __syncwarp();
        }
        const int n_pairs = g.silent ? 0 : (g.nf + 1) >> 1;
        if (g.silent) {  // log(0 + 1e-10) = -23.03 -> clamp -23 -> exactly 0.0, like the reference on silence
            for (int f = 0; f < g.nf; ++f) {
                float* row = p.out + (g.out_row + f) * p.n_mels;
                for (int m = lane; m < p.n_mels; m += 32) row[m] = 0.0f;
            }
        }
        if (tma && n_pairs == 0 && more) { __syncwarp(); issue_span6(p, gn, span, bar, lane); }
#pragma unroll 1
        for (int pr = 0; pr < n_pairs; ++pr) {
            const int f = 2 * pr;
            const int hop_b = (f + 1 < g.nf) ? p.hop : 0;  // an odd count: the last pair's second half repeats the first
            // ---- pass 1: window + 64-point real DFT over n1 (stride-32 samples), both frames
            float2 yr[33], yi[33];
            {
                const float* sa = kDirect6 ? p.wav + (long long)g.seg * p.ld_wav +
                                                 (long long)(p.first + g.j0 + f) * p.hop - 1024 + lane
                                           : span + f * p.hop + lane;
                const float* sb = sa + hop_b;
                const float4* wt = reinterpret_cast<const float4*>(s_win + lane * kWinPitch);
                float2 v[64];
                if (kDirect6) {  // all 128 global loads in flight, then the window
                    float xa[64], xb[64];
#pragma unroll
                    for (int n1 = 0; n1 < 64; ++n1) { xa[n1] = __ldg(sa + 32 * n1); xb[n1] = __ldg(sb + 32 * n1); }
#pragma unroll
                    for (int q = 0; q < 16; ++q) {
                        const float4 w = wt[q];
                        v[4 * q + 0] = __fmul2_rn(make_float2(xa[4 * q + 0], xb[4 * q + 0]), bc2(w.x));
                        v[4 * q + 1] = __fmul2_rn(make_float2(xa[4 * q + 1], xb[4 * q + 1]), bc2(w.y));
                        v[4 * q + 2] = __fmul2_rn(make_float2(xa[4 * q + 2], xb[4 * q + 2]), bc2(w.z));
                        v[4 * q + 3] = __fmul2_rn(make_float2(xa[4 * q + 3], xb[4 * q + 3]), bc2(w.w));
                    }
                } else {
#pragma unroll
                    for (int q = 0; q < 16; ++q) {
                        const float4 w = wt[q];
                        v[4 * q + 0] = __fmul2_rn(make_float2(sa[32 * (4 * q + 0)], sb[32 * (4 * q + 0)]), bc2(w.x));
                        v[4 * q + 1] = __fmul2_rn(make_float2(sa[32 * (4 * q + 1)], sb[32 * (4 * q + 1)]), bc2(w.y));
                        v[4 * q + 2] = __fmul2_rn(make_float2(sa[32 * (4 * q + 2)], sb[32 * (4 * q + 2)]), bc2(w.z));
                        v[4 * q + 3] = __fmul2_rn(make_float2(sa[32 * (4 * q + 3)], sb[32 * (4 * q + 3)]), bc2(w.w));
                    }
                }
                rdft64(v, yr, yi);
            }
            // ---- twiddle in place (rows 1..31), column 32 is real before its twiddle
#pragma unroll
            for (int k1 = 1; k1 < 32; ++k1) cmul2(yr[k1], yi[k1], s_tw[(k1 - 1) * 32 + lane]);
            float2 cr, ci;
            {
                const float2 w = s_tw[31 * 32 + lane];
                cr = __fmul2_rn(yr[32], bc2(w.x));
                ci = __fmul2_rn(yr[32], bc2(w.y));
            }
            // the unit's last span reads are consumed: fetch the next unit's samples behind the rest of this one
            if (tma && pr == n_pairs - 1 && more) { __syncwarp(); issue_span6(p, gn, span, bar, lane); }
            // ---- exchange through the warp's buffer: element (k1, n2) at 33*k1 + n2
            float2* xw = qb + lane;
            const float2* xr = qb + lane * kLaneBins;
            float2 zr[32], zi[32];
            __syncwarp();
#pragma unroll
            for (int k1 = 0; k1 < 32; ++k1) xw[k1 * kLaneBins] = yr[k1];
            __syncwarp();
#pragma unroll
            for (int n2 = 0; n2 < 32; ++n2) zr[n2] = xr[n2];
            __syncwarp();
            xw[0] = make_float2(0.0f, 0.0f);  // row 0 is purely real
#pragma unroll
            for (int k1 = 1; k1 < 32; ++k1) xw[k1 * kLaneBins] = yi[k1];
            __syncwarp();
#pragma unroll
            for (int n2 = 0; n2 < 32; ++n2) zi[n2] = xr[n2];
            __syncwarp();
            // ---- 32-point DFT across lanes for column 32
#pragma unroll
            for (int s = 0; s < 5; ++s) {
                const int hb = 16 >> s;
                const float2 orr = shfl_xor2(cr, hb), oi = shfl_xor2(ci, hb);
                const float sg = (lane & hb) ? -1.0f : 1.0f;
                cr = __ffma2_rn(cr, bc2(sg), orr);
                ci = __ffma2_rn(ci, bc2(sg), oi);
                if (s < 4) cmul2(cr, ci, s_ltw[s * 32 + lane]);
            }
            // ---- pass 2: 32-point complex DFT over n2 for k1 = lane
            cdft32(zr, zi);
            // ---- power spectrum of the pair into Q[bin] (over the dead exchange tile)
#pragma unroll
            for (int k2 = 0; k2 < 16; ++k2) qb[lane + 64 * k2] = __ffma2_rn(zi[k2], zi[k2], __fmul2_rn(zr[k2], zr[k2]));
#pragma unroll
            for (int k2 = 16; k2 < 32; ++k2)
                qb[64 - lane + 64 * (31 - k2)] = __ffma2_rn(zi[k2], zi[k2], __fmul2_rn(zr[k2], zr[k2]));
            if ((lane & 1) == 0) qb[col32_bin] = __ffma2_rn(ci, ci, __fmul2_rn(cr, cr));
            if (lane < kQ2 - 1025) qb[1025 + lane] = make_float2(0.0f, 0.0f);  // bins past the spectrum: weight 0, keep finite
            __syncwarp();
            // ---- mel: lane walks its 33 bins.  All power values and weights are requested first (the transform's
            // registers are dead by now), so the walk itself is a chain of packed FMAs without load stalls.
            {
                const float2* q = qb + lane * kLaneBins;
                const float2* w = s_w + lane;
                const Lane6 ln = s_lane[lane];
                float2 pw[kLaneBins], wt[kLaneBins];
#pragma unroll
                for (int i = 0; i < kLaneBins; ++i) { pw[i] = q[i]; wt[i] = w[i * 32]; }
                float2* mp = mb + ln.base;
                float2 lo = make_float2(0.0f, 0.0f), hi = make_float2(0.0f, 0.0f);
#pragma unroll
                for (int i = 0; i < kLaneBins; ++i) {
                    hi = __ffma2_rn(pw[i], bc2(wt[i].x), hi);
                    lo = __ffma2_rn(pw[i], bc2(wt[i].y), lo);
                    const bool cut = i < 32 ? ((ln.mask_lo >> i) & 1u) : ((ln.mask_hi >> (i - 32)) & 1u);
                    if (cut) {
                        *mp++ = lo;
                        lo = hi;
                        hi = make_float2(0.0f, 0.0f);
                    }
                }
                mp[0] = lo;
                mp[1] = hi;
                if (lane == 0) mb[kMMax - 1] = make_float2(0.0f, 0.0f);  // the slot unused list entries point to
            }
            __syncwarp();
            // ---- output: lane = filter (4 each): partial sums in lane order (missing ones read the zero slot),
            // log / clamp / affine on eight independent values
            {
                float* row_a = p.out + (g.out_row + f) * p.n_mels;
                const bool ok_a = f < g.nf, ok_b = f + 1 < g.nf;
                Comb6 c[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) c[k] = s_comb[k * 32 + lane];
                float2 v[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float2 m0 = mb[c[k].idx[0]], m1 = mb[c[k].idx[1]], m2 = mb[c[k].idx[2]];
                    v[k] = __fadd2_rn(__fadd2_rn(m0, m1), m2);
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int m = lane + 32 * k;
                    float a = __logf(v[k].x + 1e-10f), b = __logf(v[k].y + 1e-10f);
                    a = a != a ? a : fminf(fmaxf(a, -23.0f), 12.0f);    // torch.clamp keeps NaN
                    b = b != b ? b : fminf(fmaxf(b, -23.0f), 12.0f);
                    if (m < p.n_mels) {
                        if (ok_a) row_a[m] = (a + 23.0f) * (1.0f / 35.0f);
                        if (ok_b) row_a[p.n_mels + m] = (b + 23.0f) * (1.0f / 35.0f);
                    }
                }
            }
            __syncwarp();  // Q and M are rewritten by the next pair
        }
        g = gn;
    }
}

}  // namespace adtfe

using namespace adtfe;

static size_t logmel_smem_bytes() {
    return ((size_t)kSpanFloats + (size_t)kPRows * kPPitch + kSFloats + 4 * (size_t)(kW4Max + 2) + 32 * kWinPitch +
            2 * 32 * 32 + 2 * 4 * 32 + 2 * (kGroupsMax + 2)) * 4 + 16;
}

struct adtfe_mel_tables {
    MelTables t;
};

extern "C" int adtfe_mel_frames(const adtfe_mel* mel, int64_t n_samples, int32_t* first, int32_t* count) {
    ADTFE_REQUIRE(mel && first && count && n_samples >= 0, ADTFE_ERR_BAD_ARG, "adtfe_mel_frames: bad argument");
    const int64_t t_total = 1 + n_samples / mel->hop;  // centred STFT
    const int64_t c = t_total - 2 * (int64_t)mel->wpi - 1;
    *first = mel->wpi;
    *count = (int32_t)(c > 0 ? c : 0);
    return ADTFE_OK;
}

static size_t logmel6_smem_bytes(int warps, int unit) {
    return (size_t)(32 * kWinPitch + 2 * 32 * 32 + 2 * 4 * 32 + 2 * kLaneBins * 32) * 4 + 32 * sizeof(Lane6) +
           128 * sizeof(Comb6) + 128 + (size_t)warps * (w6_span(unit) + 2 * kQ2 + 2 * kMMax) * 4;
}

static int launch_logmel(const adtfe_mel* mel, const float* wav_dev, int32_t n_seg, int64_t ld_wav, int32_t first,
                         int32_t count, const adtfe_mel_row* rows_dev, float* out_dev, void* stream,
                         const int* seg_mark = nullptr, int mark_want = 0) {
    // frames per round: as many as fit the span buffer, at most 32; a segment's frames are split evenly
    const int cap = std::min(kRound, (kSpanFloats - 2048) / mel->hop + 1);
    const int rounds_per_seg = (count + cap - 1) / cap;
    const long long n_rounds = (long long)n_seg * rounds_per_seg;
    ADTFE_REQUIRE(n_rounds < (1ll << 30), ADTFE_ERR_UNSUPPORTED, "adtfe_logmel: too many frames for one launch");
    LogmelArgs a;
    a.wav = wav_dev; a.out = out_dev; a.window = mel->window; a.twiddle = mel->twiddle; a.lane_tw = mel->lane_tw;
    a.weights = mel->weights; a.groups = (const MelGroup*)mel->groups; a.rows = rows_dev; a.ld_wav = ld_wav;
    a.seg_mark = seg_mark; a.mark_want = mark_want;
    a.n_seg = n_seg; a.first = first; a.count = count;
    a.hop = mel->hop; a.n_mels = mel->n_mels; a.rounds_per_seg = rounds_per_seg; a.n_rounds = (int32_t)n_rounds;
    if (mel->v6_ok && !mel->force_generic) {
        // TMA needs 16-byte aligned unit starts; other rows are copied by the warps (same results)
        const int tma = ((uintptr_t)wav_dev & 15) == 0 && ld_wav % 4 == 0;
        const int unit = kUnitFrames, warps = kWarps6;
        const int units_per_seg = (count + unit - 1) / unit;
        const long long n_units = (long long)n_seg * units_per_seg;
        ADTFE_REQUIRE(n_units < (1ll << 30), ADTFE_ERR_UNSUPPORTED, "adtfe_logmel: too many frames for one launch");
        Logmel6Tables t6;
        t6.w = (const float2*)mel->w6; t6.lane = (const Lane6*)mel->lane6; t6.comb = (const Comb6*)mel->comb6;
        const long long ctas = (n_units + warps - 1) / warps;
        const int grid6 = (int)(ctas < mel->sm_count ? ctas : mel->sm_count);
        logmel6_kernel<kWarps6, kUnitFrames><<<grid6, kWarps6 * 32, mel->smem6_bytes, (cudaStream_t)stream>>>(
            a, t6, (int)n_units, units_per_seg, tma);
        ADTFE_CUDA(cudaGetLastError());
        return ADTFE_OK;
    }
    ADTFE_REQUIRE(!seg_mark, ADTFE_ERR_UNSUPPORTED, "adtfe_logmel: the row filter needs the warp-autonomous kernel");
    const int grid = (int)(n_rounds < mel->sm_count ? n_rounds : mel->sm_count);
    logmel_kernel<<<grid, kThreads, mel->smem_bytes, (cudaStream_t)stream>>>(a, mel->tables->t);
    ADTFE_CUDA(cudaGetLastError());
    return ADTFE_OK;
}

extern "C" int adtfe_logmel(const adtfe_mel* mel, const float* wav_dev, int32_t n_seg, int64_t ld_wav,
                            int64_t n_samples, float* out_dev, void* stream) {
    ADTFE_REQUIRE(mel && n_seg >= 0 && n_samples >= 0 && ld_wav >= n_samples, ADTFE_ERR_BAD_ARG,
                  "adtfe_logmel: bad argument (n_seg=%d ld_wav=%lld n_samples=%lld)", n_seg, (long long)ld_wav,
                  (long long)n_samples);
    int32_t first = 0, count = 0;
    adtfe_mel_frames(mel, n_samples, &first, &count);
    if (n_seg == 0 || count == 0) return ADTFE_OK;
    ADTFE_REQUIRE(wav_dev && out_dev, ADTFE_ERR_BAD_ARG, "adtfe_logmel: null buffer");
    ADTFE_REQUIRE(((uintptr_t)wav_dev & 3) == 0, ADTFE_ERR_BAD_ARG, "adtfe_logmel: wav_dev must be 4-byte aligned");
    // kept frames stay inside the signal by construction; guard against a caller-made mel
    ADTFE_REQUIRE((int64_t)first * mel->hop >= 1024 &&
                      (int64_t)(first + count - 1) * mel->hop + 1024 <= n_samples,
                  ADTFE_ERR_UNSUPPORTED, "adtfe_logmel: frame support leaves the signal");
    return launch_logmel(mel, wav_dev, n_seg, ld_wav, first, count, nullptr, out_dev, stream);
}

extern "C" int adtfe_logmel_rows(const adtfe_mel* mel, const float* wav_dev, int32_t n_seg, int64_t ld_wav,
                                 const adtfe_mel_row* rows_dev, int32_t max_count, float* out_dev, void* stream) {
    ADTFE_REQUIRE(mel && n_seg >= 0 && max_count >= 0 && ld_wav >= 0, ADTFE_ERR_BAD_ARG,
                  "adtfe_logmel_rows: bad argument (n_seg=%d ld_wav=%lld max_count=%d)", n_seg, (long long)ld_wav,
                  max_count);
    if (n_seg == 0 || max_count == 0) return ADTFE_OK;
    ADTFE_REQUIRE(wav_dev && out_dev && rows_dev, ADTFE_ERR_BAD_ARG, "adtfe_logmel_rows: null buffer");
    ADTFE_REQUIRE(((uintptr_t)wav_dev & 3) == 0 && ((uintptr_t)rows_dev & 15) == 0, ADTFE_ERR_BAD_ARG,
                  "adtfe_logmel_rows: wav_dev must be 4-byte and rows_dev 16-byte aligned");
    const int32_t first = mel->wpi;
    // the longest row must stay inside the pitch (the per-row counts live on the device: the caller's contract)
    ADTFE_REQUIRE((int64_t)first * mel->hop >= 1024 && (int64_t)(first + max_count - 1) * mel->hop + 1024 <= ld_wav,
                  ADTFE_ERR_UNSUPPORTED, "adtfe_logmel_rows: frame support leaves the row");
    return launch_logmel(mel, wav_dev, n_seg, ld_wav, first, max_count, rows_dev, out_dev, stream);
}

// The ragged form over a range of a plan's rows with a row filter (api.cu: the log-mel of a render chunk runs while
// later chunks are still being rendered; rows with an FX record follow when the FX chain is through).
bool adtfe::logmel_has_row_filter(const adtfe_mel* mel) { return mel && mel->v6_ok && !mel->force_generic; }
int adtfe::logmel_rows_filtered(const adtfe_mel* mel, const float* wav_dev, int32_t n_seg, int64_t ld_wav,
                                const adtfe_mel_row* rows_dev, int32_t max_count, float* out_dev, const int* seg_mark,
                                int mark_want, void* stream) {
    if (n_seg == 0 || max_count == 0) return ADTFE_OK;
    return launch_logmel(mel, wav_dev, n_seg, ld_wav, mel->wpi, max_count, rows_dev, out_dev, stream, seg_mark, mark_want);
}

extern "C" int adtfe_mel_destroy(adtfe_mel* mel) {
    if (!mel) return ADTFE_OK;
    cudaSetDevice(mel->device);
    cudaFree(mel->window); cudaFree(mel->twiddle); cudaFree(mel->lane_tw); cudaFree(mel->weights); cudaFree(mel->groups);
    cudaFree(mel->w6); cudaFree(mel->lane6); cudaFree(mel->comb6);
    delete mel->tables;
    delete mel;
    return ADTFE_OK;
}

// Triangular structure: every bin feeds at most two adjacent filters and the filter index never
// decreases with the bin.  Bins are then grouped into intervals j = 0..n_mels (interval j lies
// between the centres of filters j-1 and j) holding, per bin, the weight towards filter j ("up")
// and towards filter j-1 ("down"); intervals are cut into groups of kGroupBins bins.  Returns false when fb does not have that structure (or the
// packed weights do not fit): the caller then uses the per-filter path.
static bool triangular_structure(const float* fb, int n_bins, int n_mels, std::vector<int>& iv, std::vector<float>& up,
                                 std::vector<float>& dn) {
    iv.assign(n_bins, -1);
    up.assign(n_bins, 0.0f);
    dn.assign(n_bins, 0.0f);
    std::vector<int> peak(n_mels, 0);
    for (int m = 0; m < n_mels; ++m) {
        float best = -1.0f;
        for (int k = 0; k < n_bins; ++k)
            if (fb[(size_t)k * n_mels + m] > best) { best = fb[(size_t)k * n_mels + m]; peak[m] = k; }
    }
    int prev = 0;
    for (int k = 0; k < n_bins; ++k) {
        int a = -1, cnt = 0;
        for (int m = 0; m < n_mels; ++m)
            if (fb[(size_t)k * n_mels + m] != 0.0f) { if (cnt == 0) a = m; ++cnt; if (m - a > 1) return false; }
        if (cnt == 0) continue;
        if (cnt > 2) return false;
        int j;
        if (cnt == 2) {
            j = a + 1;
            dn[k] = fb[(size_t)k * n_mels + a];
            up[k] = fb[(size_t)k * n_mels + a + 1];
        } else if (a >= prev && k <= peak[a]) {  // rising edge of filter a
            j = a;
            up[k] = fb[(size_t)k * n_mels + a];
        } else {                                 // falling edge of filter a
            j = a + 1;
            dn[k] = fb[(size_t)k * n_mels + a];
        }
        if (j < prev) return false;
        iv[k] = j;
        prev = j;
    }
    return true;
}

static bool build_fast_tables(const float* fb, int n_bins, int n_mels, MelTables& t, std::vector<float>& packed,
                              std::vector<MelGroup>& groups) {
    std::vector<int> iv;
    std::vector<float> up, dn;
    if (!triangular_structure(fb, n_bins, n_mels, iv, up, dn)) return false;
    // warp ranges over the filters: warp w completes filters m0 .. m1-1, i.e. walks intervals m0 .. m1 (interval
    // m0 only for its rising edge, interval m1 only for its falling edge); the boundary intervals are walked
    // twice so that no sum crosses warps.  Balanced by the bins per range.
    std::vector<int> iv_bins(n_mels + 2, 0);
    for (int k = 0; k < n_bins; ++k)
        if (iv[k] >= 0) ++iv_bins[iv[k]];
    auto iv_groups = [&](int j) { return std::max(1, (iv_bins[j] + kGroupBins - 1) / kGroupBins); };
    std::vector<int> m_of_warp(kWarps + 1, 0);
    {
        std::vector<double> cost(n_mels, 0.0);
        double total = 0;
        for (int m = 0; m < n_mels; ++m) total += cost[m] = 16.0 * iv_groups(m + 1) + 10.0;
        int m = 0;
        double spent = 0;
        for (int w = 0; w < kWarps; ++w) {
            m_of_warp[w] = m;
            const double target = total * (w + 1) / kWarps;
            while (m < n_mels && (w == kWarps - 1 || spent + 0.5 * cost[m] <= target)) spent += cost[m++];
        }
        m_of_warp[kWarps] = n_mels;
    }
    // intervals -> groups of kGroupBins bins; an interval without bins still gets one (all-zero) group,
    // because its end is where the filter below it is completed
    packed.clear();
    groups.clear();
    memset(&t, 0, sizeof(t));
    std::vector<float> chk((size_t)n_bins * n_mels, 0.0f);
    for (int w = 0; w < kWarps; ++w) {
        t.first[w] = (int16_t)groups.size();
        const int m0 = m_of_warp[w], m1 = m_of_warp[w + 1];
        for (int j = m0; m1 > m0 && j <= m1; ++j) {
            int lo = -1, hi = -1;
            for (int k = 0; k < n_bins; ++k)
                if (iv[k] == j) { if (lo < 0) lo = k; hi = k; }
            for (int k = lo; k <= hi; ++k)
                if (lo >= 0 && iv[k] != j && iv[k] >= 0) return false;  // bins of another interval inside the range
            int g0 = lo < 0 ? 0 : lo;
            do {
                MelGroup g;
                g.b0 = std::min(g0, n_bins + 1 - kGroupBins);  // rows 0 .. n_bins exist (row n_bins is kept at zero)
                g.row = -1;
                for (int e = 0; e < kGroupBins; ++e) {
                    const int k = g.b0 + e;
                    const bool in = lo >= 0 && k >= g0 && k < g0 + kGroupBins && k <= hi && iv[k] == j;
                    const float wu = in ? up[k] : 0.0f, wd = in ? dn[k] : 0.0f;
                    packed.push_back(wu);
                    packed.push_back(wd);
                    // what the kernel will use of this group: the rising edge unless it is the range's last
                    // interval, the falling edge unless it is the range's first
                    if (in && wu != 0.0f && j < m1) { if (j >= n_mels) return false; chk[(size_t)k * n_mels + j] += wu; }
                    if (in && wd != 0.0f && j > m0) chk[(size_t)k * n_mels + j - 1] += wd;
                    if (in && wd != 0.0f && j == 0) return false;
                }
                groups.push_back(g);
                g0 += kGroupBins;
            } while (lo >= 0 && g0 <= hi);
            groups.back().row = j > m0 ? j - 1 : kDummyRow;  // the end of interval j completes filter j-1
        }
    }
    t.first[kWarps] = (int16_t)groups.size();
    if (groups.size() > (size_t)kGroupsMax) return false;
    // verify: the groups reproduce fb exactly
    for (size_t i = 0; i < chk.size(); ++i)
        if (chk[i] != fb[i]) return false;
    t.fast = 1;
    t.n_groups = (int32_t)groups.size();
    return true;
}

static void build_generic_tables(const float* fb, int n_bins, int n_mels, MelTables& t, std::vector<float>& packed) {
    packed.clear();
    memset(&t, 0, sizeof(t));
    std::vector<double> cost(n_mels, 0.0);
    for (int m = 0; m < n_mels; ++m) {
        int first = -1, last = -1;
        for (int k = 0; k < n_bins; ++k)
            if (fb[(size_t)k * n_mels + m] != 0.0f) { if (first < 0) first = k; last = k; }
        MelItem it;
        it.b0 = (int16_t)(first < 0 ? 0 : first);
        it.n = (int16_t)(first < 0 ? 0 : last - first + 1);
        it.woff = (int32_t)packed.size();
        for (int k = first; first >= 0 && k <= last; ++k) packed.push_back(fb[(size_t)k * n_mels + m]);
        t.item[m] = it;
        cost[m] = 2.5 * it.n + 8.0;
    }
    double total = 0;
    for (int m = 0; m < n_mels; ++m) total += cost[m];
    int m = 0;
    double spent = 0;
    for (int w = 0; w < kWarps; ++w) {
        t.first[w] = (int16_t)m;
        const double target = total * (w + 1) / kWarps;
        while (m < n_mels && (w == kWarps - 1 || spent + 0.5 * cost[m] <= target)) spent += cost[m++];
    }
    t.first[kWarps] = (int16_t)n_mels;
    t.fast = 0;
    t.n_groups = 0;
}

// Tables of the warp-autonomous kernel (v6): per lane the (rising, falling) weights of its 33 bins, the bit mask
// of interval changes and its first partial-sum slot; per filter the slots that add up to it.  Returns false
// when the filterbank does not fit (not triangular, an interval without bins, too many partial sums).
static bool build_lane_tables(const float* fb, int n_bins, int n_mels, std::vector<float2>& w, std::vector<Lane6>& lanes,
                              std::vector<Comb6>& comb) {
    if (n_bins > kQ2 - 1 || n_mels > kMaxMels) return false;
    std::vector<int> iv;
    std::vector<float> up, dn;
    if (!triangular_structure(fb, n_bins, n_mels, iv, up, dn)) return false;
    // bins without weights take the interval of their neighbour, so that they never cut a run
    std::vector<int> j_of(kQ2, -1);
    int last = -1;
    for (int k = 0; k < n_bins; ++k) { if (iv[k] >= 0) last = iv[k]; j_of[k] = last; }
    if (last < 0) return false;
    int first_valid = 0;
    while (j_of[first_valid] < 0) ++first_valid;
    for (int k = 0; k < first_valid; ++k) j_of[k] = j_of[first_valid];
    for (int k = n_bins; k < kQ2; ++k) j_of[k] = last;
    w.assign((size_t)kLaneBins * 32, make_float2(0.0f, 0.0f));
    lanes.assign(32, Lane6{0u, 0u, 0, 0});
    std::vector<std::vector<int>> slots_of(n_mels);  // filter -> M slots, in lane order
    int n_slots = 0;
    for (int l = 0; l < 32; ++l) {
        Lane6& ln = lanes[l];
        ln.base = n_slots;
        int filt = j_of[l * kLaneBins] - 1;  // the filter `lo` stands for
        for (int i = 0; i < kLaneBins; ++i) {
            const int k = l * kLaneBins + i;
            if (k < n_bins) w[(size_t)i * 32 + l] = make_float2(up[k], dn[k]);
            if (i + 1 < kLaneBins && j_of[k + 1] != j_of[k]) {
                if (j_of[k + 1] != j_of[k] + 1) return false;  // an interval without bins
                if (i < 32) ln.mask_lo |= 1u << i; else ln.mask_hi |= 1u << (i - 32);
                if (filt >= 0 && filt < n_mels) slots_of[filt].push_back(n_slots);
                ++n_slots;
                ++filt;
            }
        }
        for (int e = 0; e < 2; ++e, ++filt, ++n_slots)  // what is left at the end of the range: lo, then hi
            if (filt >= 0 && filt < n_mels) slots_of[filt].push_back(n_slots);
        ln.n_out = n_slots - ln.base;
    }
    if (n_slots > kMMax - 1) return false;  // the last slot is kept at zero for unused list entries
    comb.assign(128, Comb6{{(int16_t)(kMMax - 1), (int16_t)(kMMax - 1), (int16_t)(kMMax - 1)}, 0});
    for (int m = 0; m < n_mels; ++m) {
        if (slots_of[m].empty() || slots_of[m].size() > 3) return false;
        Comb6& c = comb[(size_t)(m >> 5) * 32 + (m & 31)];
        c.n = (int16_t)slots_of[m].size();
        for (size_t e = 0; e < slots_of[m].size(); ++e) c.idx[e] = (int16_t)slots_of[m][e];
    }
    // verify by replaying the walk symbolically: every weight must land on the filter fb has it on
    std::vector<float> chk((size_t)n_bins * n_mels, 0.0f);
    std::vector<int> filter_of_slot(n_slots, -1);
    for (int m = 0; m < n_mels; ++m)
        for (int sl : slots_of[m]) filter_of_slot[sl] = m;
    for (int l = 0; l < 32; ++l) {
        int slot = lanes[l].base;
        std::vector<int> lo_bins, hi_bins;  // bins currently summed in lo (with dn) / hi (with up)
        auto emit = [&](const std::vector<int>& dnb, const std::vector<int>& upb) {
            const int m = filter_of_slot[slot];
            for (int k : dnb) { if (dn[k] != 0.0f) { if (m < 0) return false; chk[(size_t)k * n_mels + m] += dn[k]; } }
            for (int k : upb) { if (up[k] != 0.0f) { if (m < 0) return false; chk[(size_t)k * n_mels + m] += up[k]; } }
            ++slot;
            return true;
        };
        std::vector<int> lo_up;  // bins whose `up` weight moved from hi into lo at a cut
        for (int i = 0; i < kLaneBins; ++i) {
            const int k = l * kLaneBins + i;
            if (k < n_bins) { hi_bins.push_back(k); lo_bins.push_back(k); }
            const bool cut = i < 32 ? (lanes[l].mask_lo >> i) & 1u : (lanes[l].mask_hi >> (i - 32)) & 1u;
            if (cut) {
                if (!emit(lo_bins, lo_up)) return false;
                lo_up = hi_bins;     // lo <- hi: the rising-edge sums so far
                lo_bins.clear();     // and no falling-edge sums yet for the new lo
                hi_bins.clear();
            }
        }
        if (!emit(lo_bins, lo_up)) return false;
        if (!emit(std::vector<int>(), hi_bins)) return false;
    }
    for (size_t i = 0; i < chk.size(); ++i)
        if (chk[i] != fb[i]) return false;
    return true;
}

extern "C" int adtfe_mel_create(int32_t n_fft, int32_t hop, int32_t n_mels, const float* window_host,
                                const float* fb_host, int device, adtfe_mel** out) {
    ADTFE_REQUIRE(out && window_host && fb_host, ADTFE_ERR_BAD_ARG, "adtfe_mel_create: null pointer");
    *out = nullptr;
    ADTFE_REQUIRE(n_fft == 2048, ADTFE_ERR_UNSUPPORTED, "adtfe_mel_create: n_fft %d unsupported (only 2048)", n_fft);
    ADTFE_REQUIRE(hop >= 1 && hop <= 2048, ADTFE_ERR_UNSUPPORTED, "adtfe_mel_create: hop %d out of range", hop);
    ADTFE_REQUIRE(n_mels >= 1 && n_mels <= kMaxMels, ADTFE_ERR_UNSUPPORTED,
                  "adtfe_mel_create: n_mels %d out of range (1..%d)", n_mels, kMaxMels);
    int rc = adtfe_device_ok(device);
    if (rc != ADTFE_OK) return rc;
    ADTFE_CUDA(cudaSetDevice(device));

    const int n_bins = n_fft / 2 + 1;
    std::vector<float2> tw(32 * 32), ltw(4 * 32);
    const double two_pi = 6.283185307179586476925286766559;
    for (int k1 = 1; k1 <= 32; ++k1)
        for (int n2 = 0; n2 < 32; ++n2) {
            const double a = -two_pi * (double)(k1 * n2) / 2048.0;
            tw[(k1 - 1) * 32 + n2] = make_float2((float)std::cos(a), (float)std::sin(a));
        }
    for (int s = 0; s < 4; ++s) {  // stage 3 (half = 2): 1 or -i, exact
        const int half = 16 >> s;
        for (int l = 0; l < 32; ++l) {
            if (s == 3) {
                ltw[s * 32 + l] = ((l & half) && (l & 1)) ? make_float2(0.0f, -1.0f) : make_float2(1.0f, 0.0f);
            } else if (l & half) {
                const double a = -two_pi * (double)(l & (half - 1)) / (double)(2 * half);
                ltw[s * 32 + l] = make_float2((float)std::cos(a), (float)std::sin(a));
            } else {
                ltw[s * 32 + l] = make_float2(1.0f, 0.0f);
            }
        }
    }

    adtfe_mel* mel = new adtfe_mel();
    mel->device = device; mel->n_fft = n_fft; mel->hop = hop; mel->n_mels = n_mels;
    mel->wpi = (n_fft / 2) / hop + 1;  // int((win/2)//hop + 1), model.py:79
    mel->sm_count = device_sm_count(device);
    mel->smem_bytes = logmel_smem_bytes();
    mel->tables = new adtfe_mel_tables();
    std::vector<float> packed;
    std::vector<MelGroup> groups;
    if (!build_fast_tables(fb_host, n_bins, n_mels, mel->tables->t, packed, groups)) {
        groups.clear();
        build_generic_tables(fb_host, n_bins, n_mels, mel->tables->t, packed);
    }
    mel->fast_path = mel->tables->t.fast;
    mel->nnz = 0;
    for (size_t i = 0; i < (size_t)n_bins * n_mels; ++i) mel->nnz += fb_host[i] != 0.0f;
    auto fail = [&](int status) { adtfe_mel_destroy(mel); return status; };
#define MEL_UPLOAD(dst, src, bytes)                                                                  \
    if (cudaMalloc((void**)&(dst), (bytes) ? (bytes) : 16) != cudaSuccess ||                         \
        cudaMemcpy((dst), (src), (bytes), cudaMemcpyHostToDevice) != cudaSuccess) {                  \
        set_error("adtfe_mel_create: upload failed: %s", cudaGetErrorString(cudaGetLastError()));    \
        return fail(ADTFE_ERR_CUDA);                                                                 \
    }
    MEL_UPLOAD(mel->window, window_host, (size_t)n_fft * 4);
    MEL_UPLOAD(mel->twiddle, tw.data(), tw.size() * sizeof(float2));
    MEL_UPLOAD(mel->lane_tw, ltw.data(), ltw.size() * sizeof(float2));
    MEL_UPLOAD(mel->weights, packed.data(), packed.size() * 4);
    MEL_UPLOAD(mel->groups, groups.data(), groups.size() * sizeof(MelGroup));
    {
        std::vector<float2> w6;
        std::vector<Lane6> lane6;
        std::vector<Comb6> comb6;
        mel->v6_ok = hop % 4 == 0 && (kUnitFrames - 1) * hop + 2048 <= w6_span(kUnitFrames) && ((n_fft / 2) / hop + 1) * hop >= 1024 &&
                     build_lane_tables(fb_host, n_bins, n_mels, w6, lane6, comb6);
        if (mel->v6_ok) {
            MEL_UPLOAD(mel->w6, w6.data(), w6.size() * sizeof(float2));
            MEL_UPLOAD(mel->lane6, lane6.data(), lane6.size() * sizeof(Lane6));
            MEL_UPLOAD(mel->comb6, comb6.data(), comb6.size() * sizeof(Comb6));
            mel->smem6_bytes = logmel6_smem_bytes(kWarps6, kUnitFrames);
            if (cudaFuncSetAttribute(logmel6_kernel<kWarps6, kUnitFrames>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)mel->smem6_bytes) != cudaSuccess) {
                cudaGetLastError();
                mel->v6_ok = 0;
            }
        }
    }
#undef MEL_UPLOAD
    if (cudaFuncSetAttribute(logmel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mel->smem_bytes) !=
        cudaSuccess) {
        set_error("adtfe_mel_create: cannot reserve %zu B of shared memory: %s", mel->smem_bytes,
                  cudaGetErrorString(cudaGetLastError()));
        return fail(ADTFE_ERR_CUDA);
    }
    *out = mel;
    return ADTFE_OK;
}

extern "C" int adtfe_mel_fast_path(const adtfe_mel* mel) { return mel ? mel->fast_path : -1; }

extern "C" int adtfe_mel_force_generic(adtfe_mel* mel, int32_t on) {
    ADTFE_REQUIRE(mel, ADTFE_ERR_BAD_ARG, "adtfe_mel_force_generic: null handle");
    mel->force_generic = on != 0;
    return ADTFE_OK;
}
