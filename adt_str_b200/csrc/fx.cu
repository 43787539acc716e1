// FX chain on the raw mix (sm_100a): Freeverb reverb, feed-forward compressor, two-stage limiter.
//
// Replaces VolumeMixer._add_fx / BoardChain (reference modules/synthetiser.py:30-87, 121-137), applied at :154
// BETWEEN the instrument sum and the global normalisation, i.e. between mix_kernel and normalise_kernel here.  The
// reference builds pedalboard plugins (JUCE DSP): pedalboard.Reverb = juce::Reverb::processMono, pedalboard.Compressor
// = juce::dsp::Compressor<float>, pedalboard.Limiter = juce::dsp::Limiter<float>.  pedalboard is third-party and not
// part of the reference tree; the arithmetic below follows the published JUCE sources (restated sample by sample in
// oracle/fx_oracle.c, which is what the tests compare against - parity against pedalboard itself is unpinned).
//
// The filters are recursions over the 61 440+ samples of a row, so the parallelism is across what IS independent:
//
// fx_reverb_kernel    one warp per row, 32 consecutive samples per step with lanes = samples.  Every delay line is
//                     longer than 32 samples (checked on the host), so inside a step nothing feeds back through a
//                     delay line; what remains sequential is the one-pole lowpass in each comb's feedback path,
//                     last[n] = y[n] (1 - damp) + damp last[n-1]: a first-order linear recurrence with a constant
//                     coefficient, solved across the lanes by a 5-step shuffle scan.  The 8 combs and then the 4
//                     series allpasses (each parallel over the step) run from 27 KB of delay lines in shared memory.
// fx_dynamics_kernel  lane = row (32 rows of the plan's FX list per warp).  An envelope follower
//                     y = |x| + (|x| > y ? cAT : cRL)(y - |x|) has a data-dependent coefficient, so it cannot be
//                     scanned: each lane walks its own row.  The three followers of the chain (compressor, limiter
//                     stage 1, limiter stage 2) are SKEWED by a block of four samples each - a trip runs stage 1 on
//                     block b, stage 2 on block b-1, stage 3 on block b-2 - so the loop-carried dependency is one
//                     follower (5 dependent instructions per sample), not three followers and two pow() in series, and
//                     the pow() themselves are exp2(e * log2(x)) on the special-function unit.  Disabled stages are
//                     the same code with an infinite threshold (gain 1), so rows with different chains share a warp.
//                     The lane also tracks max|out| of its row: the row peak the normalisation needs.
#include <math.h>

#include <algorithm>

#include "common.cuh"

namespace adtfe {

constexpr int kCombTuning[8] = {1116, 1188, 1277, 1356, 1422, 1491, 1557, 1617};   // juce_Reverb.h, at 44.1 kHz
constexpr int kAllpassTuning[4] = {556, 441, 341, 225};

struct ReverbGeom {
    int32_t comb_size[8], comb_off[8];
    int32_t ap_size[4], ap_off[4];
    int32_t total;   // floats of delay lines
};

static ReverbGeom reverb_geom(int sample_rate) {
    ReverbGeom g;
    int off = 0;
    for (int j = 0; j < 8; ++j) { g.comb_size[j] = (sample_rate * kCombTuning[j]) / 44100; g.comb_off[j] = off; off += g.comb_size[j]; }
    for (int j = 0; j < 4; ++j) { g.ap_size[j] = (sample_rate * kAllpassTuning[j]) / 44100; g.ap_off[j] = off; off += g.ap_size[j]; }
    g.total = off;
    return g;
}

// JUCE_UNDENORMALISE on Intel builds: x += 0.1f; x -= 0.1f
__device__ __forceinline__ float undenorm(float x) { return __fadd_rn(__fadd_rn(x, 0.1f), -0.1f); }

__global__ void __launch_bounds__(32) fx_reverb_kernel(const adtfe_segment* __restrict__ segments,
                                                       const adtfe_fx* __restrict__ fx, float* __restrict__ wav,
                                                       int64_t ld_wav, const ReverbGeom g) {
    extern __shared__ float s_delay[];
    const int lane = threadIdx.x;
    const adtfe_fx f = fx[blockIdx.x];   // one warp per row of the FX list
    const int seg = f.seg;
    const adtfe_segment sg = segments[seg];
    if (!(f.flags & ADTFE_FX_REVERB) || sg.flags == 0) return;
    for (int i = lane; i < g.total; i += 32) s_delay[i] = 0.0f;
    __syncwarp();
    const float wet = f.wet_level * 3.0f, dry = f.dry_level * 2.0f;
    const float wet1 = 0.5f * wet * (1.0f + f.width);
    const float damp = f.damping * 0.4f, feedback = f.room_size * 0.28f + 0.7f;
    const float one_minus_damp = 1.0f - damp;
    // damp^(2^k) for the scan, damp^(lane + 1) to carry the previous step's state in
    float dpow[5];
    dpow[0] = damp;
#pragma unroll
    for (int k = 1; k < 5; ++k) dpow[k] = dpow[k - 1] * dpow[k - 1];
    float dcarry = damp;
    for (int i = 0; i < lane; ++i) dcarry *= damp;
    float last[8];
    int cpos[8], apos[4];
#pragma unroll
    for (int j = 0; j < 8; ++j) { last[j] = 0.0f; cpos[j] = 0; }
#pragma unroll
    for (int j = 0; j < 4; ++j) apos[j] = 0;
    float* row = wav + (int64_t)seg * ld_wav;
    const int n = sg.len;
    for (int n0 = 0; n0 < n; n0 += 32) {
        const bool valid = n0 + lane < n;
        const float x = valid ? row[n0 + lane] : 0.0f;
        const float input = x * 0.015f;
        float out = 0.0f;
        // the eight combs in lock step: their scans are independent chains of shuffles, interleaved they cost the
        // latency of one
        float* slot[8];
        float y[8], v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            int idx = cpos[j] + lane;
            if (idx >= g.comb_size[j]) idx -= g.comb_size[j];
            slot[j] = s_delay + g.comb_off[j] + idx;
            y[j] = *slot[j];
            v[j] = y[j] * one_minus_damp;   // last[i] = y[i] (1 - damp) + damp last[i-1]: inclusive scan with ratio damp
        }
#pragma unroll
        for (int k = 0; k < 5; ++k) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float t = __shfl_up_sync(0xffffffffu, v[j], 1 << k);
                if (lane >= (1 << k)) v[j] = fmaf(t, dpow[k], v[j]);
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float l = undenorm(fmaf(dcarry, last[j], v[j]));   // + the carry of the previous step
            *slot[j] = undenorm(__fadd_rn(input, __fmul_rn(l, feedback)));
            out += y[j];
            last[j] = __shfl_sync(0xffffffffu, l, 31);
            cpos[j] += 32;
            if (cpos[j] >= g.comb_size[j]) cpos[j] -= g.comb_size[j];
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int idx = apos[j] + lane;
            if (idx >= g.ap_size[j]) idx -= g.ap_size[j];
            float* slot = s_delay + g.ap_off[j] + idx;
            const float buffered = *slot;
            *slot = undenorm(__fadd_rn(out, __fmul_rn(buffered, 0.5f)));
            out = buffered - out;
            apos[j] += 32;
            if (apos[j] >= g.ap_size[j]) apos[j] -= g.ap_size[j];
        }
        if (valid) row[n0 + lane] = __fadd_rn(__fmul_rn(out, wet1), __fmul_rn(x, dry));
        __syncwarp();   // the delay lines are read by other lanes in the next step
    }
}

struct Follower {   // one compressor stage of one row
    float thr, thr_inv, expo, at, rl, y;
};

__device__ __forceinline__ Follower make_follower(bool on, double exp_factor, float threshold_db, float ratio,
                                                  float attack_ms, float release_ms) {
    Follower c;
    c.thr = on ? (threshold_db > -200.0f ? powf(10.0f, threshold_db * 0.05f) : 0.0f) : __int_as_float(0x7f800000);
    c.thr_inv = 1.0f / c.thr;
    c.expo = 1.0f / ratio - 1.0f;
    c.at = attack_ms < 1.0e-3f ? 0.0f : (float)exp(exp_factor / (double)attack_ms);
    c.rl = release_ms < 1.0e-3f ? 0.0f : (float)exp(exp_factor / (double)release_ms);
    c.y = 0.0f;
    return c;
}

// one sample through a stage: returns the stage's output, advances the follower.  The gain pow(env / thr, 1/ratio - 1)
// is exp2(expo * log2(.)) on the special-function unit (relative error ~3e-7, far inside the waveform tolerance) and
// branch-free, so that the three stages of the chain interleave: the loop-carried dependency is the follower alone.
__device__ __forceinline__ float follower_step(Follower& c, float in) {
    const float rect = fabsf(in);
    const float cte = rect > c.y ? c.at : c.rl;
    const float env = __fadd_rn(rect, __fmul_rn(cte, c.y - rect));   // no contraction, like the x86 builds of JUCE
    c.y = env;
    float lg, p;   // the bare special-function instructions (no denormal / range fix-ups: the argument is a ratio >= 1)
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"(env * c.thr_inv));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p) : "f"(c.expo * lg));
    const float gain = env < c.thr ? 1.0f : p;                   // NaN envelopes take the pow value (NaN), like the C++
    return gain * in;
}

__global__ void __launch_bounds__(32) fx_dynamics_kernel(const adtfe_segment* __restrict__ segments,
                                                         const adtfe_fx* __restrict__ fx, int n_fx, float* __restrict__ wav,
                                                         int64_t ld_wav, float* __restrict__ tile_max, int max_per_seg,
                                                         int sample_rate) {
    const int r = blockIdx.x * 32 + threadIdx.x;   // lane = one row of the FX list
    if (r >= n_fx) return;                          // no warp-wide operation below
    const adtfe_fx f = fx[r];
    const int seg = f.seg;
    if (f.flags == 0 || segments[seg].flags == 0) return;
    const double exp_factor = -2.0 * 3.14159265358979323846 * 1000.0 / (double)sample_rate;
    const bool comp_on = (f.flags & ADTFE_FX_COMPRESSOR) != 0, lim_on = (f.flags & ADTFE_FX_LIMITER) != 0;
    Follower c1 = make_follower(comp_on, exp_factor, f.comp_threshold_db, f.comp_ratio, f.comp_attack_ms,
                                f.comp_release_ms);
    Follower c2 = make_follower(lim_on, exp_factor, -10.0f, 4.0f, 2.0f, 200.0f);
    Follower c3 = make_follower(lim_on, exp_factor, f.lim_threshold_db, 1000.0f, 0.001f, 100.0f);
    float out_gain = 1.0f, clip = __int_as_float(0x7f800000);
    if (lim_on) {
        out_gain = (float)pow(10.0, 10.0 * (1.0 - (double)0.25f) / 40.0);
        out_gain *= f.lim_threshold_db < 100.0f ? powf(10.0f, -f.lim_threshold_db * 0.05f) : 0.0f;
        clip = 1.0f;
    }
    float* row = wav + (int64_t)seg * ld_wav;
    const int n = segments[seg].len;
    float peak = 0.0f;
    const bool rewrite = lim_on || comp_on;
    // Four samples (one 16-byte load) per trip, the stages skewed by a whole trip: trip b runs stage 1 on block b,
    // stage 2 on stage 1's output of block b - 1 and stage 3 on stage 2's output of block b - 2.  The three followers
    // are then independent chains inside a trip (and the gains, off the chains, pipeline through the special-function
    // unit); block b - 2 leaves as one aligned 16-byte store (the row starts on a 16-byte boundary: ld_wav % 4 == 0).
    // The loads run a whole 128-byte line (eight trips) ahead: the in-place stores invalidate the line in L1, so
    // every load is an L2 round trip - issued one line early it is never waited for.
    float s1[4] = {0.0f, 0.0f, 0.0f, 0.0f}, s2[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    const float4* row4 = reinterpret_cast<const float4*>(row);
    const int n_blocks = (n + 3) / 4, pitch_blocks = (int)(ld_wav / 4);
    const float4 zero4 = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    float4 cur[8], nxt[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) cur[t] = t < pitch_blocks ? row4[t] : zero4;
    for (int line = 0; 8 * line < n_blocks + 2; ++line) {
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            const int blk = 8 * (line + 1) + t;
            nxt[t] = blk < n_blocks && blk < pitch_blocks ? row4[blk] : zero4;
        }
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            const int b = 8 * line + t;
            float xin[4] = {cur[t].x, cur[t].y, cur[t].z, cur[t].w};
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (4 * b + k >= n) xin[k] = 0.0f;   // the row's padding is not part of the signal
            float t1[4], t2[4], o[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) t1[k] = follower_step(c1, xin[k]);
#pragma unroll
            for (int k = 0; k < 4; ++k) t2[k] = follower_step(c2, s1[k]);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float v = follower_step(c3, s2[k]) * out_gain;
                o[k] = v < -clip ? -clip : (v > clip ? clip : v);   // keeps NaN, like FloatVectorOperations::clip
            }
            if (b >= 2 && 4 * (b - 2) < n) {
                const int s0 = 4 * (b - 2);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (s0 + k < n) {
                        const float a = fabsf(o[k]);
                        peak = (a != a || peak != peak) ? __int_as_float(0x7fc00000) : fmaxf(peak, a);   // torch.max keeps NaN
                    }
                }
                if (rewrite) {
                    if (s0 + 3 < n) reinterpret_cast<float4*>(row)[b - 2] = make_float4(o[0], o[1], o[2], o[3]);
                    else {
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            if (s0 + k < n) row[s0 + k] = o[k];
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) { s1[k] = t1[k]; s2[k] = t2[k]; }
        }
#pragma unroll
        for (int t = 0; t < 8; ++t) cur[t] = nxt[t];
    }
    // the normalisation takes the row peak from the tile maxima: this row's is now `peak`
    float* tm = tile_max + (size_t)seg * max_per_seg;
    tm[0] = peak;
    for (int t = 1; t < max_per_seg; ++t) tm[t] = 0.0f;
}

int fx_prepare_device() {
    // the largest delay-line set this build accepts: 48 kHz (55 KB)
    ADTFE_CUDA(cudaFuncSetAttribute(fx_reverb_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    return ADTFE_OK;
}

// FX of the rows fx_dev[r0 .. r0 + n_rows) (their `seg` fields index segments / wav / tile_max of the whole plan)
int fx_launch(const adtfe_plan* plan, int r0, int n_rows, float* wav, float* tile_max, int max_per_seg, cudaStream_t st) {
    if (n_rows <= 0) return ADTFE_OK;
    const int sr = plan->sample_rate;
    ADTFE_REQUIRE(sr >= 8000 && sr <= 96000, ADTFE_ERR_UNSUPPORTED,
                  "adtfe_render: FX need plan->sample_rate between 8000 and 96000 (got %d)", sr);
    const ReverbGeom g = reverb_geom(sr);
    ADTFE_REQUIRE(g.ap_size[3] >= 32 && (size_t)g.total * 4 <= 96 * 1024, ADTFE_ERR_UNSUPPORTED,
                  "adtfe_render: reverb delay lines of %d floats at %d Hz do not fit", g.total, sr);
    trace_open("fx_reverb", r0, st);
    fx_reverb_kernel<<<n_rows, 32, (size_t)g.total * 4, st>>>(plan->segments_dev, plan->fx_dev + r0, wav, plan->ld_wav, g);
    trace_close(st);
    ADTFE_CUDA(cudaGetLastError());
    trace_open("fx_dynamics", r0, st);
    fx_dynamics_kernel<<<(n_rows + 31) / 32, 32, 0, st>>>(plan->segments_dev, plan->fx_dev + r0, n_rows, wav, plan->ld_wav,
                                                         tile_max, max_per_seg, sr);
    trace_close(st);
    ADTFE_CUDA(cudaGetLastError());
    return ADTFE_OK;
}

}  // namespace adtfe
