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
// fx_dynamics_kernel  lane = (row, stage): 8 rows of the plan's FX list per warp, 4 lanes each.  An envelope follower
//                     y = |x| + (|x| > y ? cAT : cRL)(y - |x|) has a data-dependent coefficient, so it cannot be
//                     scanned: a row is walked sample by sample.  The three followers of the chain (compressor, limiter
//                     stage 1, limiter stage 2) and the output step sit in four neighbouring lanes, SKEWED by a block of
//                     four samples each - in a trip stage 0 works on block b, stage 1 on block b-1, ... - and pass their
//                     blocks on with shuffles, so the sequential cost of a sample is one follower (5 dependent
//                     instructions; its pow() is exp2(e * log2(x)) on the special-function unit), not three followers
//                     and two pow() in series.  Disabled stages are the same code with an infinite threshold (gain 1),
//                     so rows with different chains share a warp.  The output lane also tracks max|out| of its row: the
//                     row peak the normalisation needs.
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
    // the row's samples are fetched four steps (512 bytes) ahead: a step is ~600 clocks of dependent work, a load from
    // HBM at the top of the step would be waited for in full
    float xq[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) xq[a] = 32 * a + lane < n ? row[32 * a + lane] : 0.0f;
    for (int n0 = 0; n0 < n; n0 += 32) {
        const bool valid = n0 + lane < n;
        const float x = xq[0];
        xq[0] = xq[1]; xq[1] = xq[2]; xq[2] = xq[3];
        xq[3] = n0 + 128 + lane < n ? row[n0 + 128 + lane] : 0.0f;
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

// Four lanes per row: lane 4 j + s is stage s of the warp's row j - 0: compressor, 1 and 2: the limiter's two stages,
// 3: output (gain, clip, row peak, store).  A trip is a block of four samples; the stages are skewed by a trip, so
// that in trip t stage s works on block t - s and hands its four outputs to stage s + 1 with one shuffle each.  Every
// lane runs ONE follower, all of a row's followers advance in the same instructions, and the sequential cost of a
// sample is one follower (with its two special-function instructions) instead of three.  The arithmetic per stage and
// sample is unchanged.
//
// The rows move through shared memory in chunks of kDynChunk samples: the whole warp fetches the next chunk of its
// eight rows with coalesced 512-byte row reads a chunk ahead (the loads stay in registers for the 32 trips of the
// current chunk - never waited for), stage 0 takes its block from shared memory, the output lane parks its block
// there, and a finished chunk leaves with coalesced 512-byte row writes.  The loop body is four trips long: a single
// warp per scheduler runs as fast as its instructions arrive, and the register ring this replaces needed 16-32 unrolled
// trips (30-60 KB of code: 22 % of the stalls were instruction fetch, 21 % the loads of a row that eight lanes read
// 16 bytes at a time).
constexpr int kDynChunk = 128;                       // samples of a row per chunk: 32 trips
constexpr int kDynTrips = kDynChunk / 4;
constexpr int kDynPitch = kDynChunk + 4;             // floats between the rows of a buffer: 528 B, so that the eight rows'
                                                     // 16-byte blocks of one trip lie in different banks
constexpr int kDynBuf = 8 * kDynPitch;               // floats per buffer (eight rows)
__host__ __device__ constexpr size_t dyn_smem_bytes() { return (size_t)4 * kDynBuf * 4 + 32 * 16; }   // in[2] | out[2] | scratch

__global__ void __launch_bounds__(32) fx_dynamics_kernel(const adtfe_segment* __restrict__ segments,
                                                         const adtfe_fx* __restrict__ fx, int n_fx, float* __restrict__ wav,
                                                         int64_t ld_wav, float* __restrict__ tile_max, int max_per_seg,
                                                         int sample_rate) {
    extern __shared__ __align__(16) float s_dyn[];
    float* s_in = s_dyn;                  // [2][8][kDynPitch]
    float* s_out = s_dyn + 2 * kDynBuf;   // [2][8][kDynPitch]
    const int lane = threadIdx.x, stage = lane & 3, j = lane >> 2;
    const int r = blockIdx.x * 8 + j;
    adtfe_fx f;
    f.flags = 0; f.seg = 0;
    if (r < n_fx) f = fx[r];
    const int seg = f.seg;
    const bool live = r < n_fx && f.flags != 0 && segments[seg].flags != 0;
    if (!__any_sync(0xffffffffu, live)) return;
    const double exp_factor = -2.0 * 3.14159265358979323846 * 1000.0 / (double)sample_rate;
    const bool comp_on = live && (f.flags & ADTFE_FX_COMPRESSOR) != 0, lim_on = live && (f.flags & ADTFE_FX_LIMITER) != 0;
    // this lane's follower; the output lane (and every stage that is off) runs one with an infinite threshold: gain 1
    Follower c = make_follower(false, exp_factor, 0.0f, 1.0f, 1.0f, 1.0f);
    if (stage == 0) c = make_follower(comp_on, exp_factor, f.comp_threshold_db, f.comp_ratio, f.comp_attack_ms, f.comp_release_ms);
    if (stage == 1) c = make_follower(lim_on, exp_factor, -10.0f, 4.0f, 2.0f, 200.0f);
    if (stage == 2) c = make_follower(lim_on, exp_factor, f.lim_threshold_db, 1000.0f, 0.001f, 100.0f);
    float out_gain = 1.0f, clip = __int_as_float(0x7f800000);
    if (lim_on) {
        out_gain = (float)pow(10.0, 10.0 * (1.0 - (double)0.25f) / 40.0);
        out_gain *= f.lim_threshold_db < 100.0f ? powf(10.0f, -f.lim_threshold_db * 0.05f) : 0.0f;
        clip = 1.0f;
    }
    float* row = wav + (int64_t)seg * ld_wav;
    const int n = live ? segments[seg].len : 0;
    const bool first = stage == 0, last = stage == 3, rewrite = lim_on || comp_on;
    const int n_blocks = (n + 3) / 4;
    // the eight rows of the warp, for the chunk copies: lane l moves block l of every row's chunk
    const float4* rows4[8];
    int rows_blocks[8];
    unsigned rewrite_mask = __ballot_sync(0xffffffffu, rewrite && last);   // bit 4 j + 3: row j is written back
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const unsigned long long p = __shfl_sync(0xffffffffu, (unsigned long long)(uintptr_t)row, 4 * k);
        rows4[k] = reinterpret_cast<const float4*>((uintptr_t)p);
        rows_blocks[k] = __shfl_sync(0xffffffffu, n_blocks, 4 * k);
    }
    int max_blocks = n_blocks;
#pragma unroll
    for (int o = 16; o >= 4; o >>= 1) max_blocks = max(max_blocks, __shfl_xor_sync(0xffffffffu, max_blocks, o));
    // chunks to run: the last block leaves the output lane three trips after it entered stage 0, and a chunk is written
    // back at the end of the iteration after its own
    const int n_iter = (max_blocks + 3 + kDynTrips - 1) / kDynTrips + 1;
    const float4 zero4 = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    float4 ld[8];
    auto fetch = [&](int chunk) {   // the chunk's block `lane` of every row, zeros past the row's end
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int blk = chunk * kDynTrips + lane;
            ld[k] = blk < rows_blocks[k] ? rows4[k][blk] : zero4;
        }
    };
    auto park = [&](int chunk) {    // ... into the input buffer of that chunk
        float* dst = s_in + (chunk & 1) * kDynBuf + 4 * lane;
#pragma unroll
        for (int k = 0; k < 8; ++k) *reinterpret_cast<float4*>(dst + k * kDynPitch) = ld[k];
    };
    fetch(0);
    park(0);
    fetch(1);
    __syncwarp();
    // |max| on the float bits: non-negative floats order like unsigned integers and a NaN's bits lie above infinity's,
    // so the unsigned maximum propagates NaN like torch.max.  Blocks outside the row (pipeline fill, trips of a longer
    // row in the same warp) carry zeros through the chain and do not move it.
    unsigned peak_bits = 0u;
    float in[4] = {0.0f, 0.0f, 0.0f, 0.0f};   // what the previous stage handed over in the last trip
    // where a lane parks its output blocks: the output lane in the row's output buffers, the other lanes (whose output
    // step is not used) in a scratch block of their own - the store is then unconditional and the trip has no branch
    float* const scratch = s_dyn + 4 * kDynBuf + 4 * lane;
    for (int it = 0; it < n_iter; ++it) {
        const float* src = s_in + (it & 1) * kDynBuf + j * kDynPitch;
        // block b - 3 goes to the buffer of its chunk: the first three trips of an iteration finish the previous chunk
        float* const cur = last ? s_out + (it & 1) * kDynBuf + j * kDynPitch : scratch;
        float* const prev = last ? s_out + ((it + 1) & 1) * kDynBuf + j * kDynPitch + 4 * (kDynTrips - 3) : scratch;
        float4 x = *reinterpret_cast<const float4*>(src);   // the four lanes of a row: one address
        for (int t4 = 0; t4 < kDynTrips; t4 += 4) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int t = t4 + k;               // the trip: stage s works on block it * kDynTrips + t - s
                in[0] = first ? x.x : in[0];
                in[1] = first ? x.y : in[1];
                in[2] = first ? x.z : in[2];
                in[3] = first ? x.w : in[3];
                // the next trip's block, a trip ahead (the last trip of the chunk reads the row's padding block)
                x = *reinterpret_cast<const float4*>(src + 4 * (t + 1));
                float out[4];
#pragma unroll
                for (int m = 0; m < 4; ++m) out[m] = follower_step(c, in[m]);
                // the output step on what stage 2 handed over (block b - 3); the other lanes' results are not used
                float4 o;
                {
                    float v;
                    v = in[0] * out_gain; o.x = v < -clip ? -clip : (v > clip ? clip : v);   // keeps NaN, like FloatVectorOperations::clip
                    v = in[1] * out_gain; o.y = v < -clip ? -clip : (v > clip ? clip : v);
                    v = in[2] * out_gain; o.z = v < -clip ? -clip : (v > clip ? clip : v);
                    v = in[3] * out_gain; o.w = v < -clip ? -clip : (v > clip ? clip : v);
                }
                peak_bits = max(max(peak_bits, __float_as_uint(fabsf(o.x))), __float_as_uint(fabsf(o.y)));
                peak_bits = max(max(peak_bits, __float_as_uint(fabsf(o.z))), __float_as_uint(fabsf(o.w)));
                float* dst = (t4 == 0 && k < 3) ? prev + (last ? 4 * k : 0) : cur + (last ? 4 * (t - 3) : 0);
                *reinterpret_cast<float4*>(dst) = o;
#pragma unroll
                for (int m = 0; m < 4; ++m) in[m] = __shfl_up_sync(0xffffffffu, out[m], 1, 4);   // stage s -> stage s + 1
            }
        }
        __syncwarp();
        // chunk it - 1 is complete in its output buffer: write the rows that the chain rewrites, whole 512-byte pieces
        if (it >= 1) {
            const float* done = s_out + ((it - 1) & 1) * kDynBuf + 4 * lane;
            const int blk = (it - 1) * kDynTrips + lane;
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (((rewrite_mask >> (4 * k + 3)) & 1u) && blk < rows_blocks[k])
                    const_cast<float4*>(rows4[k])[blk] = *reinterpret_cast<const float4*>(done + k * kDynPitch);
        }
        park(it + 1);     // the chunk fetched during this iteration becomes the next input
        fetch(it + 2);
        __syncwarp();
    }
    // the last block was stored whole: what lies beyond the row's length in it goes back to exact zeros (it carried
    // zeros through the chain, but a row that went NaN would leave NaN there)
    if (live && last && rewrite)
        for (int i = n; i < 4 * n_blocks; ++i) row[i] = 0.0f;
    // the normalisation takes the row peak from the tile maxima: this row's is now the output lane's maximum
    if (live && last) {
        float* tm = tile_max + (size_t)seg * max_per_seg;
        tm[0] = __uint_as_float(peak_bits);
        for (int t = 1; t < max_per_seg; ++t) tm[t] = 0.0f;
    }
}

int fx_prepare_device() {
    // the largest delay-line set this build accepts: 48 kHz (55 KB)
    ADTFE_CUDA(cudaFuncSetAttribute(fx_reverb_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    return ADTFE_OK;
}

// FX of the rows fx_dev[r0 .. r0 + n_rows) (their `seg` fields index segments / wav / tile_max of the whole plan), in
// the chain's order: the reverb first, then compressor and limiter.  The two halves are launched separately because
// they scale differently: the reverb is one warp per row (throughput: rows per wave of warps), the dynamics kernel takes
// the same ~4 ms whatever the number of rows (32 rows per warp, a recursion over the row's samples) - so a chunked render
// runs the reverb chunk by chunk behind the tile mixer and the dynamics ONCE for all the plan's FX rows.
int fx_reverb_launch(const adtfe_plan* plan, int r0, int n_rows, float* wav, cudaStream_t st) {
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
    return ADTFE_OK;
}

int fx_dynamics_launch(const adtfe_plan* plan, int r0, int n_rows, float* wav, float* tile_max, int max_per_seg,
                       cudaStream_t st) {
    if (n_rows <= 0) return ADTFE_OK;
    trace_open("fx_dynamics", r0, st);
    fx_dynamics_kernel<<<(n_rows + 7) / 8, 32, dyn_smem_bytes(), st>>>(plan->segments_dev, plan->fx_dev + r0, n_rows, wav, plan->ld_wav,
                                                         tile_max, max_per_seg, plan->sample_rate);
    trace_close(st);
    ADTFE_CUDA(cudaGetLastError());
    return ADTFE_OK;
}

}  // namespace adtfe
