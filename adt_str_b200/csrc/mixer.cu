// Overlap-add tile mixer (sm_100a): peaks -> tiles -> normalise.
//
// Replaces the audio arithmetic of SynthDrum.__call__ (reference modules/synthetiser.py):
//   drum_rendering 214-239:  o = (1-mixup)*a + mixup*b ; o /= max|o| ; o *= vol ;
//                            track[start : start+len] += o[:len]
//   instrument_mixer 149-156: wav = sum_i track_i * gain_i ; wav / max|wav| * max_volume
// and the zero padding of collate_fn (data_modules/train_dataset.py:53).
//
// The two data-dependent maxima make it two short kernels:
//   1. peak_kernel      one CTA per (segment, instrument, 4096-sample chunk): both one-shots are
//                       read once and max|ca*a + cb*b| is taken for every note of that
//                       instrument at the same time (each note has its own mixup); chunks meet
//                       in an atomicMax on the float bits (order-independent); chunk 0 also
//                       resolves the bank lookups into ResolvedEvent records.
//   2. mix_kernel       one CTA per 2048-sample output tile; events come from the host-built
//                       CSR (tile -> events) and are added in array order into registers,
//                       so the result is deterministic and needs no atomics; emits the
//                       tile's |max|.
//                       The last CTA of a segment to finish (atomic ticket) takes the max of the
//                       tile maxima and normalises the row in place, wav / peak * max_volume
//                       (an all-zero mix gives NaN, like the reference's 0/0).
#include "common.cuh"

namespace adtfe {

constexpr int kPeakThreads = 256;
constexpr int kPeakChunk = 8;   // notes of one instrument handled per sweep over the one-shots
constexpr int kPeakSpan = ADTFE_PEAK_SPAN;  // samples of the mixed one-shot per peak work item
constexpr int kPeakIters = kPeakSpan / 4 / kPeakThreads;  // float4 per thread per one-shot
static_assert(kPeakIters * kPeakThreads * 4 == kPeakSpan, "peak span must be a multiple of 4 * threads");
constexpr int kMixThreads = 256;
constexpr int kPerThread = ADTFE_TILE / kMixThreads;  // 8 samples, stride kMixThreads
constexpr int kStage = 64;      // resolved events staged in shared memory per round

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

__device__ __forceinline__ float nan_max(float a, float b) {  // torch.max propagates NaN
    return (a != a || b != b) ? __int_as_float(0x7fc00000) : fmaxf(a, b);
}

// max|ca*a + cb*b| over the float4s held in registers, for NC notes at once
template <int NC>
__device__ __forceinline__ void peak_chunk(const float4 (&va)[kPeakIters], const float4 (&vb)[kPeakIters],
                                           const float* __restrict__ s_ca, const float* __restrict__ s_cb,
                                           float (*s_red)[kPeakThreads / 32], int tid) {
    float ca[NC], cb[NC], m[NC];
#pragma unroll
    for (int i = 0; i < NC; ++i) { ca[i] = s_ca[i]; cb[i] = s_cb[i]; m[i] = 0.0f; }
#pragma unroll
    for (int it = 0; it < kPeakIters; ++it) {
#pragma unroll
        for (int i = 0; i < NC; ++i) {
            // separate roundings, as torch's mul, mul, add (synthetiser.py:223)
            m[i] = fmaxf(m[i], fabsf(__fadd_rn(__fmul_rn(va[it].x, ca[i]), __fmul_rn(cb[i], vb[it].x))));
            m[i] = fmaxf(m[i], fabsf(__fadd_rn(__fmul_rn(va[it].y, ca[i]), __fmul_rn(cb[i], vb[it].y))));
            m[i] = fmaxf(m[i], fabsf(__fadd_rn(__fmul_rn(va[it].z, ca[i]), __fmul_rn(cb[i], vb[it].z))));
            m[i] = fmaxf(m[i], fabsf(__fadd_rn(__fmul_rn(va[it].w, ca[i]), __fmul_rn(cb[i], vb[it].w))));
        }
    }
#pragma unroll
    for (int i = 0; i < NC; ++i) {
        const float w = warp_max(m[i]);
        if ((tid & 31) == 0) s_red[i][tid >> 5] = w;
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
    __shared__ float s_red[kPeakChunk][kPeakThreads / 32];
    __shared__ float s_ca[kPeakChunk], s_cb[kPeakChunk];
    const adtfe_peak_item item = work[blockIdx.x];  // one fetch, then the data loads can start
    const int chunk = item.chunk, tid = threadIdx.x;
    const int e0 = item.first_event, e1 = e0 + item.n_events;
    if (e0 >= e1) return;
    const int64_t a_off = item.a_off, b_off = item.b_off;
    const int la = item.la, lb = item.lb, n = item.mix_len;
    // every one-shot is padded to 4 floats, so whole float4s up to the padded length are readable
    const int la4 = (la + 3) >> 2, lb4 = (lb + 3) >> 2;
    const int lo4 = chunk * (kPeakSpan / 4), hi4 = min((n + 3) >> 2, lo4 + kPeakSpan / 4);
    const float4* a4 = reinterpret_cast<const float4*>(pcm + a_off);
    const float4* b4 = reinterpret_cast<const float4*>(pcm + b_off);

    float4 va[kPeakIters], vb[kPeakIters];
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int it = 0; it < kPeakIters; ++it) {
        const int i4 = lo4 + tid + it * kPeakThreads;
        va[it] = (i4 < hi4 && i4 < la4) ? __ldg(a4 + i4) : z;
        vb[it] = (i4 < hi4 && i4 < lb4) ? __ldg(b4 + i4) : z;
    }
    if (chunk == 0) {  // resolve the bank lookups once per note for the tile mixer
        for (int e = e0 + tid; e < e1; e += kPeakThreads) {
            const adtfe_event ev = events[e];
            ResolvedEvent r;
            r.a_off = a_off; r.b_off = b_off;
            r.la = min(la, ev.len); r.lb = min(lb, ev.len);
            r.start = ev.start; r.len = ev.len; r.ca = ev.ca; r.cb = ev.cb; r.gain = ev.gain; r.pad = 0;
            resolved[e] = r;
        }
    }
#pragma unroll
    for (int it = 0; it < kPeakIters; ++it) {  // last float4 of a one-shot: ignore whatever pads it
        const int i4 = lo4 + tid + it * kPeakThreads;
        if (4 * i4 + 3 >= la) {
            if (4 * i4 + 1 >= la) va[it].y = 0.f;
            if (4 * i4 + 2 >= la) va[it].z = 0.f;
            va[it].w = 0.f;
        }
        if (4 * i4 + 3 >= lb) {
            if (4 * i4 + 1 >= lb) vb[it].y = 0.f;
            if (4 * i4 + 2 >= lb) vb[it].z = 0.f;
            vb[it].w = 0.f;
        }
    }
    for (int c0 = e0; c0 < e1; c0 += kPeakChunk) {
        const int nc = min(kPeakChunk, e1 - c0);
        __syncthreads();
        if (tid < kPeakChunk) {
            const bool live = tid < nc;
            s_ca[tid] = live ? events[c0 + tid].ca : 0.0f;
            s_cb[tid] = live ? events[c0 + tid].cb : 0.0f;
        }
        __syncthreads();
        if (nc <= 1) peak_chunk<1>(va, vb, s_ca, s_cb, s_red, tid);
        else if (nc <= 2) peak_chunk<2>(va, vb, s_ca, s_cb, s_red, tid);
        else if (nc <= 4) peak_chunk<4>(va, vb, s_ca, s_cb, s_red, tid);
        else peak_chunk<8>(va, vb, s_ca, s_cb, s_red, tid);
        __syncthreads();
        if (tid < nc) {
            float peak = 0.0f;
#pragma unroll
            for (int w = 0; w < kPeakThreads / 32; ++w) peak = fmaxf(peak, s_red[tid][w]);
            if (peak > 0.0f) atomicMax(peak_bits + c0 + tid, __float_as_int(peak));
        }
    }
}

// One source of one note inside one tile: acc[n] += coef * src[n - start] for start <= n < start+len.
struct SubEvent {
    const float* src;
    int32_t start, len;
    float coef;
    int32_t pad;
};

__global__ void __launch_bounds__(kMixThreads) mix_kernel(
    const float* __restrict__ pcm, const ResolvedEvent* __restrict__ resolved, const int* __restrict__ peak_bits,
    const int32_t* __restrict__ tile_ptr, const int32_t* __restrict__ tile_events,
    const adtfe_segment* __restrict__ segments, int tiles_per_seg, int64_t ld_wav, float* __restrict__ wav,
    float* __restrict__ tile_max, int* __restrict__ seg_done) {
    __shared__ __align__(16) SubEvent s_sub[2 * kStage];
    __shared__ float s_red[kMixThreads / 32];
    __shared__ int s_last;
    const int tile_id = blockIdx.x, tid = threadIdx.x;
    const int seg = tile_id / tiles_per_seg, tile = tile_id - seg * tiles_per_seg;
    const int lo = tile * ADTFE_TILE;
    float acc[kPerThread];
#pragma unroll
    for (int j = 0; j < kPerThread; ++j) acc[j] = 0.0f;

    const int p0 = tile_ptr[tile_id], p1 = tile_ptr[tile_id + 1];
    for (int base = p0; base < p1; base += kStage) {
        const int nst = min(kStage, p1 - base);
        __syncthreads();
        // (1 - mixup) * a and mixup * b become two independent sources, each scaled by
        // gain / peak: o = (ca*a + cb*b) / peak * vol * gain up to float rounding (1e-7 relative)
        if (tid < nst) {
            const int e = tile_events[base + tid];
            const ResolvedEvent ev = resolved[e];
            // an all-zero one-shot has peak 0: gain/0 = inf (or NaN), and 0 * inf = NaN over the whole
            // note - the reference's o / 0 (synthetiser.py:225)
            const float scale = ev.gain / __int_as_float(peak_bits[e]);
            SubEvent a, b;
            a.src = pcm + ev.a_off; a.start = ev.start; a.len = ev.la; a.coef = ev.ca * scale; a.pad = 0;
            b.src = pcm + ev.b_off; b.start = ev.start; b.len = ev.lb; b.coef = ev.cb * scale; b.pad = 0;
            s_sub[2 * tid] = a;
            s_sub[2 * tid + 1] = b;
        }
        __syncthreads();
        // the two sources of one note are processed together: 16 loads in flight per thread
        for (int k = 0; k < 2 * nst; k += 2) {
            const SubEvent ea = s_sub[k], eb = s_sub[k + 1];
            const int r0 = lo + tid - ea.start;  // same start for both
            const float* sa = ea.src + r0;
            const float* sb = eb.src + r0;
            float va[kPerThread], vb[kPerThread];
#pragma unroll
            for (int j = 0; j < kPerThread; ++j) {
                const unsigned r = (unsigned)(r0 + j * kMixThreads);
                va[j] = r < (unsigned)ea.len ? __ldg(sa + j * kMixThreads) : 0.0f;
                vb[j] = r < (unsigned)eb.len ? __ldg(sb + j * kMixThreads) : 0.0f;
            }
#pragma unroll
            for (int j = 0; j < kPerThread; ++j) {
                const unsigned r = (unsigned)(r0 + j * kMixThreads);
                // samples outside the note stay untouched even when coef is inf / NaN
                if (r < (unsigned)ea.len) acc[j] = fmaf(va[j], ea.coef, acc[j]);
                if (r < (unsigned)eb.len) acc[j] = fmaf(vb[j], eb.coef, acc[j]);
            }
        }
    }
    float m = 0.0f;
    float* row = wav + (int64_t)seg * ld_wav;
#pragma unroll
    for (int j = 0; j < kPerThread; ++j) {
        const int n = lo + tid + j * kMixThreads;
        if (n < ld_wav) row[n] = acc[j];
        m = nan_max(m, fabsf(acc[j]));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = nan_max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((tid & 31) == 0) s_red[tid >> 5] = m;
    __syncthreads();
    if (tid == 0) {
        float r = s_red[0];
        for (int i = 1; i < kMixThreads / 32; ++i) r = nan_max(r, s_red[i]);
        tile_max[tile_id] = r;
    }
    // ---- the last CTA of a segment to get here normalises the whole row (wav / peak * max_volume,
    // synthetiser.py:142-144,156).  Release: tile + tile_max written, fence, then the ticket;
    // acquire: ticket, fence, then read through L2 (__ldcg) what the other CTAs wrote.
    const adtfe_segment sg = segments[seg];
    if (sg.flags == 0) return;  // empty note list: the row is already all zeros
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = atomicAdd(seg_done + seg, 1) == tiles_per_seg - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    float peak = 0.0f;
    for (int t = tid; t < tiles_per_seg; t += kMixThreads) peak = nan_max(peak, __ldcg(tile_max + seg * tiles_per_seg + t));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) peak = nan_max(peak, __shfl_xor_sync(0xffffffffu, peak, o));
    if ((tid & 31) == 0) s_red[tid >> 5] = peak;
    __syncthreads();
    peak = s_red[0];
    for (int i = 1; i < kMixThreads / 32; ++i) peak = nan_max(peak, s_red[i]);
    const float vol = sg.max_volume;
    float4* row4 = reinterpret_cast<float4*>(row);  // ld_wav is a multiple of 4 and the base is 16-byte aligned
    const int n4 = sg.len >> 2;
    for (int i = tid; i < n4; i += kMixThreads) {
        float4 v = __ldcg(row4 + i);
        v.x = __fmul_rn(__fdiv_rn(v.x, peak), vol); v.y = __fmul_rn(__fdiv_rn(v.y, peak), vol);
        v.z = __fmul_rn(__fdiv_rn(v.z, peak), vol); v.w = __fmul_rn(__fdiv_rn(v.w, peak), vol);
        row4[i] = v;
    }
    // the reference's row ends at len; beyond it collate_fn pads with exact zeros
    for (int i = 4 * n4 + tid; i < sg.len; i += kMixThreads) row[i] = __fmul_rn(__fdiv_rn(__ldcg(row + i), peak), vol);
}

}  // namespace adtfe

using namespace adtfe;

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

extern "C" size_t adtfe_render_workspace_bytes(int32_t n_events, int32_t n_seg, int32_t tiles_per_seg) {
    if (n_events < 0 || n_seg < 0 || tiles_per_seg < 0) return 0;
    return align256((size_t)n_events * sizeof(ResolvedEvent)) + align256((size_t)n_events * 4 + (size_t)n_seg * 4) +
           align256((size_t)n_seg * tiles_per_seg * 4) + 256;
}

extern "C" int adtfe_render(const adtfe_bank* bank, const adtfe_plan* plan, float* wav_out_dev, void* workspace_dev,
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
    cudaStream_t st = (cudaStream_t)stream;
    char* ws = (char*)(((uintptr_t)workspace_dev + 255) & ~(uintptr_t)255);
    ResolvedEvent* resolved = (ResolvedEvent*)ws;
    int* peak_bits = (int*)(ws + align256((size_t)plan->n_events * sizeof(ResolvedEvent)));
    int* seg_done = peak_bits + plan->n_events;  // zeroed together with the peaks
    float* tile_max = (float*)((char*)peak_bits + align256((size_t)plan->n_events * 4 + (size_t)plan->n_seg * 4));
    const int n_tiles = plan->n_seg * plan->tiles_per_seg;
    ADTFE_CUDA(cudaMemsetAsync(peak_bits, 0, ((size_t)plan->n_events + plan->n_seg) * 4, st));
    if (plan->n_events > 0) {
        peak_kernel<<<plan->n_peak_work, kPeakThreads, 0, st>>>(bank->pcm, plan->events_dev, plan->peak_work_dev,
                                                               resolved, peak_bits);
        ADTFE_CUDA(cudaGetLastError());
    }
    mix_kernel<<<n_tiles, kMixThreads, 0, st>>>(bank->pcm, resolved, peak_bits, plan->tile_ptr_dev,
                                               plan->tile_events_dev, plan->segments_dev, plan->tiles_per_seg,
                                               plan->ld_wav, wav_out_dev, tile_max, seg_done);
    ADTFE_CUDA(cudaGetLastError());
    return ADTFE_OK;
}
