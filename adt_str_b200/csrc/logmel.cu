// Fused STFT -> |X|^2 -> mel -> log kernel (sm_100a).
//
// Replaces ComputeMelSpectrogram.forward (reference model.py:81-97) and the torchaudio
// MelSpectrogram it calls (model.py:71-78,89): frame (n_fft 2048, centred), periodic Hann,
// real FFT, power, (n_fft/2+1 x n_mels) triangular filterbank, log(x + 1e-10),
// clamp[-23, 12], (x + 23) / 35, keep frames wpi .. T-wpi-2.  Only kept frames are
// computed; their support never reaches the reflect padding.
//
// A CTA (16 warps) works in rounds of 32 frames; nothing but the (32 x n_mels) result tile
// goes to HBM.
//  FFT phase - one warp owns one frame at a time (two per round):
//   pass 1  lane n2 holds x[32*n1 + n2] * hann, n1 = 0..63, and runs a 64-point real DFT
//           over n1 in registers (generated straight-line code, tools/gen_fft.py)
//   twiddle Y[k1][n2] *= W_2048^(k1*n2)
//   exchange through the warp's private shared tile, real parts then imaginary parts
//           (row pitch 36 floats: STS.32 rows and LDS.128 columns are both conflict-free)
//   pass 2  lane k1 (0..31) runs a 32-point complex DFT over n2 -> X[k1 + 64*k2];
//           bins above 1024 are the mirror images of bins 64-k1 + 64*(31-k2)
//   column k1 = 32 (bins 32 + 64*k2) is a 32-point DFT across lanes with shuffles
//   |X|^2 goes to the CTA's power matrix P[bin][frame] (pitch 33: consecutive bins of one
//           frame land in distinct banks; only the 16 column-32 bins collide, once per frame)
//  mel phase - lanes are the 32 frames, each warp owns a contiguous, cost-balanced range of
//   filters: acc += w * P[bin][lane] is conflict-free, weights are broadcast reads (4 per
//   LDS.128) from the warp's (now idle) exchange tile; then log / clamp / affine, staged in
//   the same tile and written out in runs of consecutive filters.
//  The samples of a warp's next frame are requested before the mel phase, so their latency
//  hides behind it.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.cuh"
#include "fft_gen.cuh"

namespace adtfe {

constexpr int kWarps = 16;
constexpr int kThreads = kWarps * 32;
constexpr int kRound = 32;                  // frames per CTA round (lanes of the mel phase)
constexpr int kXPitch = 36;                 // floats per exchange row
constexpr int kXFloats = 32 * kXPitch;      // 1152 floats per warp
constexpr int kPPitch = 33;
constexpr int kPRows = 1025 + 3;               // 3 zero rows: filter weights are padded to groups of 4
constexpr int kPFloats = kPRows * kPPitch;
constexpr int kWtsMax = 384;                // packed filter weights one warp parks in its tile ...
constexpr int kStagePitchMax = 23;          // ... followed by its (32 frames x n filters) result tile
static_assert(kWtsMax + 32 * kStagePitchMax <= kXFloats, "weights + staging must fit in one exchange tile");
constexpr int kMaxMels = 128;

__device__ __forceinline__ int p_index(int bin) { return bin * kPPitch; }

// Per-warp schedule of the mel phase, passed by value (kernel-parameter constant bank).  The
// weights themselves are packed per warp in global memory (adtfe_mel::sched_w) and copied into
// the warp's idle exchange tile every round, where they are read with broadcast LDS.
struct MelEntry {
    int16_t lo, cnt4, woff, pad;  // first bin, weights padded to a multiple of 4, offset in the warp's pack
};
struct MelTables {
    MelEntry entry[kMaxMels];
    int16_t first[kWarps + 1];   // warp w owns the contiguous filters first[w] .. first[w+1]-1
    int16_t wbase[kWarps + 1];   // its packed weights: sched_w[wbase[w] .. wbase[w+1])
};

struct LogmelArgs {
    const float* wav;
    float* out;
    const float* window;
    const float2* twiddle;
    const float2* lane_tw;
    const float* sched_w;
    long long* trace;  // ADTFE_TRACE builds only: per-warp clock64 stamps
    int64_t ld_wav;
    int32_t n_seg, first, count, hop, n_mels;
};

// Mel phase of one warp: lane = frame, the warp owns a contiguous range of filters.
// acc += w * P[bin][lane]: the P read is conflict-free (pitch 33), four weights come from one
// broadcast LDS.128 out of the warp's own tile.  Results are staged in the same tile and
// copied out in runs of n consecutive filters per frame.
__device__ __forceinline__ void mel_phase(const MelTables& tab, int warp, int lane, const float* __restrict__ pl,
                                          float* __restrict__ tile, float* __restrict__ out, long long g0,
                                          long long total, int n_mels) {
    const int f0 = tab.first[warp], n = tab.first[warp + 1] - f0;
    const int pitch = n | 1;
    float* stage = tile + kWtsMax;
    for (int i = 0; i < n; ++i) {
        const MelEntry en = tab.entry[f0 + i];
        const float* pp = pl + p_index(en.lo);
        const float4* ww = reinterpret_cast<const float4*>(tile + en.woff);
        float acc0 = 0.0f, acc1 = 0.0f;
        for (int k = 0; k < en.cnt4; k += 4, pp += 4 * kPPitch) {
            const float4 w = ww[k >> 2];
            acc0 = fmaf(w.x, pp[0], acc0);
            acc1 = fmaf(w.y, pp[kPPitch], acc1);
            acc0 = fmaf(w.z, pp[2 * kPPitch], acc0);
            acc1 = fmaf(w.w, pp[3 * kPPitch], acc1);
        }
        float v = __logf((acc0 + acc1) + 1e-10f);           // |err| ~1e-7 in ln, 35x below the tolerance
        v = v != v ? v : fminf(fmaxf(v, -23.0f), 12.0f);    // torch.clamp keeps NaN
        stage[lane * pitch + i] = (v + 23.0f) * (1.0f / 35.0f);
    }
    __syncwarp();
    for (int idx = lane; idx < 32 * n; idx += 32) {
        const int f = idx / n, i = idx - f * n;
        const long long g = g0 + f;
        if (g < total) out[g * n_mels + f0 + i] = stage[f * pitch + i];
    }
}

__global__ void __launch_bounds__(kThreads, 1) logmel_kernel(const LogmelArgs p, const __grid_constant__ MelTables tab) {
    extern __shared__ __align__(16) float smem[];
    float* s_win = smem;                                           // 2048
    float2* s_tw = reinterpret_cast<float2*>(s_win + 2048);        // 32*32
    float2* s_ltw = s_tw + 32 * 32;                                // 3*32
    float* s_x = reinterpret_cast<float*>(s_ltw + 3 * 32);         // kWarps * kXFloats (also the staging tile)
    float* s_p = s_x + kWarps * kXFloats;                          // kPFloats

    const int tid = threadIdx.x, lane = tid & 31;
    // broadcast from lane 0 so the compiler knows the warp index (and everything the mel phase
    // derives from it: schedule, filter bounds, weight index) is warp-uniform
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    for (int i = tid; i < 2048; i += kThreads) s_win[i] = p.window[i];
    for (int i = tid; i < 32 * 32; i += kThreads) s_tw[i] = p.twiddle[i];
    for (int i = tid; i < 3 * 32; i += kThreads) s_ltw[i] = p.lane_tw[i];
    __syncthreads();

#ifdef ADTFE_TRACE
#define TRACE(slot) do { if (lane == 0 && p.trace && tr_round < 3) p.trace[((blockIdx.x * kWarps + warp) * 3 + tr_round) * 16 + (slot)] = clock64(); } while (0)
    int tr_round = 0;
#else
#define TRACE(slot) do {} while (0)
#endif
    float* xw = s_x + warp * kXFloats;
    const int col32_bin = 32 + 64 * (int)(__brev((unsigned)lane) >> 27);
    const int wb = tab.wbase[warp], wn = tab.wbase[warp + 1] - wb;
    for (int i = tid; i < 3 * kPPitch; i += kThreads) s_p[1025 * kPPitch + i] = 0.0f;  // padding rows

    const long long total = (long long)p.n_seg * p.count;
    const long long n_rounds = (total + kRound - 1) / kRound;

    auto frame_ptr = [&](long long g) -> const float* {
        const int seg = (int)(g / p.count);
        const int j = (int)(g - (long long)seg * p.count);
        return p.wav + (long long)seg * p.ld_wav + (long long)(p.first + j) * p.hop - 1024 + lane;
    };

    // samples of the warp's first frame of the coming round, fetched one round ahead
    float v[64];
    {
        const long long g = (long long)blockIdx.x * kRound + warp;
        if (g < total) {
            const float* x = frame_ptr(g);
#pragma unroll
            for (int n1 = 0; n1 < 64; ++n1) v[n1] = __ldg(x + 32 * n1);
        }
    }

    for (long long round = blockIdx.x; round < n_rounds; round += gridDim.x) {
        const long long g0 = round * kRound;
        TRACE(0);
        // ================= FFT phase: frames g0 + warp and g0 + warp + 16 =================
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
            const int f = warp + half * kWarps;
            const long long g = g0 + f;
            if (g >= total) break;  // warp-uniform
            if (half == 1) {
                const float* x = frame_ptr(g);
#pragma unroll
                for (int n1 = 0; n1 < 64; ++n1) v[n1] = __ldg(x + 32 * n1);
            }
            // ---- pass 1: window + 64-point real DFT over n1 (stride-32 samples)
            float yr[33], yi[33];
#pragma unroll
            for (int n1 = 0; n1 < 64; ++n1) v[n1] *= s_win[32 * n1 + lane];
            TRACE(1 + half * 5);
            rdft64(v, yr, yi);
            TRACE(2 + half * 5);
            // ---- twiddle in place (rows 1..31), column 32 is real before its twiddle
#pragma unroll
            for (int k1 = 1; k1 < 32; ++k1) {
                const float2 w = s_tw[(k1 - 1) * 32 + lane];
                const float a = yr[k1] * w.x - yi[k1] * w.y;
                yi[k1] = yr[k1] * w.y + yi[k1] * w.x;
                yr[k1] = a;
            }
            float cr, ci;
            {
                const float2 w = s_tw[31 * 32 + lane];
                cr = yr[32] * w.x;
                ci = yr[32] * w.y;
            }
            // ---- exchange, real parts
            float zr[32], zi[32];
#pragma unroll
            for (int k1 = 0; k1 < 32; ++k1) xw[k1 * kXPitch + lane] = yr[k1];
            __syncwarp();
            {
                const float4* row = reinterpret_cast<const float4*>(xw + lane * kXPitch);
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const float4 t = row[q];
                    zr[4 * q] = t.x; zr[4 * q + 1] = t.y; zr[4 * q + 2] = t.z; zr[4 * q + 3] = t.w;
                }
            }
            __syncwarp();
            // ---- exchange, imaginary parts (row 0 is purely real)
            xw[lane] = 0.0f;
#pragma unroll
            for (int k1 = 1; k1 < 32; ++k1) xw[k1 * kXPitch + lane] = yi[k1];
            __syncwarp();
            {
                const float4* row = reinterpret_cast<const float4*>(xw + lane * kXPitch);
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const float4 t = row[q];
                    zi[4 * q] = t.x; zi[4 * q + 1] = t.y; zi[4 * q + 2] = t.z; zi[4 * q + 3] = t.w;
                }
            }
            __syncwarp();

            // ---- 32-point DFT across lanes for column 32 (decimation in frequency, bit-reversed)
#pragma unroll
            for (int s = 0; s < 5; ++s) {
                const int hb = 16 >> s;
                const float orr = __shfl_xor_sync(0xffffffffu, cr, hb);
                const float oi = __shfl_xor_sync(0xffffffffu, ci, hb);
                const bool upper = (lane & hb) != 0;
                const float dr = upper ? orr - cr : cr + orr;
                const float di = upper ? oi - ci : ci + oi;
                if (s < 3) {            // W_{2*hb}^(lane mod hb) on the upper half, 1 on the lower
                    const float2 w = s_ltw[s * 32 + lane];
                    cr = dr * w.x - di * w.y;
                    ci = dr * w.y + di * w.x;
                } else if (s == 3) {    // hb = 2: twiddle is 1 or -i
                    const bool rot = upper && (lane & 1);
                    cr = rot ? di : dr;
                    ci = rot ? -dr : di;
                } else {
                    cr = dr;
                    ci = di;
                }
            }

            // ---- pass 2: 32-point complex DFT over n2 for k1 = lane
            TRACE(3 + half * 5);
            cdft32(zr, zi);
            TRACE(4 + half * 5);

            // ---- power spectrum into P[bin][f]
            float* pf = s_p + f;
#pragma unroll
            for (int k2 = 0; k2 < 16; ++k2) pf[p_index(lane + 64 * k2)] = zr[k2] * zr[k2] + zi[k2] * zi[k2];
#pragma unroll
            for (int k2 = 16; k2 < 32; ++k2)
                pf[p_index(64 - lane + 64 * (31 - k2))] = zr[k2] * zr[k2] + zi[k2] * zi[k2];
            if ((lane & 1) == 0) pf[p_index(col32_bin)] = cr * cr + ci * ci;
            TRACE(5 + half * 5);
        }
        // the exchange tile is idle until the next round: park this warp's filter weights in it
        for (int i = lane; i < wn; i += 32) xw[i] = __ldg(p.sched_w + wb + i);
        // and start fetching the first frame of the next round; it lands during the mel phase
        {
            const long long g = (round + gridDim.x) * kRound + warp;
            if (g < total) {
                const float* x = frame_ptr(g);
#pragma unroll
                for (int n1 = 0; n1 < 64; ++n1) v[n1] = __ldg(x + 32 * n1);
            }
        }
        TRACE(11);
        __syncthreads();
        TRACE(12);

        // ================= mel phase: lane = frame, the warp's own filter range =================
        mel_phase(tab, warp, lane, s_p + lane, xw, p.out, g0, total, p.n_mels);
        TRACE(13);
        __syncthreads();  // P and the exchange tiles are reused by the next round
        TRACE(14);
#ifdef ADTFE_TRACE
        ++tr_round;
#endif
    }
}

}  // namespace adtfe

using namespace adtfe;

// ADTFE_TRACE builds: a device buffer address handed over in the environment by tools/trace_logmel.py
static void* getenv_trace() {
#ifdef ADTFE_TRACE
    const char* e = getenv("ADTFE_TRACE_PTR");
    return e ? (void*)strtoull(e, nullptr, 0) : nullptr;
#else
    return nullptr;
#endif
}

static size_t logmel_smem_bytes() {
    return ((size_t)2048 + 2 * 32 * 32 + 2 * 3 * 32 + (size_t)kWarps * kXFloats + kPFloats) * 4;
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

extern "C" int adtfe_logmel(const adtfe_mel* mel, const float* wav_dev, int32_t n_seg, int64_t ld_wav,
                            int64_t n_samples, float* out_dev, void* stream) {
    ADTFE_REQUIRE(mel && n_seg >= 0 && n_samples >= 0 && ld_wav >= n_samples, ADTFE_ERR_BAD_ARG,
                  "adtfe_logmel: bad argument (n_seg=%d ld_wav=%lld n_samples=%lld)", n_seg, (long long)ld_wav,
                  (long long)n_samples);
    int32_t first = 0, count = 0;
    adtfe_mel_frames(mel, n_samples, &first, &count);
    if (n_seg == 0 || count == 0) return ADTFE_OK;
    ADTFE_REQUIRE(wav_dev && out_dev, ADTFE_ERR_BAD_ARG, "adtfe_logmel: null buffer");
    // kept frames stay inside the signal by construction; guard against a caller-made mel
    ADTFE_REQUIRE((int64_t)first * mel->hop >= 1024 &&
                      (int64_t)(first + count - 1) * mel->hop + 1024 <= n_samples,
                  ADTFE_ERR_UNSUPPORTED, "adtfe_logmel: frame support leaves the signal");
    LogmelArgs a;
    a.wav = wav_dev; a.out = out_dev; a.window = mel->window; a.twiddle = mel->twiddle; a.lane_tw = mel->lane_tw;
    a.sched_w = mel->sched_w; a.trace = (long long*)getenv_trace(); a.ld_wav = ld_wav; a.n_seg = n_seg; a.first = first; a.count = count; a.hop = mel->hop; a.n_mels = mel->n_mels;
    const long long total = (long long)n_seg * count;
    const long long n_rounds = (total + kRound - 1) / kRound;
    const int grid = (int)(n_rounds < mel->sm_count ? n_rounds : mel->sm_count);
    logmel_kernel<<<grid, kThreads, mel->smem_bytes, (cudaStream_t)stream>>>(a, mel->tables->t);
    ADTFE_CUDA(cudaGetLastError());
    return ADTFE_OK;
}

extern "C" int adtfe_mel_destroy(adtfe_mel* mel) {
    if (!mel) return ADTFE_OK;
    cudaSetDevice(mel->device);
    cudaFree(mel->window); cudaFree(mel->twiddle); cudaFree(mel->lane_tw); cudaFree(mel->sched_w);
    delete mel->tables;
    delete mel;
    return ADTFE_OK;
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
    std::vector<float> w;
    std::vector<int32_t> ptr(n_mels + 1, 0), lo(n_mels, 0);
    for (int m = 0; m < n_mels; ++m) {
        int first = -1, last = -1;
        for (int k = 0; k < n_bins; ++k)
            if (fb_host[(size_t)k * n_mels + m] != 0.0f) { if (first < 0) first = k; last = k; }
        ptr[m] = (int32_t)w.size();
        lo[m] = first < 0 ? 0 : first;
        if (first >= 0)
            for (int k = first; k <= last; ++k) w.push_back(fb_host[(size_t)k * n_mels + m]);
    }
    ptr[n_mels] = (int32_t)w.size();

    std::vector<float2> tw(32 * 32), ltw(3 * 32);
    const double two_pi = 6.283185307179586476925286766559;
    for (int k1 = 1; k1 <= 32; ++k1)
        for (int n2 = 0; n2 < 32; ++n2) {
            const double a = -two_pi * (double)(k1 * n2) / 2048.0;
            tw[(k1 - 1) * 32 + n2] = make_float2((float)std::cos(a), (float)std::sin(a));
        }
    for (int s = 0; s < 3; ++s) {
        const int half = 16 >> s;
        for (int l = 0; l < 32; ++l) {
            if (l & half) {
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
    mel->nnz = (int32_t)w.size();
    mel->sm_count = device_sm_count(device);
    mel->smem_bytes = logmel_smem_bytes();
    mel->tables = new adtfe_mel_tables();
    std::vector<float> packed;
    {
        MelTables& t = mel->tables->t;
        memset(&t, 0, sizeof(t));
        // contiguous filter ranges for the 16 warps of the mel phase, balanced by estimated cost
        auto cost = [&](int m) { return 1.5 * (ptr[m + 1] - ptr[m]) + 14.0; };
        double total_cost = 0;
        for (int m = 0; m < n_mels; ++m) total_cost += cost(m);
        int m = 0;
        bool fits = true;
        double spent = 0;
        for (int wv = 0; wv < kWarps; ++wv) {
            t.first[wv] = (int16_t)m;
            t.wbase[wv] = (int16_t)packed.size();
            const size_t base = packed.size();
            const double target = total_cost * (wv + 1) / kWarps;
            int taken = 0;
            while (m < n_mels && taken < kStagePitchMax &&
                   (wv == kWarps - 1 || taken == 0 || spent + 0.5 * cost(m) <= target)) {
                MelEntry en;
                const int cnt = ptr[m + 1] - ptr[m];
                en.lo = (int16_t)lo[m]; en.cnt4 = (int16_t)((cnt + 3) & ~3);
                en.woff = (int16_t)(packed.size() - base); en.pad = 0;
                t.entry[m] = en;
                packed.insert(packed.end(), w.begin() + ptr[m], w.begin() + ptr[m + 1]);
                packed.resize(base + en.woff + en.cnt4, 0.0f);
                fits = fits && lo[m] + en.cnt4 <= kPRows;
                spent += cost(m);
                ++m; ++taken;
            }
            fits = fits && packed.size() - base <= (size_t)kWtsMax;
        }
        t.first[kWarps] = (int16_t)m;
        t.wbase[kWarps] = (int16_t)packed.size();
        if (!fits || m != n_mels || packed.size() > 32000) {
            set_error("adtfe_mel_create: filterbank does not fit the mel-phase tiles (%zu stored weights, %d of %d "
                      "filters placed)", packed.size(), m, n_mels);
            adtfe_mel_destroy(mel);
            return ADTFE_ERR_UNSUPPORTED;
        }
    }
    auto fail = [&](int status) { adtfe_mel_destroy(mel); return status; };
#define MEL_UPLOAD(dst, src, bytes)                                                                  \
    if (cudaMalloc((void**)&(dst), (bytes) ? (bytes) : 4) != cudaSuccess ||                          \
        cudaMemcpy((dst), (src), (bytes), cudaMemcpyHostToDevice) != cudaSuccess) {                  \
        set_error("adtfe_mel_create: upload failed: %s", cudaGetErrorString(cudaGetLastError()));    \
        return fail(ADTFE_ERR_CUDA);                                                                 \
    }
    MEL_UPLOAD(mel->window, window_host, (size_t)n_fft * 4);
    MEL_UPLOAD(mel->twiddle, tw.data(), tw.size() * sizeof(float2));
    MEL_UPLOAD(mel->lane_tw, ltw.data(), ltw.size() * sizeof(float2));
    MEL_UPLOAD(mel->sched_w, packed.data(), packed.size() * 4);
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
