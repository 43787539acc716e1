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
//  mel phase - lanes are the 32 frames, each warp takes the filters the host scheduled for it
//   (longest-processing-time balance): acc += w * P[bin][lane] is conflict-free and the
//   weight is warp-uniform, read from the kernel-parameter constant bank; then
//   log / clamp / affine into a staging tile, and a coalesced copy-out.
#include "common.cuh"
#include "fft_gen.cuh"

namespace adtfe {

constexpr int kWarps = 16;
constexpr int kThreads = kWarps * 32;
constexpr int kRound = 32;                  // frames per CTA round (lanes of the mel phase)
constexpr int kXPitch = 36;                 // floats per exchange row
constexpr int kXFloats = 32 * kXPitch;      // 1152 floats per warp
constexpr int kPPitch = 33;
constexpr int kPFloats = 1025 * kPPitch;
constexpr int kOPitch = 129;                // staging tile pitch (aliases the exchange tiles)
constexpr int kMaxNnz = 4096;
constexpr int kMaxMels = 128;

__device__ __forceinline__ int p_index(int bin) { return bin * kPPitch; }

// Filterbank and its per-warp schedule, passed by value: lives in the kernel-parameter
// constant bank, so the mel phase reads weights with uniform constant loads.
struct MelTables {
    float w[kMaxNnz];
    int16_t ptr[kMaxMels + 1];   // weights of filter m: w[ptr[m] .. ptr[m+1])
    int16_t lo[kMaxMels];        // first bin of filter m
    int16_t sched_ptr[kWarps + 1];
    uint8_t sched[kMaxMels];     // filters of warp w: sched[sched_ptr[w] .. sched_ptr[w+1])
};

struct LogmelArgs {
    const float* wav;
    float* out;
    const float* window;
    const float2* twiddle;
    const float2* lane_tw;
    int64_t ld_wav;
    int32_t n_seg, first, count, hop, n_mels;
};

__global__ void __launch_bounds__(kThreads, 1) logmel_kernel(const LogmelArgs p, const __grid_constant__ MelTables tab) {
    extern __shared__ __align__(16) float smem[];
    float* s_win = smem;                                           // 2048
    float2* s_tw = reinterpret_cast<float2*>(s_win + 2048);        // 32*32
    float2* s_ltw = s_tw + 32 * 32;                                // 3*32
    float* s_x = reinterpret_cast<float*>(s_ltw + 3 * 32);         // kWarps * kXFloats (also the staging tile)
    float* s_p = s_x + kWarps * kXFloats;                          // kPFloats

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < 2048; i += kThreads) s_win[i] = p.window[i];
    for (int i = tid; i < 32 * 32; i += kThreads) s_tw[i] = p.twiddle[i];
    for (int i = tid; i < 3 * 32; i += kThreads) s_ltw[i] = p.lane_tw[i];
    __syncthreads();

    float* xw = s_x + warp * kXFloats;
    const int col32_bin = 32 + 64 * (int)(__brev((unsigned)lane) >> 27);

    const long long total = (long long)p.n_seg * p.count;
    const long long n_rounds = (total + kRound - 1) / kRound;
    for (long long round = blockIdx.x; round < n_rounds; round += gridDim.x) {
        const long long g0 = round * kRound;
        // ================= FFT phase: frames g0 + warp and g0 + warp + 16 =================
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
            const int f = warp + half * kWarps;
            const long long g = g0 + f;
            if (g >= total) break;  // warp-uniform
            const int seg = (int)(g / p.count);
            const int j = (int)(g - (long long)seg * p.count);
            const float* x = p.wav + (long long)seg * p.ld_wav + (long long)(p.first + j) * p.hop - 1024 + lane;

            // ---- pass 1: window + 64-point real DFT over n1 (stride-32 samples)
            float yr[33], yi[33];
            {
                float v[64];
#pragma unroll
                for (int n1 = 0; n1 < 64; ++n1) v[n1] = __ldg(x + 32 * n1) * s_win[32 * n1 + lane];
                rdft64(v, yr, yi);
            }
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
            cdft32(zr, zi);

            // ---- power spectrum into P[bin][f]
            float* pf = s_p + f;
#pragma unroll
            for (int k2 = 0; k2 < 16; ++k2) pf[p_index(lane + 64 * k2)] = zr[k2] * zr[k2] + zi[k2] * zi[k2];
#pragma unroll
            for (int k2 = 16; k2 < 32; ++k2)
                pf[p_index(64 - lane + 64 * (31 - k2))] = zr[k2] * zr[k2] + zi[k2] * zi[k2];
            if ((lane & 1) == 0) pf[p_index(col32_bin)] = cr * cr + ci * ci;
        }
        __syncthreads();

        // ================= mel phase: lane = frame, filters from the warp's schedule =================
        {
            const float* pl = s_p + lane;
            float* stage = s_x + lane * kOPitch;
            const int s0 = tab.sched_ptr[warp], s1 = tab.sched_ptr[warp + 1];
            for (int si = s0; si < s1; ++si) {
                const int m = tab.sched[si];
                const int b = tab.ptr[m], e = tab.ptr[m + 1];
                const float* pp = pl + p_index(tab.lo[m]);
                float acc = 0.0f;
                for (int i = b; i < e; ++i, pp += kPPitch) acc = fmaf(tab.w[i], *pp, acc);
                float v = logf(acc + 1e-10f);
                v = v != v ? v : fminf(fmaxf(v, -23.0f), 12.0f);  // torch.clamp keeps NaN
                stage[m] = (v + 23.0f) / 35.0f;
            }
        }
        __syncthreads();

        // ================= copy-out: one warp per frame row, coalesced =================
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
            const int f = warp + half * kWarps;
            const long long g = g0 + f;
            if (g >= total) break;
            float* out = p.out + g * p.n_mels;
            for (int m = lane; m < p.n_mels; m += 32) out[m] = s_x[f * kOPitch + m];
        }
        __syncthreads();  // staging tile and P are reused by the next round
    }
}

}  // namespace adtfe

using namespace adtfe;

static size_t logmel_smem_bytes() {
    return ((size_t)2048 + 2 * 32 * 32 + 2 * 3 * 32 + (size_t)kWarps * kXFloats + kPFloats) * 4;
}
static_assert(kWarps * kXFloats >= kRound * kOPitch, "staging tile must fit in the exchange tiles");

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
    a.ld_wav = ld_wav; a.n_seg = n_seg; a.first = first; a.count = count; a.hop = mel->hop; a.n_mels = mel->n_mels;
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
    cudaFree(mel->window); cudaFree(mel->twiddle); cudaFree(mel->lane_tw);
    delete mel->tables;
    delete mel;
    return ADTFE_OK;
}

#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

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

    ADTFE_REQUIRE((int)w.size() <= kMaxNnz, ADTFE_ERR_UNSUPPORTED,
                  "adtfe_mel_create: filterbank has %zu weights between first and last non-zero bins (max %d)",
                  w.size(), kMaxNnz);
    adtfe_mel* mel = new adtfe_mel();
    mel->device = device; mel->n_fft = n_fft; mel->hop = hop; mel->n_mels = n_mels;
    mel->wpi = (n_fft / 2) / hop + 1;  // int((win/2)//hop + 1), model.py:79
    mel->nnz = (int32_t)w.size();
    mel->sm_count = device_sm_count(device);
    mel->smem_bytes = logmel_smem_bytes();
    mel->tables = new adtfe_mel_tables();
    {
        MelTables& t = mel->tables->t;
        memset(&t, 0, sizeof(t));
        for (size_t i = 0; i < w.size(); ++i) t.w[i] = w[i];
        for (int m = 0; m <= n_mels; ++m) t.ptr[m] = (int16_t)ptr[m];
        for (int m = 0; m < n_mels; ++m) t.lo[m] = (int16_t)lo[m];
        // longest-processing-time schedule of filters onto the 16 warps of the mel phase
        std::vector<int> order(n_mels);
        for (int m = 0; m < n_mels; ++m) order[m] = m;
        std::stable_sort(order.begin(), order.end(),
                         [&](int a, int b) { return ptr[a + 1] - ptr[a] > ptr[b + 1] - ptr[b]; });
        std::vector<std::vector<int>> bins(kWarps);
        std::vector<long> load(kWarps, 0);
        for (int m : order) {
            int best = 0;
            for (int wv = 1; wv < kWarps; ++wv)
                if (load[wv] < load[best]) best = wv;
            bins[best].push_back(m);
            load[best] += (ptr[m + 1] - ptr[m]) + 12;  // + the log / store epilogue
        }
        int pos = 0;
        for (int wv = 0; wv < kWarps; ++wv) {
            t.sched_ptr[wv] = (int16_t)pos;
            for (int m : bins[wv]) t.sched[pos++] = (uint8_t)m;
        }
        t.sched_ptr[kWarps] = (int16_t)pos;
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
