// Fused STFT -> |X|^2 -> mel -> log kernel (sm_100a).
//
// Replaces ComputeMelSpectrogram.forward (reference model.py:81-97) and the torchaudio
// MelSpectrogram it calls (model.py:71-78,89): frame (n_fft 2048, centred), periodic Hann,
// real FFT, power, (n_fft/2+1 x n_mels) triangular filterbank, log(x + 1e-10),
// clamp[-23, 12], (x + 23) / 35, keep frames wpi .. T-wpi-2.  Only kept frames are
// computed; their support never reaches the reflect padding.
//
// One warp owns one frame end to end; nothing but the (n_mels) result row goes to HBM.
//   pass 1  lane n2 holds x[32*n1 + n2] * hann, n1 = 0..63, and runs a 64-point real DFT
//           over n1 in registers (generated straight-line code, tools/gen_fft.py)
//   twiddle Y[k1][n2] *= W_2048^(k1*n2)
//   exchange through the warp's private shared-memory tile (row pitch 34 complex: the
//           STS.64 rows and the LDS.128 columns are both conflict-free)
//   pass 2  lane k1 (0..31) runs a 32-point complex DFT over n2 -> X[k1 + 64*k2];
//           bins above 1024 are the mirror images of bins 64-k1 + 64*(31-k2)
//   column k1 = 32 (bins 32 + 64*k2) is a 32-point DFT across lanes with shuffles
//   |X|^2 goes back to the same shared tile in natural bin order, then every lane
//   accumulates filters lane, lane+32, ... from the CSR filterbank and writes log-mel.
#include "common.cuh"
#include "fft_gen.cuh"

namespace adtfe {

constexpr int kWarps = 16;                 // frames in flight per CTA
constexpr int kThreads = kWarps * 32;
constexpr int kRow = 34;                   // complex per exchange row (32 + 2 pad)
constexpr int kWarpFloats = 32 * kRow * 2; // 2176 floats = 8704 B, also holds 1025 powers

struct LogmelArgs {
    const float* wav;
    float* out;
    const float* window;
    const float2* twiddle;
    const float2* lane_tw;
    const float* fb_w;
    const int32_t* fb_ptr;
    const int32_t* fb_lo;
    int64_t ld_wav;
    int32_t n_seg, first, count, hop, n_mels, nnz;
};

__global__ void __launch_bounds__(kThreads, 1) logmel_kernel(const LogmelArgs p) {
    extern __shared__ __align__(16) float smem[];
    float* s_win = smem;                                           // 2048
    float2* s_tw = reinterpret_cast<float2*>(s_win + 2048);        // 32*32
    float2* s_ltw = s_tw + 32 * 32;                                // 3*32
    float* s_y = reinterpret_cast<float*>(s_ltw + 3 * 32);         // kWarps * kWarpFloats
    float* s_fbw = s_y + kWarps * kWarpFloats;                     // nnz (padded to 4)
    int32_t* s_ptr = reinterpret_cast<int32_t*>(s_fbw + ((p.nnz + 3) & ~3));  // n_mels+1
    int32_t* s_lo = s_ptr + p.n_mels + 1;                          // n_mels

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < 2048; i += kThreads) s_win[i] = p.window[i];
    for (int i = tid; i < 32 * 32; i += kThreads) s_tw[i] = p.twiddle[i];
    for (int i = tid; i < 3 * 32; i += kThreads) s_ltw[i] = p.lane_tw[i];
    for (int i = tid; i < p.nnz; i += kThreads) s_fbw[i] = p.fb_w[i];
    for (int i = tid; i <= p.n_mels; i += kThreads) s_ptr[i] = p.fb_ptr[i];
    for (int i = tid; i < p.n_mels; i += kThreads) s_lo[i] = p.fb_lo[i];
    __syncthreads();

    float2* y = reinterpret_cast<float2*>(s_y + warp * kWarpFloats);
    float* pw = reinterpret_cast<float*>(y);
    const int col32_bin = 32 + 64 * (int)(__brev((unsigned)lane) >> 27);

    const long long total = (long long)p.n_seg * p.count;
    const long long n_blocks = (total + kWarps - 1) / kWarps;
    for (long long blk = blockIdx.x; blk < n_blocks; blk += gridDim.x) {
        const long long g = blk * kWarps + warp;
        if (g >= total) continue;  // warp-uniform
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
        // ---- twiddle and hand rows 0..31 to the lanes that own them
        y[lane] = make_float2(yr[0], 0.0f);
#pragma unroll
        for (int k1 = 1; k1 < 32; ++k1) {
            const float2 w = s_tw[(k1 - 1) * 32 + lane];
            y[k1 * kRow + lane] = make_float2(yr[k1] * w.x - yi[k1] * w.y, yr[k1] * w.y + yi[k1] * w.x);
        }
        // column k1 = 32 is real before the twiddle
        float cr, ci;
        {
            const float2 w = s_tw[31 * 32 + lane];
            cr = yr[32] * w.x;
            ci = yr[32] * w.y;
        }
        __syncwarp();

        // ---- 32-point DFT across lanes (decimation in frequency, result bit-reversed)
#pragma unroll
        for (int s = 0; s < 5; ++s) {
            const int half = 16 >> s;
            const float orr = __shfl_xor_sync(0xffffffffu, cr, half);
            const float oi = __shfl_xor_sync(0xffffffffu, ci, half);
            const bool upper = (lane & half) != 0;
            const float dr = upper ? orr - cr : cr + orr;
            const float di = upper ? oi - ci : ci + oi;
            if (s < 3) {            // W_{2*half}^(lane mod half) on the upper half, 1 on the lower
                const float2 w = s_ltw[s * 32 + lane];
                cr = dr * w.x - di * w.y;
                ci = dr * w.y + di * w.x;
            } else if (s == 3) {    // half = 2: twiddle is 1 or -i
                const bool rot = upper && (lane & 1);
                cr = rot ? di : dr;
                ci = rot ? -dr : di;
            } else {
                cr = dr;
                ci = di;
            }
        }

        // ---- pass 2: 32-point complex DFT over n2 for k1 = lane
        float zr[32], zi[32];
        {
            const float4* row = reinterpret_cast<const float4*>(y + lane * kRow);
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                const float4 t = row[q];
                zr[2 * q] = t.x; zi[2 * q] = t.y; zr[2 * q + 1] = t.z; zi[2 * q + 1] = t.w;
            }
        }
        cdft32(zr, zi);
        __syncwarp();  // every lane has read its row: the tile can take the powers

        // ---- power spectrum in natural bin order
#pragma unroll
        for (int k2 = 0; k2 < 16; ++k2) pw[lane + 64 * k2] = zr[k2] * zr[k2] + zi[k2] * zi[k2];
#pragma unroll
        for (int k2 = 16; k2 < 32; ++k2)
            pw[64 - lane + 64 * (31 - k2)] = zr[k2] * zr[k2] + zi[k2] * zi[k2];
        if ((lane & 1) == 0) pw[col32_bin] = cr * cr + ci * ci;
        __syncwarp();

        // ---- mel filterbank (CSR by filter) + log / clamp / affine
        float* out = p.out + g * p.n_mels;
        for (int m = lane; m < p.n_mels; m += 32) {
            const int b = s_ptr[m], e = s_ptr[m + 1];
            const float* pp = pw + s_lo[m] - b;
            float acc = 0.0f;
            for (int i = b; i < e; ++i) acc = fmaf(s_fbw[i], pp[i], acc);
            float v = logf(acc + 1e-10f);
            v = fminf(fmaxf(v, -23.0f), 12.0f);
            out[m] = (v + 23.0f) / 35.0f;
        }
        __syncwarp();  // powers consumed before the next frame reuses the tile
    }
}

}  // namespace adtfe

using namespace adtfe;

static size_t logmel_smem_bytes(int nnz, int n_mels) {
    size_t floats = 2048 + 2 * 32 * 32 + 2 * 3 * 32 + (size_t)kWarps * kWarpFloats + ((nnz + 3) & ~3);
    return floats * 4 + (size_t)(2 * n_mels + 1) * 4;
}

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
    a.fb_w = mel->fb_w; a.fb_ptr = mel->fb_ptr; a.fb_lo = mel->fb_lo; a.ld_wav = ld_wav; a.n_seg = n_seg;
    a.first = first; a.count = count; a.hop = mel->hop; a.n_mels = mel->n_mels; a.nnz = mel->nnz;
    const long long total = (long long)n_seg * count;
    const long long n_blocks = (total + kWarps - 1) / kWarps;
    const int grid = (int)(n_blocks < mel->sm_count ? n_blocks : mel->sm_count);
    logmel_kernel<<<grid, kThreads, mel->smem_bytes, (cudaStream_t)stream>>>(a);
    ADTFE_CUDA(cudaGetLastError());
    return ADTFE_OK;
}

extern "C" int adtfe_mel_destroy(adtfe_mel* mel) {
    if (!mel) return ADTFE_OK;
    cudaSetDevice(mel->device);
    cudaFree(mel->window); cudaFree(mel->twiddle); cudaFree(mel->lane_tw);
    cudaFree(mel->fb_w); cudaFree(mel->fb_ptr); cudaFree(mel->fb_lo);
    delete mel;
    return ADTFE_OK;
}

#include <cmath>
#include <vector>

extern "C" int adtfe_mel_create(int32_t n_fft, int32_t hop, int32_t n_mels, const float* window_host,
                                const float* fb_host, int device, adtfe_mel** out) {
    ADTFE_REQUIRE(out && window_host && fb_host, ADTFE_ERR_BAD_ARG, "adtfe_mel_create: null pointer");
    *out = nullptr;
    ADTFE_REQUIRE(n_fft == 2048, ADTFE_ERR_UNSUPPORTED, "adtfe_mel_create: n_fft %d unsupported (only 2048)", n_fft);
    ADTFE_REQUIRE(hop >= 1 && hop <= 2048, ADTFE_ERR_UNSUPPORTED, "adtfe_mel_create: hop %d out of range", hop);
    ADTFE_REQUIRE(n_mels >= 1 && n_mels <= 256, ADTFE_ERR_UNSUPPORTED, "adtfe_mel_create: n_mels %d out of range", n_mels);
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
    mel->smem_bytes = logmel_smem_bytes(mel->nnz, n_mels);
    auto fail = [&](int status) { adtfe_mel_destroy(mel); return status; };
    if (mel->smem_bytes > 227 * 1024) {
        set_error("adtfe_mel_create: filterbank needs %zu B of shared memory", mel->smem_bytes);
        return fail(ADTFE_ERR_UNSUPPORTED);
    }
#define MEL_UPLOAD(dst, src, bytes)                                                                  \
    if (cudaMalloc((void**)&(dst), (bytes) ? (bytes) : 4) != cudaSuccess ||                          \
        cudaMemcpy((dst), (src), (bytes), cudaMemcpyHostToDevice) != cudaSuccess) {                  \
        set_error("adtfe_mel_create: upload failed: %s", cudaGetErrorString(cudaGetLastError()));    \
        return fail(ADTFE_ERR_CUDA);                                                                 \
    }
    MEL_UPLOAD(mel->window, window_host, (size_t)n_fft * 4);
    MEL_UPLOAD(mel->twiddle, tw.data(), tw.size() * sizeof(float2));
    MEL_UPLOAD(mel->lane_tw, ltw.data(), ltw.size() * sizeof(float2));
    MEL_UPLOAD(mel->fb_w, w.data(), w.size() * 4);
    MEL_UPLOAD(mel->fb_ptr, ptr.data(), ptr.size() * 4);
    MEL_UPLOAD(mel->fb_lo, lo.data(), lo.size() * 4);
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
