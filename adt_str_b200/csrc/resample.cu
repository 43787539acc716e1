// Long-form front of eval / inference (sm_100a): polyphase sinc resampling, channel mean, peak normalisation.
//
// Replaces, for audio that is already in device memory:
//   utils/audio_utils.py:17-19  resample()  = torchaudio.transforms.Resample(orig, new)(wav)      (also
//   inference.py:82-84, data_modules/eval_dataset.py:69)
//   utils/audio_utils.py:12     wav.mean(0)   / inference.py:86-87  waveform.mean(dim=0, keepdim=True)
//   utils/audio_utils.py:22-23  normalize()  = wav / wav.abs().max()                               (eval_dataset.py:70)
// The chunking that follows (inference.py:35-48 _chunk_audio) is a view: non-overlapping chunks of a zero-padded
// signal are the rows of a (n_chunks, chunk_samples) matrix, which is what adtfe_logmel takes.
//
// torchaudio's algorithm (functional._apply_sinc_resample_kernel): with o = orig/gcd, n = new/gcd, K = 2*width + o,
//   y[j*n + p] = sum_{k<K} xpad[j*o + k] * kernel[p][k],   xpad = `width` zeros, x, `width + o` zeros,
// truncated to ceil(n * len / o) samples: a strided conv1d.  Here one CTA stages the input span of a tile of
// groups j (and, when it fits, the filter bank) in shared memory; a thread owns a 4 x 4 register tile of groups x
// phases.  Every output is one chain of float32 FMAs over ascending k.
#include <algorithm>
#include <vector>

#include "common.cuh"

struct adtfe_resampler {
    int device = 0;
    int32_t orig = 1, neu = 1;  // frequencies divided by their gcd
    int32_t width = 0, taps = 0;
    float* table = nullptr;     // [taps][neu]: kernel[p][k] transposed
    int32_t groups_per_tile = 1;
    int32_t phase_stride = 1;   // ceil(neu / kResampleP): a thread's phases are p0, p0 + stride, ...
    bool table_in_smem = false;
    size_t smem_bytes = 0;
};

namespace adtfe {

constexpr int kResampleThreads = 256;
constexpr int kResampleJ = 4;            // groups per thread
constexpr int kResampleP = 4;            // phases per thread
constexpr int kResampleSpanMax = 24576;  // floats of input staged per tile (96 KB)
constexpr size_t kResampleSmemMax = 200 * 1024;

// grid (tiles, rows).  A thread owns kResampleJ consecutive groups x kResampleP phases (p0, p0 + ps, p0 + 2 ps, ...):
// per tap kResampleP table values (lanes hold consecutive phases: conflict-free rows of the [k][p] table) and
// kResampleJ input samples (the same address across a warp: broadcasts) feed 16 FMAs, so the shared-memory pipe sees
// 0.5-0.75 wavefronts per warp FMA instead of the 1.25 of one load pair per FMA.  kTableInSmem: the filter bank is
// staged behind the input span (it fits for the usual rate pairs; 44.1 -> 16 kHz has 304 KB and stays in L1/L2).
// absmax_bits (optional): atomicMax of the float bits of |y| - non-negative floats order like integers and NaN's
// bits are above every number's, so a NaN wins, as in torch.max.
template <bool kTableInSmem>
__global__ void __launch_bounds__(kResampleThreads) resample_kernel(
    const float* __restrict__ x, int64_t ld_in, int64_t n_in, float* __restrict__ y, int64_t ld_out, int64_t n_out,
    const float* __restrict__ table, int orig, int neu, int width, int taps, int groups_per_tile, int phase_stride,
    int64_t n_tiles, int* __restrict__ absmax_bits) {
    extern __shared__ __align__(16) float xs[];
    const int tid = threadIdx.x;
    const float* xr = x + (int64_t)blockIdx.y * ld_in;
    float* yr = y + (int64_t)blockIdx.y * ld_out;
    const int span = (groups_per_tile - 1) * orig + taps;
    const float* tbl = table;
    if (kTableInSmem) {  // staged once: the CTA is persistent over its tiles
        float* ts = xs + ((span + 3) & ~3);
        for (int i = tid; i < taps * neu; i += kResampleThreads) ts[i] = __ldg(table + i);
        tbl = ts;
    }
    const int quads = groups_per_tile / kResampleJ;
    float m = 0.0f;
    bool nan = false;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t j0 = tile * groups_per_tile;
    const int64_t first = j0 * orig - width;  // x index of xs[0]
    __syncthreads();                          // the previous tile's span is no longer read
    for (int base = 0; base < span; base += 8 * kResampleThreads) {  // eight loads in flight per thread
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int i = base + u * kResampleThreads + tid;
            const int64_t g = first + i;
            v[u] = (i < span && g >= 0 && g < n_in) ? __ldg(xr + g) : 0.0f;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int i = base + u * kResampleThreads + tid;
            if (i < span) xs[i] = v[u];
        }
    }
    __syncthreads();
    for (int w = tid; w < quads * phase_stride; w += kResampleThreads) {
        const int jq = w / phase_stride, p0 = w - jq * phase_stride;
        const float* xa = xs + jq * kResampleJ * orig;
        int pi[kResampleP];   // phases past the last one read (and discard) phase p0
#pragma unroll
        for (int q = 0; q < kResampleP; ++q) pi[q] = p0 + q * phase_stride < neu ? p0 + q * phase_stride : p0;
        float acc[kResampleJ][kResampleP];
#pragma unroll
        for (int r = 0; r < kResampleJ; ++r)
#pragma unroll
            for (int q = 0; q < kResampleP; ++q) acc[r][q] = 0.0f;
        // one pointer per table column and per input row, bumped per tap: no index arithmetic in the loop
        const float* tp[kResampleP];
        const float* xp[kResampleJ];
#pragma unroll
        for (int q = 0; q < kResampleP; ++q) tp[q] = tbl + pi[q];
#pragma unroll
        for (int r = 0; r < kResampleJ; ++r) xp[r] = xa + r * orig;
#pragma unroll 4
        for (int k = 0; k < taps; ++k) {
            float t[kResampleP], v[kResampleJ];
#pragma unroll
            for (int q = 0; q < kResampleP; ++q) {
                t[q] = kTableInSmem ? *tp[q] : __ldg(tp[q]);
                tp[q] += neu;
            }
#pragma unroll
            for (int r = 0; r < kResampleJ; ++r) v[r] = *xp[r]++;
#pragma unroll
            for (int r = 0; r < kResampleJ; ++r)
#pragma unroll
                for (int q = 0; q < kResampleP; ++q) acc[r][q] = fmaf(v[r], t[q], acc[r][q]);
        }
#pragma unroll
        for (int r = 0; r < kResampleJ; ++r)
#pragma unroll
            for (int q = 0; q < kResampleP; ++q) {
                const int p = p0 + q * phase_stride;
                const int64_t o = (j0 + jq * kResampleJ + r) * neu + p;
                if (p < neu && o < n_out) {
                    yr[o] = acc[r][q];
                    nan |= acc[r][q] != acc[r][q];
                    m = fmaxf(m, fabsf(acc[r][q]));
                }
            }
    }
    }
    if (absmax_bits) {
        unsigned bits = nan ? 0x7fc00000u : __float_as_uint(m);
        bits = __reduce_max_sync(0xffffffffu, bits);
        if ((tid & 31) == 0 && bits != 0u) atomicMax(absmax_bits, (int)bits);
    }
}

// out[i] = (x[0][i] + x[1][i] + ...) / n_rows, channels added in order (torch.mean over dim 0)
__global__ void downmix_kernel(const float* __restrict__ x, int n_rows, int64_t ld, int64_t n, float* __restrict__ out) {
    const float count = (float)n_rows;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float s = x[i];
        for (int c = 1; c < n_rows; ++c) s += x[(int64_t)c * ld + i];
        out[i] = __fdiv_rn(s, count);
    }
}

__global__ void absmax_kernel(const float* __restrict__ x, int64_t n, int* __restrict__ absmax_bits) {
    float m = 0.0f;
    bool nan = false;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float v = x[i];
        nan |= v != v;
        m = fmaxf(m, fabsf(v));
    }
    unsigned bits = nan ? 0x7fc00000u : __float_as_uint(m);
    bits = __reduce_max_sync(0xffffffffu, bits);
    if ((threadIdx.x & 31) == 0 && bits != 0u) atomicMax(absmax_bits, (int)bits);
}

// x / max|x| (IEEE division, like torch); max = 0 gives the reference's 0/0 = NaN
__global__ void divide_kernel(float* __restrict__ x, int64_t n, const int* __restrict__ absmax_bits) {
    const float m = __int_as_float(*absmax_bits);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        x[i] = __fdiv_rn(x[i], m);
}

static int elementwise_grid(int64_t n, int sm_count) {
    const int64_t blocks = (n + 4 * 256 - 1) / (4 * 256);
    return (int)std::max<int64_t>(1, std::min<int64_t>(blocks, (int64_t)sm_count * 8));
}

}  // namespace adtfe

using namespace adtfe;

static int64_t gcd64(int64_t a, int64_t b) {
    while (b) { const int64_t t = a % b; a = b; b = t; }
    return a;
}

extern "C" int adtfe_resampler_create(int32_t orig_freq, int32_t new_freq, int32_t width, const float* kernel_host,
                                      int device, adtfe_resampler** out) {
    ADTFE_REQUIRE(out, ADTFE_ERR_BAD_ARG, "adtfe_resampler_create: null out");
    *out = nullptr;
    ADTFE_REQUIRE(orig_freq > 0 && new_freq > 0 && width >= 0 && kernel_host, ADTFE_ERR_BAD_ARG,
                  "adtfe_resampler_create: bad argument (orig=%d new=%d width=%d)", orig_freq, new_freq, width);
    const int64_t g = gcd64(orig_freq, new_freq);
    const int64_t o = orig_freq / g, n = new_freq / g, taps = 2 * (int64_t)width + o;
    ADTFE_REQUIRE(taps + (kResampleJ - 1) * o <= kResampleSpanMax && n * taps <= (1ll << 24), ADTFE_ERR_UNSUPPORTED,
                  "adtfe_resampler_create: %d -> %d Hz needs %lld taps x %lld phases (rates with a tiny gcd are not "
                  "supported)", orig_freq, new_freq, (long long)taps, (long long)n);
    int rc = adtfe_device_ok(device);
    if (rc != ADTFE_OK) return rc;
    ADTFE_CUDA(cudaSetDevice(device));
    std::vector<float> t((size_t)taps * n);
    for (int64_t p = 0; p < n; ++p)
        for (int64_t k = 0; k < taps; ++k) t[(size_t)k * n + p] = kernel_host[(size_t)p * taps + k];
    adtfe_resampler* r = new adtfe_resampler();
    r->device = device; r->orig = (int32_t)o; r->neu = (int32_t)n; r->width = width; r->taps = (int32_t)taps;
    // a tile: as many groups as give every thread of the CTA one (4 groups x 4 phases) item, a multiple of
    // kResampleJ, with the input span within the staging limit
    r->phase_stride = (int32_t)((n + kResampleP - 1) / kResampleP);
    int64_t groups = (int64_t)kResampleJ * std::max<int64_t>(1, kResampleThreads / r->phase_stride);
    groups = std::min<int64_t>(groups, (kResampleSpanMax - taps) / o + 1);
    groups = std::max<int64_t>(kResampleJ, groups / kResampleJ * kResampleJ);
    r->groups_per_tile = (int32_t)groups;
    const size_t span_bytes = (size_t)(((groups - 1) * o + taps + 3) & ~(int64_t)3) * 4;
    r->table_in_smem = span_bytes + (size_t)taps * n * 4 <= kResampleSmemMax;
    r->smem_bytes = span_bytes + (r->table_in_smem ? (size_t)taps * n * 4 : 0);
    if (cudaMalloc((void**)&r->table, t.size() * 4) != cudaSuccess ||
        cudaMemcpy(r->table, t.data(), t.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess ||
        (r->table_in_smem ? cudaFuncSetAttribute(resample_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 (int)kResampleSmemMax)
                          : cudaFuncSetAttribute(resample_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 (int)kResampleSmemMax)) != cudaSuccess) {
        set_error("adtfe_resampler_create: device set-up failed: %s", cudaGetErrorString(cudaGetLastError()));
        cudaFree(r->table);
        delete r;
        return ADTFE_ERR_CUDA;
    }
    *out = r;
    return ADTFE_OK;
}

extern "C" int adtfe_resampler_destroy(adtfe_resampler* r) {
    if (!r) return ADTFE_OK;
    cudaSetDevice(r->device);
    cudaFree(r->table);
    delete r;
    return ADTFE_OK;
}

extern "C" int64_t adtfe_resample_length(const adtfe_resampler* r, int64_t n_in) {
    if (!r || n_in < 0) return -1;
    return (n_in * r->neu + r->orig - 1) / r->orig;  // ceil(new * length / orig)
}

extern "C" int adtfe_resample(const adtfe_resampler* r, const float* x_dev, int32_t n_rows, int64_t ld_in, int64_t n_in,
                              float* y_dev, int64_t ld_out, int32_t* absmax_bits_dev, void* stream) {
    ADTFE_REQUIRE(r && n_rows >= 0 && n_in >= 0 && ld_in >= n_in, ADTFE_ERR_BAD_ARG,
                  "adtfe_resample: bad argument (n_rows=%d ld_in=%lld n_in=%lld)", n_rows, (long long)ld_in,
                  (long long)n_in);
    const int64_t n_out = adtfe_resample_length(r, n_in);
    ADTFE_REQUIRE(ld_out >= n_out, ADTFE_ERR_BAD_ARG, "adtfe_resample: ld_out %lld < %lld output samples",
                  (long long)ld_out, (long long)n_out);
    if (n_rows == 0 || n_out == 0) return ADTFE_OK;
    ADTFE_REQUIRE(x_dev && y_dev, ADTFE_ERR_BAD_ARG, "adtfe_resample: null buffer");
    ADTFE_REQUIRE(n_rows <= 65535, ADTFE_ERR_UNSUPPORTED, "adtfe_resample: more than 65535 rows");
    const int64_t per_tile = (int64_t)r->groups_per_tile * r->neu;
    const int64_t tiles = (n_out + per_tile - 1) / per_tile;
    ADTFE_REQUIRE(tiles < (1ll << 31), ADTFE_ERR_UNSUPPORTED, "adtfe_resample: signal too long for one launch");
    // persistent CTAs: two per SM and row (shared-memory bound), each walks its tiles with the filter bank staged once
    int device = 0, sms = 148;
    ADTFE_CUDA(cudaGetDevice(&device));
    sms = device_sm_count(device);
    const int64_t per_row = std::max<int64_t>(1, (2 * (int64_t)sms + n_rows - 1) / n_rows);
    const dim3 grid((unsigned)std::min<int64_t>(tiles, per_row), (unsigned)n_rows);
    if (r->table_in_smem)
        resample_kernel<true><<<grid, kResampleThreads, r->smem_bytes, (cudaStream_t)stream>>>(
            x_dev, ld_in, n_in, y_dev, ld_out, n_out, r->table, r->orig, r->neu, r->width, r->taps, r->groups_per_tile,
            r->phase_stride, tiles, (int*)absmax_bits_dev);
    else
        resample_kernel<false><<<grid, kResampleThreads, r->smem_bytes, (cudaStream_t)stream>>>(
            x_dev, ld_in, n_in, y_dev, ld_out, n_out, r->table, r->orig, r->neu, r->width, r->taps, r->groups_per_tile,
            r->phase_stride, tiles, (int*)absmax_bits_dev);
    ADTFE_CUDA(cudaGetLastError());
    return ADTFE_OK;
}

extern "C" int adtfe_downmix(const float* x_dev, int32_t n_rows, int64_t ld, int64_t n, float* out_dev, void* stream) {
    ADTFE_REQUIRE(n_rows >= 1 && n >= 0 && ld >= n, ADTFE_ERR_BAD_ARG, "adtfe_downmix: bad argument");
    if (n == 0) return ADTFE_OK;
    ADTFE_REQUIRE(x_dev && out_dev, ADTFE_ERR_BAD_ARG, "adtfe_downmix: null buffer");
    int device = 0;
    ADTFE_CUDA(cudaGetDevice(&device));
    downmix_kernel<<<elementwise_grid(n, device_sm_count(device)), 256, 0, (cudaStream_t)stream>>>(x_dev, n_rows, ld, n,
                                                                                                out_dev);
    ADTFE_CUDA(cudaGetLastError());
    return ADTFE_OK;
}

extern "C" int adtfe_peak_normalise(float* x_dev, int64_t n, int32_t* absmax_bits_dev, int32_t have_absmax,
                                    void* stream) {
    ADTFE_REQUIRE(n >= 0 && absmax_bits_dev, ADTFE_ERR_BAD_ARG, "adtfe_peak_normalise: bad argument");
    if (n == 0) return ADTFE_OK;
    ADTFE_REQUIRE(x_dev, ADTFE_ERR_BAD_ARG, "adtfe_peak_normalise: null buffer");
    int device = 0;
    ADTFE_CUDA(cudaGetDevice(&device));
    const int grid = elementwise_grid(n, device_sm_count(device));
    cudaStream_t st = (cudaStream_t)stream;
    if (!have_absmax) {
        ADTFE_CUDA(cudaMemsetAsync(absmax_bits_dev, 0, 4, st));
        absmax_kernel<<<grid, 256, 0, st>>>(x_dev, n, (int*)absmax_bits_dev);
        ADTFE_CUDA(cudaGetLastError());
    }
    divide_kernel<<<grid, 256, 0, st>>>(x_dev, n, (const int*)absmax_bits_dev);
    ADTFE_CUDA(cudaGetLastError());
    return ADTFE_OK;
}
