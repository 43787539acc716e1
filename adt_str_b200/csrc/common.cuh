// Shared declarations for libadtfe (sm_100a).  See include/adtfe.h for the ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <mutex>
#include <string>

#include "../../include/adtfe.h"

namespace adtfe {

void set_error(const char* fmt, ...);

#define ADTFE_CUDA(call)                                                                   \
    do {                                                                                   \
        cudaError_t err__ = (call);                                                        \
        if (err__ != cudaSuccess) {                                                        \
            adtfe::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(err__), __FILE__, __LINE__); \
            return ADTFE_ERR_CUDA;                                                         \
        }                                                                                  \
    } while (0)

#define ADTFE_REQUIRE(cond, status, ...)       \
    do {                                       \
        if (!(cond)) {                         \
            adtfe::set_error(__VA_ARGS__);     \
            return (status);                   \
        }                                      \
    } while (0)

// Event with its bank lookups resolved by the peak pass, so the tile mixer does one
// broadcast load per event instead of a three-deep dependent chain (48 bytes).
struct ResolvedEvent {
    int64_t a_off, b_off;  // float offsets of the main / sub one-shot in the bank
    int32_t la, lb;        // their lengths, cut to the note's copy length
    int32_t start, len;
    float ca, cb;
    float gain;            // vel_to_vol * instrument gain; divided by the peak in the mixer
    int32_t pad;
};
static_assert(sizeof(ResolvedEvent) == 48, "ResolvedEvent layout");
static_assert(sizeof(adtfe_event) == 32, "adtfe_event layout");
static_assert(sizeof(adtfe_segment) == 16, "adtfe_segment layout");
static_assert(sizeof(adtfe_peak_item) == 40, "adtfe_peak_item layout");
static_assert(sizeof(adtfe_fx) == 48 && sizeof(adtfe_chunk) == 16, "adtfe_fx / adtfe_chunk layout");

int device_sm_count(int device);

#ifdef __CUDACC__
// ---- mbarrier / TMA bulk-copy helpers (sm_90+ PTX) ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("{ .reg .b64 t; mbarrier.arrive.shared::cta.b64 t, [%0]; }" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("{ .reg .b64 t; mbarrier.arrive.expect_tx.shared::cta.b64 t, [%0], %1; }" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// the same, letting the hardware park the thread for up to `ns` nanoseconds per try: a waiter that expects a
// long wait (data still on its way from L2) then leaves the issue slots to the other warps of the SM
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity, uint32_t ns) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAITR_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@p bra DONER_%=;\n\t"
        "bra WAITR_%=;\n\t"
        "DONER_%=:\n\t}" ::"r"(smem_u32(bar)),
        "r"(parity), "r"(ns)
        : "memory");
}
// TMA 1-D bulk copy global -> shared, completion counted in bytes on `bar`
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
#endif

}  // namespace adtfe

struct adtfe_mel_tables;
struct adtfe_bank;
struct adtfe_mel;

namespace adtfe {
int mixer_prepare_device();
int bank_build_blockmax(struct ::adtfe_bank* b, const int32_t* lengths_host);
int fx_prepare_device();
// FX kernels over the rows fx_dev[r0 .. r0 + n_rows) of the plan, between the tile mixer and the normalisation
int fx_reverb_launch(const adtfe_plan* plan, int r0, int n_rows, float* wav, cudaStream_t st);
int fx_dynamics_launch(const adtfe_plan* plan, int r0, int n_rows, float* wav, float* tile_max, int max_per_seg,
                       cudaStream_t st);
// Called by render_impl behind the normalisation of every chunk (rows [seg0, seg0 + n_seg) are final except those
// marked in seg_fx, which wait for the FX chain): `done` is on the stream that ran the normalisation.
struct ChunkHook {
    int (*fn)(void* ctx, int chunk, int seg0, int n_seg, const int* seg_fx, cudaStream_t done);
    void* ctx;
};
int render_impl(const adtfe_bank* bank, const adtfe_plan* plan, float* wav_out_dev, void* workspace_dev,
                size_t workspace_bytes, void* stream, const ChunkHook* hook = nullptr, const int** seg_fx_out = nullptr);
bool logmel_has_row_filter(const struct ::adtfe_mel* mel);
int logmel_rows_filtered(const struct ::adtfe_mel* mel, const float* wav_dev, int32_t n_seg, int64_t ld_wav,
                         const adtfe_mel_row* rows_dev, int32_t max_count, float* out_dev, const int* seg_mark,
                         int mark_want, void* stream);
// Diagnostics (adtfe_trace_begin / adtfe_trace_dump): a pair of timing events around every kernel launch.
void trace_open(const char* kernel, int index, cudaStream_t st);
void trace_close(cudaStream_t st);
}  // namespace adtfe

#ifndef ADTFE_BANK_STREAMS
#define ADTFE_BANK_STREAMS 7
#endif
constexpr int kBankStreams = ADTFE_BANK_STREAMS;
constexpr int kStageEvents = 8;
struct adtfe_bank {
    int device = 0;
    int sm_count = 148;
    int32_t n = 0;
    int64_t total = 0;
    float* pcm = nullptr;
    int64_t* offsets = nullptr;
    int32_t* lengths = nullptr;
    // max |x| of every ADTFE_PEAK_BLOCK-sample block of every one-shot (blocks of one-shot i start at bm_off[i]): the
    // peak pass bounds the mixed one-shot's blocks with them and scans only the blocks that can hold the peak
    float* blockmax = nullptr;
    int32_t* bm_off = nullptr;
    // internal streams for chunked plans: forked from / joined into the caller's stream
    int n_streams = 0;
    cudaStream_t streams[kBankStreams] = {};
    cudaEvent_t fork_event = nullptr, join_events[kBankStreams] = {};
    // pipeline of a chunked render: stage_events[k][c % kStageEvents] = chunk c left stage k (an event can be
    // re-recorded as soon as the wait on its previous recording has been enqueued)
    cudaEvent_t stage_events[3][kStageEvents] = {};   // [2]: chunk c is normalised (the log-mel of the chunk waits for it)
    mutable std::mutex mu;
};

struct adtfe_mel {
    int device = 0;
    int sm_count = 148;
    int32_t n_fft = 2048, hop = 240, n_mels = 128, wpi = 5;
    int32_t nnz = 0;        // stored filter weights (first..last non-zero bin of every filter)
    float* window = nullptr;   // n_fft
    float2* twiddle = nullptr; // 32 x 32: W_2048^(k1*n2), k1 = 1..32, n2 = lane
    float2* lane_tw = nullptr; // 4 x 32: per-lane twiddles of the cross-lane 32-point DFT (stages 0..3)
    float* weights = nullptr;  // filter weights in mel-phase order (groups of (up, down) pairs, or per filter)
    void* groups = nullptr;    // fast path: MelGroup records {first bin, S row to flush into}
    // warp-autonomous kernel (v6): lane-walk weights, per-lane masks, per-filter partial-sum lists
    int32_t v6_ok = 0;
    void *w6 = nullptr, *lane6 = nullptr, *comb6 = nullptr;
    size_t smem6_bytes = 0;
    int32_t force_generic = 0; // adtfe_mel_force_generic: take the generic kernel even where v6 applies (cross-checks)
    int32_t fast_path = 0;     // 1: the filterbank is triangular (<= 2 adjacent filters per bin)
    struct adtfe_mel_tables* tables = nullptr;  // mel-phase items + warp schedule, passed as a kernel parameter
    size_t smem_bytes = 0;
};
