// libadtfe: error state, device checks, bank, fused and host-buffer entry points.
#include <stdarg.h>
#include <stdlib.h>

#include <vector>

#include "common.cuh"

namespace adtfe {

static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

// ---- launch trace: cudaEvent pairs around the kernels of adtfe_render / adtfe_render_logmel --------------------
struct TraceRec {
    const char* kernel;
    int index;
    cudaEvent_t a, b;
};
static std::vector<TraceRec> g_trace;
static bool g_trace_on = false;
static std::mutex g_trace_mu;

void trace_open(const char* kernel, int index, cudaStream_t st) {
    if (!g_trace_on) return;
    std::lock_guard<std::mutex> lock(g_trace_mu);
    TraceRec r = {kernel, index, nullptr, nullptr};
    if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return;
    cudaEventRecord(r.a, st);
    g_trace.push_back(r);
}
void trace_close(cudaStream_t st) {
    if (!g_trace_on) return;
    std::lock_guard<std::mutex> lock(g_trace_mu);
    if (!g_trace.empty()) cudaEventRecord(g_trace.back().b, st);
}

int device_sm_count(int device) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || n <= 0) n = 148;
    return n;
}

}  // namespace adtfe

using namespace adtfe;

extern "C" int adtfe_version(void) { return ADTFE_VERSION; }
extern "C" const char* adtfe_last_error(void) { return g_error; }

extern "C" int adtfe_trace_begin(void) {
    std::lock_guard<std::mutex> lock(g_trace_mu);
    for (auto& r : g_trace) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    g_trace.clear();
    g_trace_on = true;
    return ADTFE_OK;
}

extern "C" int adtfe_trace_dump(const char* path) {
    ADTFE_REQUIRE(path, ADTFE_ERR_BAD_ARG, "adtfe_trace_dump: null path");
    ADTFE_CUDA(cudaDeviceSynchronize());
    std::lock_guard<std::mutex> lock(g_trace_mu);
    g_trace_on = false;
    FILE* f = fopen(path, "w");
    ADTFE_REQUIRE(f, ADTFE_ERR_BAD_ARG, "adtfe_trace_dump: cannot write %s", path);
    fprintf(f, "kernel,index,start_ms,end_ms\n");
    for (auto& r : g_trace) {
        float t0 = 0.f, t1 = 0.f;
        if (cudaEventElapsedTime(&t0, g_trace.front().a, r.a) == cudaSuccess &&
            cudaEventElapsedTime(&t1, g_trace.front().a, r.b) == cudaSuccess)
            fprintf(f, "%s,%d,%.4f,%.4f\n", r.kernel, r.index, t0, t1);
    }
    for (auto& r : g_trace) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    cudaGetLastError();
    g_trace.clear();
    fclose(f);
    return ADTFE_OK;
}

extern "C" int adtfe_device_ok(int device) {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) {
        cudaGetLastError();
        set_error("no CUDA device %d (found %d)", device, count);
        return ADTFE_ERR_NO_DEVICE;
    }
    int major = 0;
    ADTFE_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
    ADTFE_REQUIRE(major == 10, ADTFE_ERR_NO_DEVICE,
                  "device %d is compute capability %d.x; this library is built for sm_100a only", device, major);
    return ADTFE_OK;
}

extern "C" int adtfe_bank_destroy(adtfe_bank* bank) {
    if (!bank) return ADTFE_OK;
    cudaSetDevice(bank->device);
    cudaFree(bank->pcm);
    cudaFree(bank->offsets);
    cudaFree(bank->lengths);
    cudaFree(bank->blockmax);
    cudaFree(bank->bm_off);
    for (int k = 0; k < bank->n_streams; ++k) {
        if (bank->streams[k]) cudaStreamDestroy(bank->streams[k]);
        if (bank->join_events[k]) cudaEventDestroy(bank->join_events[k]);
    }
    if (bank->fork_event) cudaEventDestroy(bank->fork_event);
    for (auto& stage : bank->stage_events)
        for (cudaEvent_t e : stage)
            if (e) cudaEventDestroy(e);
    delete bank;
    return ADTFE_OK;
}

extern "C" int64_t adtfe_bank_bytes(const adtfe_bank* bank) { return bank ? bank->total * 4 : 0; }

extern "C" int adtfe_bank_create(const float* pcm_host, int64_t total_floats, const int64_t* offsets_host,
                                 const int32_t* lengths_host, int32_t n_oneshots, int device, adtfe_bank** out) {
    ADTFE_REQUIRE(out, ADTFE_ERR_BAD_ARG, "adtfe_bank_create: null out");
    *out = nullptr;
    ADTFE_REQUIRE(total_floats >= 0 && n_oneshots >= 0, ADTFE_ERR_BAD_ARG, "adtfe_bank_create: negative size");
    ADTFE_REQUIRE(n_oneshots == 0 || (pcm_host && offsets_host && lengths_host), ADTFE_ERR_BAD_ARG,
                  "adtfe_bank_create: null array");
    for (int32_t i = 0; i < n_oneshots; ++i) {
        const int64_t padded = ((int64_t)lengths_host[i] + 3) & ~(int64_t)3;
        ADTFE_REQUIRE(offsets_host[i] >= 0 && offsets_host[i] % 4 == 0 && lengths_host[i] >= 0 &&
                          offsets_host[i] + padded <= total_floats,
                      ADTFE_ERR_BAD_ARG,
                      "adtfe_bank_create: one-shot %d (offset %lld, length %d) is misaligned or outside the bank "
                      "(starts must be multiples of 4 floats and storage padded to 4)",
                      i, (long long)offsets_host[i], lengths_host[i]);
    }
    int rc = adtfe_device_ok(device);
    if (rc != ADTFE_OK) return rc;
    ADTFE_CUDA(cudaSetDevice(device));
    adtfe_bank* b = new adtfe_bank();
    b->device = device;
    b->sm_count = device_sm_count(device);
    b->n = n_oneshots;
    b->total = total_floats;
    auto up = [&](void** dst, const void* src, size_t bytes) -> bool {
        if (cudaMalloc(dst, bytes ? bytes : 16) != cudaSuccess) return false;
        return bytes == 0 || cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice) == cudaSuccess;
    };
    if (!up((void**)&b->pcm, pcm_host, (size_t)total_floats * 4) ||
        !up((void**)&b->offsets, offsets_host, (size_t)n_oneshots * 8) ||
        !up((void**)&b->lengths, lengths_host, (size_t)n_oneshots * 4)) {
        set_error("adtfe_bank_create: upload failed: %s", cudaGetErrorString(cudaGetLastError()));
        adtfe_bank_destroy(b);
        return ADTFE_ERR_CUDA;
    }
    rc = mixer_prepare_device();
    if (rc == ADTFE_OK) rc = bank_build_blockmax(b, lengths_host);
    if (rc == ADTFE_OK) rc = fx_prepare_device();
    if (rc != ADTFE_OK) { adtfe_bank_destroy(b); return rc; }
    bool ok = cudaEventCreateWithFlags(&b->fork_event, cudaEventDisableTiming) == cudaSuccess;
    for (int k = 0; ok && k < kBankStreams; ++k) {
        ok = cudaStreamCreateWithFlags(&b->streams[k], cudaStreamNonBlocking) == cudaSuccess &&
             cudaEventCreateWithFlags(&b->join_events[k], cudaEventDisableTiming) == cudaSuccess;
        if (ok) b->n_streams = k + 1;
    }
    for (auto& stage : b->stage_events)
        for (cudaEvent_t& e : stage) ok = ok && cudaEventCreateWithFlags(&e, cudaEventDisableTiming) == cudaSuccess;
    if (!ok) {
        set_error("adtfe_bank_create: cannot create streams: %s", cudaGetErrorString(cudaGetLastError()));
        adtfe_bank_destroy(b);
        return ADTFE_ERR_CUDA;
    }
    *out = b;
    return ADTFE_OK;
}

// The log-mel of a render chunk is launched as soon as the chunk is normalised, on the caller's stream, and runs while
// the next chunks are rendered: the log-mel kernel is bound by instruction issue and shared memory and uses a tenth of
// the HBM bandwidth, the normalisation is bound by HBM and uses neither - under one log-mel launch at the end of the
// render the two ran one after the other.  (Peaks and the tile mixer need the shared memory the log-mel CTAs hold, so
// they take turns with them.)  Rows with an FX record are left out of the per-chunk launches and featurised by one more
// launch when the FX chain is through.
namespace {
struct MelPerChunk {
    const adtfe_bank* bank;
    const adtfe_mel* mel;
    const adtfe_plan* plan;
    const float* wav;
    float* out;
    cudaStream_t user;
    bool fx;
};
int mel_chunk_hook(void* ctx, int chunk, int seg0, int n_seg, const int* seg_fx, cudaStream_t done) {
    const MelPerChunk* m = (const MelPerChunk*)ctx;
    if (done != m->user) {
        cudaEvent_t e = m->bank->stage_events[2][chunk % kStageEvents];
        ADTFE_CUDA(cudaEventRecord(e, done));
        ADTFE_CUDA(cudaStreamWaitEvent(m->user, e, 0));
    }
    trace_open("logmel", chunk, m->user);
    const int rc = logmel_rows_filtered(m->mel, m->wav + (size_t)seg0 * m->plan->ld_wav, n_seg, m->plan->ld_wav,
                                        m->plan->mel_rows_dev + seg0, m->plan->mel_max_count, m->out,
                                        m->fx ? seg_fx + seg0 : nullptr, 0, m->user);
    trace_close(m->user);
    return rc;
}
}  // namespace

extern "C" int adtfe_render_logmel(const adtfe_bank* bank, const adtfe_mel* mel, const adtfe_plan* plan,
                                   int64_t n_samples, float* wav_out_dev, float* mel_out_dev, void* workspace_dev,
                                   size_t workspace_bytes, void* stream) {
    ADTFE_REQUIRE(plan && (plan->mel_rows_dev || n_samples <= plan->ld_wav), ADTFE_ERR_BAD_ARG,
                  "adtfe_render_logmel: n_samples exceeds the row pitch");
    if (plan->mel_rows_dev && plan->n_chunks > 1 && plan->chunks_host && mel && logmel_has_row_filter(mel) &&
        mel_out_dev && plan->mel_max_count > 0) {
        ADTFE_REQUIRE(((uintptr_t)plan->mel_rows_dev & 15) == 0 &&
                          (int64_t)mel->wpi * mel->hop >= 1024 &&
                          (int64_t)(mel->wpi + plan->mel_max_count - 1) * mel->hop + 1024 <= plan->ld_wav,
                      ADTFE_ERR_UNSUPPORTED, "adtfe_render_logmel: frame support leaves the row");
        MelPerChunk m{bank, mel, plan, wav_out_dev, mel_out_dev, (cudaStream_t)stream, plan->n_fx > 0};
        const ChunkHook hook{mel_chunk_hook, &m};
        const int* seg_fx = nullptr;
        int rc = render_impl(bank, plan, wav_out_dev, workspace_dev, workspace_bytes, stream, &hook, &seg_fx);
        if (rc != ADTFE_OK || plan->n_fx == 0) return rc;
        trace_open("logmel_fx", -1, (cudaStream_t)stream);   // the FX rows: the render has joined its FX stream into `stream`
        rc = logmel_rows_filtered(mel, wav_out_dev, plan->n_seg, plan->ld_wav, plan->mel_rows_dev, plan->mel_max_count,
                                  mel_out_dev, seg_fx, 1, stream);
        trace_close((cudaStream_t)stream);
        return rc;
    }
    int rc = adtfe_render(bank, plan, wav_out_dev, workspace_dev, workspace_bytes, stream);
    if (rc != ADTFE_OK) return rc;
    if (plan->mel_rows_dev) {
        trace_open("logmel", -1, (cudaStream_t)stream);
        rc = adtfe_logmel_rows(mel, wav_out_dev, plan->n_seg, plan->ld_wav, plan->mel_rows_dev, plan->mel_max_count,
                               mel_out_dev, stream);
        trace_close((cudaStream_t)stream);
        return rc;
    }
    return adtfe_logmel(mel, wav_out_dev, plan->n_seg, plan->ld_wav, n_samples, mel_out_dev, stream);
}

static size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }

extern "C" int adtfe_plan_blob_layout(const adtfe_plan* s, size_t offsets[7], size_t* blob_bytes) {
    ADTFE_REQUIRE(s && offsets && blob_bytes, ADTFE_ERR_BAD_ARG, "adtfe_plan_blob_layout: null pointer");
    ADTFE_REQUIRE(s->n_events >= 0 && s->n_seg >= 0 && s->tiles_per_seg >= 0 && s->n_peak_work >= 0 && s->n_fx >= 0,
                  ADTFE_ERR_BAD_ARG, "adtfe_plan_blob_layout: negative count");
    size_t o = 0;
    offsets[0] = o; o = align16(o + (size_t)s->n_events * sizeof(adtfe_event));
    offsets[1] = o; o = align16(o + (size_t)s->n_seg * sizeof(adtfe_segment));
    offsets[2] = o; o = align16(o + ((size_t)s->n_seg * s->tiles_per_seg + 1) * 4);
    offsets[3] = o; o = align16(o + (size_t)s->n_peak_work * sizeof(adtfe_peak_item));
    offsets[4] = o; o = align16(o + (s->mel_total_rows > 0 ? (size_t)s->n_seg * sizeof(adtfe_mel_row) : 0));
    offsets[5] = o; o = align16(o + (size_t)s->n_fx * sizeof(adtfe_fx));
    offsets[6] = o;  // tile_events runs to the end of the blob; its length is tile_ptr's last entry
    *blob_bytes = o;
    return ADTFE_OK;
}

extern "C" int adtfe_frontend_host(const adtfe_bank* bank, const adtfe_mel* mel, const adtfe_plan* shape,
                                   int64_t n_samples, const void* blob_host, size_t blob_bytes, void* blob_dev,
                                   float* wav_dev, float* mel_dev, void* workspace_dev, size_t workspace_bytes,
                                   float* mel_out_host, float* wav_out_host, void* stream, void* copy_stream) {
    ADTFE_REQUIRE(bank && mel && shape && blob_host && blob_dev && mel_out_host, ADTFE_ERR_BAD_ARG,
                  "adtfe_frontend_host: null pointer");
    size_t off[7], fixed = 0;
    int rc = adtfe_plan_blob_layout(shape, off, &fixed);
    if (rc != ADTFE_OK) return rc;
    ADTFE_REQUIRE(blob_bytes >= fixed, ADTFE_ERR_BAD_ARG, "adtfe_frontend_host: plan blob %zu B < %zu B", blob_bytes,
                  fixed);
    cudaStream_t st = (cudaStream_t)stream;
    ADTFE_CUDA(cudaMemcpyAsync(blob_dev, blob_host, blob_bytes, cudaMemcpyHostToDevice, st));
    adtfe_plan p = *shape;
    const char* d = (const char*)blob_dev;
    p.events_dev = (const adtfe_event*)(d + off[0]);
    p.segments_dev = (const adtfe_segment*)(d + off[1]);
    p.tile_ptr_dev = (const int32_t*)(d + off[2]);
    p.peak_work_dev = (const adtfe_peak_item*)(d + off[3]);
    p.mel_rows_dev = shape->mel_total_rows > 0 ? (const adtfe_mel_row*)(d + off[4]) : nullptr;
    p.fx_dev = shape->n_fx > 0 ? (const adtfe_fx*)(d + off[5]) : nullptr;
    p.tile_events_dev = (const int32_t*)(d + off[6]);
    rc = adtfe_render_logmel(bank, mel, &p, n_samples, wav_dev, mel_dev, workspace_dev, workspace_bytes, stream);
    if (rc != ADTFE_OK) return rc;
    int32_t first = 0, count = 0;
    adtfe_mel_frames(mel, n_samples, &first, &count);
    const size_t mel_bytes = p.mel_rows_dev ? (size_t)p.mel_total_rows * mel->n_mels * 4
                                            : (size_t)p.n_seg * count * mel->n_mels * 4;
    cudaStream_t cs = st;
    if (copy_stream && copy_stream != stream) {  // results leave on their own stream, after the kernels
        cs = (cudaStream_t)copy_stream;
        std::lock_guard<std::mutex> lock(bank->mu);
        ADTFE_CUDA(cudaEventRecord(bank->fork_event, st));
        ADTFE_CUDA(cudaStreamWaitEvent(cs, bank->fork_event, 0));
    }
    if (mel_bytes) ADTFE_CUDA(cudaMemcpyAsync(mel_out_host, mel_dev, mel_bytes, cudaMemcpyDeviceToHost, cs));
    if (wav_out_host && p.n_seg)
        ADTFE_CUDA(cudaMemcpyAsync(wav_out_host, wav_dev, (size_t)p.n_seg * p.ld_wav * 4, cudaMemcpyDeviceToHost, cs));
    return ADTFE_OK;
}
