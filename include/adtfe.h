/*
 * adtfe - B200 (sm_100a) render + log-mel front end, C ABI.
 *
 * The reference (pier-maker92/ADT_STR) is pure Python, so there is no FFI in it to
 * bind; this header is the boundary a maintainer would bind with ctypes (see
 * INTEGRATION.md).  Every entry point names the reference code it replaces:
 *
 *   adtfe_bank_*          one-shot storage: the per-note h5py.File open + two gzip dataset
 *                         reads of modules/synthetiser.py:273,283-284 (bank layout written by
 *                         data_modules/convert_augmented_to_hdf5.py:70-138)
 *   adtfe_render          SynthDrum.__call__ audio arithmetic: drum_rendering
 *                         (modules/synthetiser.py:214-239), VolumeMixer.instrument_mixer and
 *                         _normalize_audio (:142-156), and the zero padding of collate_fn
 *                         (data_modules/train_dataset.py:53)
 *   adtfe_mel_* / adtfe_logmel
 *                         ComputeMelSpectrogram.__init__/forward (model.py:68-97) incl. the
 *                         torchaudio MelSpectrogram it delegates to (model.py:71-78,89)
 *   adtfe_render_logmel   both, back to back on one stream (train.py:52 H2D + model.py:248)
 *   adtfe_frontend_host   the same with HOST buffers: plan blob in, log-mel (and optionally
 *                         the waveform) out, copies included - the end-to-end entry
 *   adtfe_linear_*        ADTModel.project_to_mel (model.py:224-226, 249): the bf16 Linear the log-mel feeds
 *   adtfe_resample* / adtfe_downmix / adtfe_peak_normalise
 *                         the audio front of eval and inference: utils/audio_utils.py:17-23
 *                         (resample = torchaudio.transforms.Resample, normalize = wav / max|wav|),
 *                         the channel mean of utils/audio_utils.py:12 and inference.py:86-87; the
 *                         chunking of inference.py:35-48 is then a view of the zero-padded signal
 *
 * Conventions: plain pointers and sizes; pointers named *_dev are device memory on the
 * bank's / mel's device, *_host are host memory (pinned for asynchronous copies).  All
 * work is enqueued on `stream` (a cudaStream_t passed as void*; NULL = legacy default
 * stream) and is asynchronous unless stated.  The library never allocates memory the
 * caller sees; scratch comes from a caller workspace sized by adtfe_*_workspace_bytes.
 * Every function returns ADTFE_OK (0) or a negative adtfe_status; adtfe_last_error()
 * gives a thread-local message for the last failure.
 */
#ifndef ADTFE_H
#define ADTFE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ADTFE_VERSION 5
#define ADTFE_TILE 2048      /* output samples owned by one mixer CTA */
#define ADTFE_PEAK_BLOCK 256 /* samples per block of the bank's block maxima (peak pass: branch and bound) */
#define ADTFE_PEAK_NOTES 8   /* notes the peak pass bounds together: planners cut a group into items of that many */

typedef enum adtfe_status {
    ADTFE_OK = 0,
    ADTFE_ERR_BAD_ARG = -1,     /* null pointer, negative size, misaligned pitch ... */
    ADTFE_ERR_UNSUPPORTED = -2, /* n_fft != 2048, n_mels > 256, hop < 1 ... */
    ADTFE_ERR_WORKSPACE = -3,   /* workspace too small */
    ADTFE_ERR_CUDA = -4,        /* a CUDA call failed; message holds cudaGetErrorString */
    ADTFE_ERR_NO_DEVICE = -5    /* no sm_100 device / wrong architecture */
} adtfe_status;

/* One rendered note (32 bytes).  Events of a batch are sorted by (seg, instrument
 * first-appearance, note order); the mixer adds them in array order. */
typedef struct adtfe_event {
    int32_t start;   /* first output sample: (int)(onset * sr), float math as the reference */
    int32_t len;     /* samples written: min(max(len_main, len_sub), seg_len - start), >= 0 */
    int32_t main_id; /* one-shot ids in the bank */
    int32_t sub_id;
    float ca;        /* (float)(1 - mixup) */
    float cb;        /* (float)mixup */
    float gain;      /* vel_to_vol(velocity) * instrument gain */
    int32_t seg;     /* segment (row of the waveform matrix) */
} adtfe_event;

/* One segment = one SynthDrum.__call__ (16 bytes). */
typedef struct adtfe_segment {
    int32_t len;         /* wave_length in samples; the row is zero beyond it */
    int32_t flags;       /* 0: no notes, all-zero row; 1: wav / max|wav| * max_volume; ADTFE_SEG_RAW (2): the row keeps
                          * the raw instrument sum (no normalisation) - for a caller that post-processes it itself */
    float max_volume;    /* vel_to_vol(max velocity) */
    int32_t first_event; /* index of the segment's first event */
} adtfe_segment;

#define ADTFE_SEG_RAW 2

/* One work item of the peak pass (40 bytes): the mixed one-shot shared by the notes first_event ..
 * first_event+n_events-1 (notes of one instrument of one segment: they share the two one-shots).  Any number of
 * notes per item is handled; the planners cut a group into items of ADTFE_PEAK_NOTES notes (one warp each).  chunk = 0
 * (ABI <= 4 split a group into per-span items chunk = 0, 1, ...: items with chunk != 0 are ignored).  Offsets and lengths are the bank's,
 * resolved on the host so the kernel starts its loads after a single record fetch; events[first_event].main_id /
 * sub_id name the two one-shots (block maxima lookup). */
typedef struct adtfe_peak_item {
    int64_t a_off, b_off; /* float offsets of the main / sub one-shot in the bank */
    int32_t la, lb;       /* their lengths */
    int32_t mix_len;      /* max(la, lb) */
    int32_t first_event, n_events, chunk;
} adtfe_peak_item;

/* Ragged log-mel: segment s writes `count` frames (the first kept frame is always the window-pad index)
 * to rows out_row .. out_row+count-1 of the (rows, n_mels) output matrix.  Lets one launch featurise many
 * collated batches at once - each batch has its own width, hence its own frame count (model.py:95-97). */
typedef struct adtfe_mel_row {
    int64_t out_row;
    int32_t count; /* >= 0; the frames must lie inside the row: (first+count-1)*hop + n_fft/2 <= ld_wav */
    int32_t flags; /* 0, or ADTFE_MEL_ROW_SILENT */
} adtfe_mel_row;
/* The caller knows the row is all zeros (an empty segment: synthetiser.py:257-258 returns zeros): its log-mel frames
 * are exactly 0.0 (log(1e-10) clamped at -23) and are written without running the transform.  Purely a hint - a row
 * of zeros without the flag gives the same output. */
#define ADTFE_MEL_ROW_SILENT 1

/* A chunk of a plan: segments, events, peak work items and FX rows [x[c], x[c+1]) are rendered together;
 * chunks are independent and go through the library's internal streams as a pipeline (so that a chunk's
 * one-shots are still in L2 when its tile mixer re-reads them, and the kernels of different chunks overlap). */
typedef struct adtfe_chunk {
    int32_t seg, event, peak_work, fx_row;
} adtfe_chunk;

/* FX chain of one segment (48 bytes): VolumeMixer._add_fx / BoardChain of the reference
 * (modules/synthetiser.py:30-87,121-137), applied between the instrument sum and the global normalisation
 * (:154).  The parameters are the keyword arguments the reference hands to pedalboard.Reverb / Compressor /
 * Limiter; the host planner draws them from the reference's RNG streams in the reference's order.  Records are
 * sorted by `seg`; segments without a record get no FX. */
#define ADTFE_FX_REVERB 1
#define ADTFE_FX_COMPRESSOR 2
#define ADTFE_FX_LIMITER 4
typedef struct adtfe_fx {
    int32_t seg;
    int32_t flags; /* ADTFE_FX_* (0: the FX coin hit but no plugin was drawn: the empty board) */
    float room_size, damping, wet_level, dry_level, width;                 /* Reverb (freeze_mode 0) */
    float comp_threshold_db, comp_ratio, comp_attack_ms, comp_release_ms; /* Compressor */
    float lim_threshold_db;                                                /* Limiter (release_ms 100) */
} adtfe_fx;

typedef struct adtfe_bank adtfe_bank; /* one-shot bank resident in HBM */
typedef struct adtfe_mel adtfe_mel;   /* window, mel filterbank (CSR) and twiddles on device */

/* A planned batch, all arrays in device memory. */
typedef struct adtfe_plan {
    const adtfe_event* events_dev;        /* n_events */
    const adtfe_segment* segments_dev;    /* n_seg */
    const int32_t* tile_ptr_dev;          /* n_seg*tiles_per_seg+1: CSR tile -> tile_events */
    const int32_t* tile_events_dev;       /* event ids, ascending inside a tile */
    const adtfe_peak_item* peak_work_dev; /* n_peak_work */
    int32_t n_events, n_seg, tiles_per_seg, n_peak_work;
    int64_t ld_wav; /* row pitch of the waveform matrix in floats, multiple of 4, <= tiles_per_seg*ADTFE_TILE */
    /* optional (NULL / 0 = absent) */
    const adtfe_mel_row* mel_rows_dev;   /* n_seg rows: ragged log-mel output (several collated batches per plan) */
    int64_t mel_total_rows;              /* rows of the log-mel matrix when mel_rows_dev is given */
    int32_t mel_max_count;               /* largest count in mel_rows_dev */
    int32_t n_chunks;                    /* boundaries in chunks_host: n_chunks + 1 records, first all-zero, */
    const adtfe_chunk* chunks_host;      /* last = {n_seg, n_events, n_peak_work}; HOST memory */
    int32_t n_tile_events;               /* entries of tile_events_dev (= the last tile_ptr entry); required */
    /* optional FX chain (NULL / 0 = no segment has FX) */
    int32_t n_fx;                        /* records in fx_dev */
    const adtfe_fx* fx_dev;              /* sorted by seg */
    int32_t sample_rate;                 /* of the rendered audio; required when n_fx > 0 (filter tunings) */
} adtfe_plan;

int adtfe_version(void);
const char* adtfe_last_error(void);
/* 0 when device `device` exists and is compute capability 10.x */
int adtfe_device_ok(int device);

/* ---- bank -------------------------------------------------------------------------- */
/* Copies `total_floats` floats of PCM (all one-shots, each start a multiple of 4 floats)
 * and the per-one-shot offsets/lengths to `device`.  Synchronous. */
int adtfe_bank_create(const float* pcm_host, int64_t total_floats, const int64_t* offsets_host,
                      const int32_t* lengths_host, int32_t n_oneshots, int device, adtfe_bank** out);
int adtfe_bank_destroy(adtfe_bank* bank);
int64_t adtfe_bank_bytes(const adtfe_bank* bank);

/* ---- render ------------------------------------------------------------------------ */
size_t adtfe_render_workspace_bytes(int32_t n_events, int32_t n_seg, int32_t tiles_per_seg, int32_t n_tile_events);
/* Writes the (n_seg, ld_wav) float32 waveform matrix: every row normalised as the reference
 * does and zero-padded to ld_wav.  Four kernels per chunk: per-note peak of the mixed one-shot (branch and bound
 * over the bank's block maxima), the per-tile slice records, the tile mixer, the row normalisation.  With
 * plan->n_fx > 0 the rows with an FX record are left out of the per-chunk normalisation: their reverb runs behind
 * the chunk's mixer, then one dynamics launch and one normalisation cover all the plan's FX rows.  A plan with chunks
 * (plan->chunks_host) is rendered chunk by chunk on the bank's internal streams, forked from and joined back into
 * `stream`; calls on one bank handle must not be made from several host threads at once.  The waveform matrix must
 * not be read between the chunks: rows are final when the call's work on `stream` is. */
int adtfe_render(const adtfe_bank* bank, const adtfe_plan* plan, float* wav_out_dev, void* workspace_dev,
                 size_t workspace_bytes, void* stream);

/* ---- log-mel ----------------------------------------------------------------------- */
/* window_host: n_fft floats; fb_host: (n_fft/2+1, n_mels) row-major float32 - the two
 * buffers torchaudio's MelSpectrogram keeps in the state dict.  Synchronous. */
int adtfe_mel_create(int32_t n_fft, int32_t hop, int32_t n_mels, const float* window_host, const float* fb_host,
                     int device, adtfe_mel** out);
int adtfe_mel_destroy(adtfe_mel* mel);
/* Frames the reference keeps for an n_samples-long input: t = *first .. *first+*count-1 of
 * the centred STFT (model.py:79,95-97). */
int adtfe_mel_frames(const adtfe_mel* mel, int64_t n_samples, int32_t* first, int32_t* count);
/* 1 when the filterbank has the triangular structure (at most two adjacent filters per bin) the fast
 * mel phase needs, 0 when the per-filter path is used (any fb works), < 0 on a null handle. */
int adtfe_mel_fast_path(const adtfe_mel* mel);
/* on != 0: this handle runs the generic kernel (any hop, any filterbank) even where the warp-autonomous one applies -
 * the two are cross-checked against each other in the tests.  Per handle; not meant to be flipped while launches of
 * the handle are being enqueued from another thread. */
int adtfe_mel_force_generic(adtfe_mel* mel, int32_t on);
/* wav_dev: (n_seg, ld_wav) rows of n_samples valid floats; out_dev: (n_seg, count, n_mels). */
int adtfe_logmel(const adtfe_mel* mel, const float* wav_dev, int32_t n_seg, int64_t ld_wav, int64_t n_samples,
                 float* out_dev, void* stream);
/* Ragged form: one launch over rows with their own frame counts (rows_dev[n_seg], max_count = the largest). */
int adtfe_logmel_rows(const adtfe_mel* mel, const float* wav_dev, int32_t n_seg, int64_t ld_wav,
                      const adtfe_mel_row* rows_dev, int32_t max_count, float* out_dev, void* stream);

/* ---- fused call -------------------------------------------------------------------- */
/* adtfe_render then adtfe_logmel over the first n_samples (<= plan->ld_wav) floats of every
 * row: n_samples is the collated batch width (longest segment), which sets the frame count.
 * With plan->mel_rows_dev the ragged form is used instead and n_samples is ignored; a plan with chunks then gets
 * the log-mel of every chunk launched behind the chunk's normalisation (on `stream`, while later chunks are being
 * rendered) and the rows with an FX record featurised by one more launch at the end - same results. */
int adtfe_render_logmel(const adtfe_bank* bank, const adtfe_mel* mel, const adtfe_plan* plan, int64_t n_samples,
                        float* wav_out_dev, float* mel_out_dev, void* workspace_dev, size_t workspace_bytes,
                        void* stream);
/* ---- host-buffer entry (end to end) ------------------------------------------------ */
/* Plan blob layout (host, 16-byte aligned sections in this order):
 *   events | segments | tile_ptr | peak_work | mel_rows | fx | tile_events
 * with the counts in `shape` (a plan whose device pointers are ignored; mel_rows is n_seg records when
 * shape->mel_total_rows > 0, else empty; fx is shape->n_fx records; shape->chunks_host is used as given).  The blob is copied to
 * `blob_dev` (>= blob_bytes), the batch rendered and featurised, then the log-mel matrix
 * (and the waveform when wav_out_host != NULL) copied back.  Asynchronous on `stream`:
 * the host buffers must be pinned and stay alive until the stream is synchronised.
 * copy_stream: NULL, or a second stream for the device->host copies - they then wait for the kernels through
 * an event and overlap the next call's work on `stream`; synchronise copy_stream before reading the results. */
int adtfe_frontend_host(const adtfe_bank* bank, const adtfe_mel* mel, const adtfe_plan* shape, int64_t n_samples,
                        const void* blob_host, size_t blob_bytes, void* blob_dev, float* wav_dev, float* mel_dev,
                        void* workspace_dev, size_t workspace_bytes, float* mel_out_host, float* wav_out_host,
                        void* stream, void* copy_stream);
/* Byte offsets of the seven sections inside a plan blob (offsets[7]; *blob_bytes = offsets[6], where
 * tile_events starts - it runs to the end of the blob). */
int adtfe_plan_blob_layout(const adtfe_plan* shape, size_t offsets[7], size_t* blob_bytes);

/* ---- long-form audio front (eval / inference) ------------------------------------------ */
typedef struct adtfe_resampler adtfe_resampler;
/* kernel_host: the (new_freq/gcd, 2*width + orig_freq/gcd) float32 filter bank torchaudio's Resample keeps in its
 * `kernel` buffer (functional._get_sinc_resample_kernel), width its `width`.  Synchronous.  Rate pairs whose
 * gcd is so small that the bank would not fit (taps > ~24k or > 16M coefficients) return ADTFE_ERR_UNSUPPORTED. */
int adtfe_resampler_create(int32_t orig_freq, int32_t new_freq, int32_t width, const float* kernel_host, int device,
                           adtfe_resampler** out);
int adtfe_resampler_destroy(adtfe_resampler* resampler);
/* ceil(new_freq * n_in / orig_freq): samples torchaudio returns for an n_in-long input; < 0 on a bad argument */
int64_t adtfe_resample_length(const adtfe_resampler* resampler, int64_t n_in);
/* x_dev: n_rows rows (channels / items) of n_in floats, pitch ld_in; y_dev: rows of adtfe_resample_length floats,
 * pitch ld_out.  absmax_bits_dev: NULL, or one zero-initialised int32 that receives the float bits of max|y|
 * over everything written (NaN if any output is NaN) - hand it to adtfe_peak_normalise with have_absmax = 1. */
int adtfe_resample(const adtfe_resampler* resampler, const float* x_dev, int32_t n_rows, int64_t ld_in, int64_t n_in,
                   float* y_dev, int64_t ld_out, int32_t* absmax_bits_dev, void* stream);
/* out[i] = (x[0][i] + ... + x[n_rows-1][i]) / n_rows - torch.mean over the channel dimension */
int adtfe_downmix(const float* x_dev, int32_t n_rows, int64_t ld, int64_t n, float* out_dev, void* stream);
/* x / max|x| in place (IEEE division; an all-zero signal gives NaN like the reference's 0/0).  absmax_bits_dev:
 * one int32 of scratch; have_absmax != 0 when it already holds the bits of max|x| (from adtfe_resample). */
int adtfe_peak_normalise(float* x_dev, int64_t n, int32_t* absmax_bits_dev, int32_t have_absmax, void* stream);

/* ---- project_to_mel (the layer behind the log-mel) ------------------------------------- */
/* nn.Linear(n_mels, d_query * nhead) of the reference model (model.py:224-226), applied to the log-mel at
 * model.py:249 under bf16 autocast: out = bf16(bf16(x) @ bf16(W)^T + bf16(b)) with float32 accumulation - on the
 * tcgen05 tensor cores, the cast of x fused in.  weight_host: (n_out, n_in) row-major float32 (nn.Linear.weight),
 * bias_host: n_out floats or NULL.  n_in must be 128, n_out a multiple of 32 up to 768.  Synchronous. */
typedef struct adtfe_linear adtfe_linear;
int adtfe_linear_create(int32_t n_in, int32_t n_out, const float* weight_host, const float* bias_host, int device,
                        adtfe_linear** out);
int adtfe_linear_destroy(adtfe_linear* linear);
/* x_dev: (n_rows, 128) float32, contiguous, 16-byte aligned - e.g. the (rows, n_mels) matrix adtfe_logmel writes;
 * out_bf16_dev: (n_rows, n_out) bfloat16. */
int adtfe_linear_forward(const adtfe_linear* linear, const float* x_dev, int64_t n_rows, void* out_bf16_dev,
                         void* stream);

/* ---- diagnostics ----------------------------------------------------------------------- */
/* Launch trace: after adtfe_trace_begin every kernel launched by adtfe_render / adtfe_render_logmel is bracketed
 * by a pair of timing events on its stream; adtfe_trace_dump synchronises the device, writes
 * "kernel,index,start_ms,end_ms" lines (times relative to the first launch; index = chunk or group) and stops
 * tracing.  Shows how the chunks' kernels and the log-mel groups overlap across the internal streams. */
int adtfe_trace_begin(void);
int adtfe_trace_dump(const char* path);

/* ---- host planner (no GPU work) ------------------------------------------------------ */
/* C++ restatement of the per-note bookkeeping of SynthDrum.__call__ (modules/synthetiser.py:255-292:
 * RNG draws in the reference's order from Python's MT19937 stream, float32 index rules, velocity
 * curve, track order) plus the tile bucketing - what adt_str_b200/planner.py does in Python.
 * group_ptr[28]: for pitch 35+i the admitted similarity groups are entries group_ptr[i]..group_ptr[i+1]
 * of (group_first, group_count) = contiguous one-shot id ranges, best group first; gain[27] = mixing gain
 * per pitch (< 0: the reference raises KeyError); inverse_ptr/inverse_pitch: ADTOF class -> member pitches. */
typedef struct adtfe_planner adtfe_planner;
int adtfe_planner_create(int32_t sample_rate, double input_sec, double mixup_range, double use_fx_prob,
                         int32_t adtof_mapping, const int32_t* lengths, const int64_t* offsets, int32_t n_oneshots,
                         const int32_t* group_ptr,
                         const int32_t* group_first, const int32_t* group_count, const float* gain,
                         const int32_t* inverse_ptr, const int32_t* inverse_pitch, adtfe_planner** out);
int adtfe_planner_destroy(adtfe_planner* planner);
/* Probabilities of the FX chain's three plugins (SynthDrumConfig.use_reverb_prob / use_compression_prob /
 * use_limiter_prob; BoardChain.get_board, synthetiser.py:79-86).  Without this call an FX coin hit
 * (use_fx_prob > 0) draws an empty board. */
int adtfe_planner_set_fx(adtfe_planner* planner, double use_reverb_prob, double use_compression_prob,
                         double use_limiter_prob);
/* notes: float32 rows [onset, offset, pitch, velocity] of all segments back to back, counts[n_seg] rows
 * each.  mt_state[625]: random.getstate()[1], advanced in place.  ld_wav_in: 0 = derive.  Returns 0, a
 * negative adtfe_status, or 1 invalid note (ValueError) / 2 no admitted group (IndexError) / 3 KeyError
 * with info = {segment, note}.
 * out_counts = {n_events, n_groups, n_seg, tiles_per_seg, n_peak_work, n_tile_events, ld_wav, max_len, n_fx}.
 * FX: when the segment's FX coin hits (random() < use_fx_prob, synthetiser.py:154) the planner draws what
 * BoardChain.get_board draws from `random` - the three plugin coins and the reverb's four uniforms - and
 * leaves the compressor / limiter parameters, which the reference draws from torch's generator
 * (utils/utils.py:266-269), as NaN for the caller to fill in record order (adtfe_planner_export_fx).
 * mix_len / group_ptr are exported for inspection only; the device plan does not need them. */
int adtfe_planner_plan(adtfe_planner* planner, const float* notes, const int32_t* counts, int32_t n_seg,
                       uint32_t* mt_state, int64_t ld_wav_in, int64_t* out_counts, int32_t* info);
int adtfe_planner_export(const adtfe_planner* planner, adtfe_event* events, int32_t* mix_len, int32_t* group_ptr,
                         adtfe_segment* segments, int32_t* tile_ptr, adtfe_peak_item* peak_work, int32_t* tile_events);
/* The FX records of the last plan (out_counts[8] of them). */
int adtfe_planner_export_fx(const adtfe_planner* planner, adtfe_fx* fx);

/* The last plan as n_batches collated batches laid end to end (batch_sizes[b] segments each), written straight into
 * a plan blob for adtfe_frontend_host - RenderPlan.set_batches + PlanBuffers.pack of the Python host without the
 * interpreter.  Every batch keeps its own width (its longest segment: collate_fn, train_dataset.py:53) and frame
 * count max(0, 1 + width / hop - 2 * wpi - 1) (model.py:79,95-97); one render chunk per chunk_batches batches.
 * shape_out: counts, ld_wav, mel_*, n_fx, sample_rate and chunk fields filled (device pointers NULL; chunks_host =
 * chunks_out, which must hold n_batches + 1 records and stay alive while the shape is used).  fx_offset_out (may be
 * NULL): byte offset of the blob's FX section, whose NaN fields the caller fills before the blob is uploaded.  batch_width_out / batch_frames_out:
 * n_batches values each (may be NULL).  *blob_bytes_out = bytes the blob takes; with blob_host NULL or
 * blob_capacity too small nothing is written and ADTFE_ERR_WORKSPACE is returned (size query). */
int adtfe_planner_pack_batches(const adtfe_planner* planner, const int32_t* batch_sizes, int32_t n_batches,
                               int32_t chunk_batches, int32_t hop, int32_t wpi, void* blob_host, size_t blob_capacity,
                               adtfe_plan* shape_out, adtfe_chunk* chunks_out, int64_t* batch_width_out,
                               int64_t* batch_frames_out, size_t* blob_bytes_out, size_t* fx_offset_out);

#ifdef __cplusplus
}
#endif
#endif /* ADTFE_H */
