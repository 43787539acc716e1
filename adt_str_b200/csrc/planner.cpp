// Host planner in C++: notes + Python's random stream -> event records, tile buckets, peak work.
//
// Same semantics as adt_str_b200/planner.py (which stays the readable restatement and the
// fallback for float64 note arrays); this one exists because at >1e6 audio-s/s the Python
// per-note loop is the bottleneck (SURVEY §7 "host planner throughput").
//
// The reference draws from Python's `random` module (modules/synthetiser.py:194-199, 217, 154).
// To stay on the *same stream* the caller hands over `random.getstate()` (624 MT19937 words +
// index) and gets the advanced state back for `random.setstate()`.  The generator, the 53-bit
// `random()`, `getrandbits(k)`, `_randbelow` (rejection on k = bit_length(n) bits), `choice` and
// `uniform` below follow CPython 3.12 (Modules/_randommodule.c, Lib/random.py).
//
// Index rules are float32 like `torch.tensor(notes)` arithmetic (synthetiser.py:229, 262, 243).
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "../../include/adtfe.h"

namespace adtfe {
void set_error(const char* fmt, ...);
}

namespace {

struct MT {
    uint32_t s[624];
    int pos;
    uint32_t next() {
        if (pos >= 624) {
            static const uint32_t mag01[2] = {0u, 0x9908b0dfu};
            int kk;
            uint32_t y;
            for (kk = 0; kk < 624 - 397; kk++) {
                y = (s[kk] & 0x80000000u) | (s[kk + 1] & 0x7fffffffu);
                s[kk] = s[kk + 397] ^ (y >> 1) ^ mag01[y & 1u];
            }
            for (; kk < 623; kk++) {
                y = (s[kk] & 0x80000000u) | (s[kk + 1] & 0x7fffffffu);
                s[kk] = s[kk + (397 - 624)] ^ (y >> 1) ^ mag01[y & 1u];
            }
            y = (s[623] & 0x80000000u) | (s[0] & 0x7fffffffu);
            s[623] = s[396] ^ (y >> 1) ^ mag01[y & 1u];
            pos = 0;
        }
        uint32_t y = s[pos++];
        y ^= (y >> 11);
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= (y >> 18);
        return y;
    }
    double random() {  // genrand_res53
        const uint32_t a = next() >> 5, b = next() >> 6;
        return (a * 67108864.0 + b) * (1.0 / 9007199254740992.0);
    }
    uint32_t below(uint32_t n) {  // Random._randbelow_with_getrandbits, n >= 1
        int k = 0;
        for (uint32_t t = n; t; t >>= 1) ++k;  // n.bit_length()
        uint32_t r = next() >> (32 - k);
        while (r >= n) r = next() >> (32 - k);
        return r;
    }
};

struct GroupRange {
    int32_t first, count;
};

}  // namespace

struct adtfe_planner {
    int32_t sample_rate = 24000;
    double input_sec = 2.56, mixup_range = 0.8, use_fx_prob = 0.0;
    double use_reverb_prob = 0.0, use_compression_prob = 0.0, use_limiter_prob = 0.0;   // adtfe_planner_set_fx
    int adtof = 0;
    std::vector<int32_t> lengths;                 // one-shot lengths
    std::vector<int64_t> offsets;                 // one-shot starts in the bank (floats)
    std::vector<GroupRange> groups[27];           // pitch 35..61 -> admitted (first id, count), best group first
    float gain[27];                               // per-pitch mixing gain, < 0 = KeyError in the reference
    std::vector<int32_t> inverse[27];             // ADTOF class -> member pitches (empty = KeyError)
    // last plan
    std::vector<adtfe_event> events;
    std::vector<int32_t> mix_len, group_ptr, tile_ptr, tile_events;
    std::vector<adtfe_peak_item> peak_work;
    std::vector<adtfe_segment> segments;
    std::vector<adtfe_fx> fx;                     // one record per segment whose FX coin hit, ascending seg
    int64_t ld_wav = 0;
    int32_t tiles_per_seg = 0;
};

extern "C" int adtfe_planner_create(int32_t sample_rate, double input_sec, double mixup_range, double use_fx_prob,
                                    int32_t adtof_mapping, const int32_t* lengths, const int64_t* offsets,
                                    int32_t n_oneshots,
                                    const int32_t* group_ptr /*28*/, const int32_t* group_first,
                                    const int32_t* group_count, const float* gain /*27*/,
                                    const int32_t* inverse_ptr /*28*/, const int32_t* inverse_pitch,
                                    adtfe_planner** out) {
    if (!out || !lengths || !offsets || !group_ptr || !gain || !inverse_ptr || n_oneshots < 0) {
        adtfe::set_error("adtfe_planner_create: bad argument");
        return ADTFE_ERR_BAD_ARG;
    }
    adtfe_planner* p = new adtfe_planner();
    p->sample_rate = sample_rate; p->input_sec = input_sec; p->mixup_range = mixup_range;
    p->use_fx_prob = use_fx_prob; p->adtof = adtof_mapping;
    p->lengths.assign(lengths, lengths + n_oneshots);
    p->offsets.assign(offsets, offsets + n_oneshots);
    for (int i = 0; i < 27; ++i) {
        for (int j = group_ptr[i]; j < group_ptr[i + 1]; ++j) p->groups[i].push_back({group_first[j], group_count[j]});
        p->gain[i] = gain[i];
        for (int j = inverse_ptr[i]; j < inverse_ptr[i + 1]; ++j) p->inverse[i].push_back(inverse_pitch[j]);
    }
    *out = p;
    return ADTFE_OK;
}

extern "C" int adtfe_planner_destroy(adtfe_planner* p) {
    delete p;
    return ADTFE_OK;
}

extern "C" int adtfe_planner_set_fx(adtfe_planner* p, double use_reverb_prob, double use_compression_prob,
                                    double use_limiter_prob) {
    if (!p) {
        adtfe::set_error("adtfe_planner_set_fx: null planner");
        return ADTFE_ERR_BAD_ARG;
    }
    p->use_reverb_prob = use_reverb_prob; p->use_compression_prob = use_compression_prob;
    p->use_limiter_prob = use_limiter_prob;
    return ADTFE_OK;
}

static float vel_to_vol(float v) {  // synthetiser.py:204-212, float32 throughout
    if (v == 0.0f) return 0.0f;
    const float c = v < 0.0f ? 0.0f : (v > 127.0f ? 127.0f : v);
    const float x = c / 127.0f;
    const float pw = powf(6.0f, x);
    const float t = 0.9f * (pw - 1.0f);
    return 0.1f + t / 5.0f;
}

// status: 0 ok; 1 ValueError (invalid note); 2 IndexError (no admitted group); 3 KeyError;
// info[0] = segment, info[1] = note index of the failure.
extern "C" int adtfe_planner_plan(adtfe_planner* P, const float* notes, const int32_t* counts, int32_t n_seg,
                                  uint32_t* mt_state /*625, updated*/, int64_t ld_wav_in, int64_t* out_counts /*9*/,
                                  int32_t* info /*2*/) {
    if (!P || !counts || !mt_state || !out_counts || !info || n_seg < 0) {
        adtfe::set_error("adtfe_planner_plan: bad argument");
        return ADTFE_ERR_BAD_ARG;
    }
    MT rng;
    memcpy(rng.s, mt_state, 624 * 4);
    rng.pos = (int)mt_state[624];
    const float sr_f = (float)P->sample_rate;
    P->events.clear(); P->mix_len.clear(); P->segments.clear(); P->fx.clear();
    P->group_ptr.assign(1, 0);
    int status = 0;
    info[0] = info[1] = -1;

    std::vector<adtfe_event> ev;
    std::vector<int32_t> ml, rank_of;
    std::vector<int> order;
    const float* row = notes;
    int64_t max_len = 0;
    for (int32_t s = 0; s < n_seg && status == 0; ++s) {
        const int32_t n = counts[s];
        adtfe_segment seg;
        seg.first_event = (int32_t)P->events.size();
        if (n == 0) {  // synthetiser.py:257-258
            seg.len = (int32_t)(P->input_sec * P->sample_rate);
            seg.flags = 0; seg.max_volume = 0.0f;
            P->segments.push_back(seg);
            max_len = std::max<int64_t>(max_len, seg.len);
            continue;
        }
        float max_off = row[1], max_vel = 0.0f;
        for (int32_t i = 0; i < n; ++i) {
            max_off = std::max(max_off, row[4 * i + 1]);
            max_vel = std::max(max_vel, row[4 * i + 3]);
        }
        const float end = max_off + 0.1f;  // synthetiser.py:262
        const int32_t wave_length = end < (float)P->input_sec ? (int32_t)(P->input_sec * P->sample_rate)
                                                               : (int32_t)(end * sr_f);
        int32_t main_of[27], sub_of[27], rank[27], n_rank = 0;
        for (int i = 0; i < 27; ++i) rank[i] = -1;
        ev.assign(n, adtfe_event());
        ml.assign(n, 0);
        rank_of.assign(n, 0);
        for (int32_t i = 0; i < n && status == 0; ++i) {
            const float on = row[4 * i], off = row[4 * i + 1], pitch = row[4 * i + 2], vel = row[4 * i + 3];
            if (!(pitch >= 35.0f && pitch <= 61.0f && off >= on) || on < 0.0f) { status = 1; info[0] = s; info[1] = i; break; }
            const int inst = (int)pitch;
            if ((float)inst != pitch) { status = 3; info[0] = s; info[1] = i; break; }
            const int pi = inst - 35;
            if (rank[pi] < 0) {
                int32_t chosen[2];
                for (int c = 0; c < 2 && status == 0; ++c) {  // synthetiser.py:192-202, main then sub
                    int gp = pi;
                    if (P->adtof) {
                        if (P->inverse[pi].empty()) { status = 3; break; }
                        gp = P->inverse[pi][rng.below((uint32_t)P->inverse[pi].size())] - 35;
                    }
                    const std::vector<GroupRange>& g = P->groups[gp];
                    if (g.empty()) { status = 2; break; }
                    const GroupRange& r = g[rng.below((uint32_t)g.size())];
                    chosen[c] = r.first + (int32_t)rng.below((uint32_t)r.count);
                }
                if (status) { info[0] = s; info[1] = i; break; }
                main_of[pi] = chosen[0]; sub_of[pi] = chosen[1];
                rank[pi] = n_rank++;
            }
            const double alpha = 0.0 + (P->mixup_range - 0.0) * rng.random();  // random.uniform(0, mixup_range)
            adtfe_event& e = ev[i];
            e.main_id = main_of[pi]; e.sub_id = sub_of[pi];
            e.ca = (float)(1.0 - alpha); e.cb = (float)alpha;
            e.start = (int32_t)(on * sr_f);
            const int32_t mix = std::max(P->lengths[e.main_id], P->lengths[e.sub_id]);
            e.len = std::max(0, std::min(mix, wave_length - e.start));
            e.gain = vel_to_vol(vel);
            e.seg = s;
            ml[i] = mix;
            rank_of[i] = rank[pi];
        }
        if (status) break;
        for (int pi = 0; pi < 27; ++pi)  // gain lookup happens in instrument_mixer, after the loop
            if (rank[pi] >= 0 && P->gain[pi] < 0.0f) { status = 3; info[0] = s; info[1] = -1; }
        if (status) break;
        if (rng.random() < P->use_fx_prob) {  // synthetiser.py:154 -> _add_fx -> BoardChain.get_board (:79-86)
            const float nan = nanf("");
            adtfe_fx f;
            f.seg = s; f.flags = 0;
            f.room_size = f.damping = f.wet_level = f.dry_level = f.width = 0.0f;
            f.comp_threshold_db = f.comp_ratio = f.comp_attack_ms = f.comp_release_ms = f.lim_threshold_db = 0.0f;
            if (rng.random() < P->use_reverb_prob) {  // _add_reverb (:44-61): four random.uniform(a, b) = a + (b - a) * random()
                const double room = 0.2 + (0.8 - 0.2) * rng.random();
                const double damping = 0.2 + (0.8 - 0.2) * rng.random();
                const double wet = 0.1 + (0.4 - 0.1) * rng.random();
                const double width = 0.6 + (1.0 - 0.6) * rng.random();
                f.flags |= ADTFE_FX_REVERB;
                f.room_size = (float)room; f.damping = (float)damping; f.wet_level = (float)wet;
                f.dry_level = (float)(1.0 - wet); f.width = (float)width;
            }
            if (rng.random() < P->use_compression_prob) {  // _add_compression (:63-75): four draws from torch's generator
                f.flags |= ADTFE_FX_COMPRESSOR;
                f.comp_threshold_db = f.comp_ratio = f.comp_attack_ms = f.comp_release_ms = nan;
            }
            if (rng.random() < P->use_limiter_prob) {  // _add_limiter (:77-79): one draw from torch's generator
                f.flags |= ADTFE_FX_LIMITER;
                f.lim_threshold_db = nan;
            }
            P->fx.push_back(f);
        }
        for (int32_t i = 0; i < n; ++i) ev[i].gain *= P->gain[(int)row[4 * i + 2] - 35];
        // track order: instrument by first appearance, then note order
        order.resize(n);
        for (int i = 0; i < n; ++i) order[i] = i;
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return rank_of[a] < rank_of[b]; });
        for (int k = 0; k < n; ++k) {
            const int i = order[k];
            if (k > 0 && rank_of[i] != rank_of[order[k - 1]]) P->group_ptr.push_back((int32_t)P->events.size());
            P->events.push_back(ev[i]);
            P->mix_len.push_back(ml[i]);
        }
        P->group_ptr.push_back((int32_t)P->events.size());
        seg.len = wave_length; seg.flags = 1; seg.max_volume = vel_to_vol(std::max(0.0f, max_vel));
        P->segments.push_back(seg);
        max_len = std::max<int64_t>(max_len, wave_length);
        row += 4 * (int64_t)n;
    }
    memcpy(mt_state, rng.s, 624 * 4);
    mt_state[624] = (uint32_t)rng.pos;
    if (status) return status;

    // ---- geometry, tile CSR, peak work list
    int64_t ld = ld_wav_in > 0 ? ld_wav_in : ((max_len + 3) / 4) * 4;
    if (ld < max_len || ld % 4) {
        adtfe::set_error("adtfe_planner_plan: ld_wav must be a multiple of 4 and cover the longest segment");
        return ADTFE_ERR_BAD_ARG;
    }
    P->ld_wav = ld;
    P->tiles_per_seg = (int32_t)((ld + ADTFE_TILE - 1) / ADTFE_TILE);
    const int64_t n_tiles = (int64_t)n_seg * P->tiles_per_seg;
    P->tile_ptr.assign(n_tiles + 1, 0);
    for (const adtfe_event& e : P->events)
        if (e.len > 0)
            for (int t = e.start / ADTFE_TILE; t <= (e.start + e.len - 1) / ADTFE_TILE; ++t)
                ++P->tile_ptr[(int64_t)e.seg * P->tiles_per_seg + t + 1];
    for (int64_t t = 0; t < n_tiles; ++t) P->tile_ptr[t + 1] += P->tile_ptr[t];
    P->tile_events.assign(P->tile_ptr[n_tiles], 0);
    {
        std::vector<int32_t> cursor(P->tile_ptr.begin(), P->tile_ptr.end() - 1);
        for (int32_t i = 0; i < (int32_t)P->events.size(); ++i) {
            const adtfe_event& e = P->events[i];
            if (e.len > 0)
                for (int t = e.start / ADTFE_TILE; t <= (e.start + e.len - 1) / ADTFE_TILE; ++t)
                    P->tile_events[cursor[(int64_t)e.seg * P->tiles_per_seg + t]++] = i;  // ascending event id
        }
    }
    P->peak_work.clear();
    for (int32_t g = 0; g + 1 < (int32_t)P->group_ptr.size(); ++g) {
        const int32_t e0 = P->group_ptr[g];
        const adtfe_event& head = P->events[e0];
        adtfe_peak_item it;
        it.a_off = P->offsets[head.main_id]; it.b_off = P->offsets[head.sub_id];
        it.la = P->lengths[head.main_id]; it.lb = P->lengths[head.sub_id];
        it.mix_len = P->mix_len[e0];
        it.chunk = 0;
        // one item (= one warp of the peak pass) per ADTFE_PEAK_NOTES notes of the group: a long-form render has a few
        // groups of hundreds of notes each, which one warp per group would walk one after the other
        for (int32_t e = e0; e < P->group_ptr[g + 1]; e += ADTFE_PEAK_NOTES) {
            it.first_event = e;
            it.n_events = std::min<int32_t>(ADTFE_PEAK_NOTES, P->group_ptr[g + 1] - e);
            P->peak_work.push_back(it);
        }
    }
    out_counts[0] = (int64_t)P->events.size();
    out_counts[1] = (int64_t)P->group_ptr.size() - 1;
    out_counts[2] = n_seg;
    out_counts[3] = P->tiles_per_seg;
    out_counts[4] = (int64_t)P->peak_work.size();
    out_counts[5] = (int64_t)P->tile_events.size();
    out_counts[6] = ld;
    out_counts[7] = max_len;
    out_counts[8] = (int64_t)P->fx.size();
    return ADTFE_OK;
}

// Copies the last plan into caller arrays (any may be NULL to skip).
extern "C" int adtfe_planner_export(const adtfe_planner* P, adtfe_event* events, int32_t* mix_len, int32_t* group_ptr,
                                    adtfe_segment* segments, int32_t* tile_ptr, adtfe_peak_item* peak_work,
                                    int32_t* tile_events) {
    if (!P) return ADTFE_ERR_BAD_ARG;
    if (events && !P->events.empty()) memcpy(events, P->events.data(), P->events.size() * sizeof(adtfe_event));
    if (mix_len && !P->mix_len.empty()) memcpy(mix_len, P->mix_len.data(), P->mix_len.size() * 4);
    if (group_ptr) memcpy(group_ptr, P->group_ptr.data(), P->group_ptr.size() * 4);
    if (segments && !P->segments.empty()) memcpy(segments, P->segments.data(), P->segments.size() * sizeof(adtfe_segment));
    if (tile_ptr) memcpy(tile_ptr, P->tile_ptr.data(), P->tile_ptr.size() * 4);
    if (peak_work && !P->peak_work.empty())
        memcpy(peak_work, P->peak_work.data(), P->peak_work.size() * sizeof(adtfe_peak_item));
    if (tile_events && !P->tile_events.empty()) memcpy(tile_events, P->tile_events.data(), P->tile_events.size() * 4);
    return ADTFE_OK;
}

extern "C" int adtfe_planner_export_fx(const adtfe_planner* P, adtfe_fx* fx) {
    if (!P) return ADTFE_ERR_BAD_ARG;
    if (fx && !P->fx.empty()) memcpy(fx, P->fx.data(), P->fx.size() * sizeof(adtfe_fx));
    return ADTFE_OK;
}

// The last plan as `n_batches` collated batches laid end to end, written straight into a plan blob (the layout of
// adtfe_plan_blob_layout) - what RenderPlan.set_batches + PlanBuffers.pack do in Python, without the interpreter:
// every batch keeps its own width (its longest segment, train_dataset.py:53) and therefore its own frame count
// max(0, 1 + width / hop - 2 * wpi - 1) (model.py:79,95-97); one render chunk per `chunk_batches` batches.
extern "C" int adtfe_planner_pack_batches(const adtfe_planner* P, const int32_t* batch_sizes, int32_t n_batches,
                                          int32_t chunk_batches, int32_t hop, int32_t wpi, void* blob_host,
                                          size_t blob_capacity, adtfe_plan* shape_out, adtfe_chunk* chunks_out,
                                          int64_t* batch_width_out, int64_t* batch_frames_out, size_t* blob_bytes_out,
                                          size_t* fx_offset_out) {
    if (!P || !batch_sizes || n_batches <= 0 || hop <= 0 || wpi < 0 || !shape_out || !chunks_out || !blob_bytes_out) {
        adtfe::set_error("adtfe_planner_pack_batches: bad argument");
        return ADTFE_ERR_BAD_ARG;
    }
    const int32_t n_seg = (int32_t)P->segments.size();
    int64_t total = 0;
    for (int32_t b = 0; b < n_batches; ++b) {
        if (batch_sizes[b] <= 0) {
            adtfe::set_error("adtfe_planner_pack_batches: batch %d is empty", b);
            return ADTFE_ERR_BAD_ARG;
        }
        total += batch_sizes[b];
    }
    if (total != n_seg) {
        adtfe::set_error("adtfe_planner_pack_batches: batch sizes add up to %lld, the plan has %d segments",
                         (long long)total, n_seg);
        return ADTFE_ERR_BAD_ARG;
    }
    const int32_t step = std::max(1, chunk_batches);
    std::vector<adtfe_mel_row> rows((size_t)n_seg);
    int64_t row0 = 0;
    int32_t s0 = 0, n_chunks = 0, max_count = 0;
    // peak work items are ordered by first_event (FX records by seg), so a chunk's first one is found by walking forward
    size_t pw = 0, fxr = 0;
    for (int32_t b = 0; b < n_batches; ++b) {
        const int32_t s1 = s0 + batch_sizes[b];
        int64_t width = 0;
        for (int32_t s = s0; s < s1; ++s) width = std::max<int64_t>(width, P->segments[s].len);
        const int64_t frames = std::max<int64_t>(0, 1 + width / hop - 2 * (int64_t)wpi - 1);
        for (int32_t s = s0; s < s1; ++s) {
            rows[s].out_row = row0 + (int64_t)(s - s0) * frames;
            rows[s].count = (int32_t)frames;
            rows[s].flags = P->segments[s].flags == 0 ? ADTFE_MEL_ROW_SILENT : 0;  // an empty segment is all zeros
        }
        if (batch_width_out) batch_width_out[b] = width;
        if (batch_frames_out) batch_frames_out[b] = frames;
        max_count = std::max<int32_t>(max_count, (int32_t)frames);
        if (b % step == 0) {
            const int32_t ev = P->segments[s0].first_event;
            while (pw < P->peak_work.size() && P->peak_work[pw].first_event < ev) ++pw;
            while (fxr < P->fx.size() && P->fx[fxr].seg < s0) ++fxr;
            chunks_out[n_chunks].seg = s0;
            chunks_out[n_chunks].event = ev;
            chunks_out[n_chunks].peak_work = (int32_t)pw;
            chunks_out[n_chunks].fx_row = (int32_t)fxr;
            ++n_chunks;
        }
        row0 += (int64_t)batch_sizes[b] * frames;
        s0 = s1;
    }
    chunks_out[n_chunks].seg = n_seg;
    chunks_out[n_chunks].event = (int32_t)P->events.size();
    chunks_out[n_chunks].peak_work = (int32_t)P->peak_work.size();
    chunks_out[n_chunks].fx_row = (int32_t)P->fx.size();

    adtfe_plan shape;
    memset(&shape, 0, sizeof(shape));
    shape.n_events = (int32_t)P->events.size();
    shape.n_seg = n_seg;
    shape.tiles_per_seg = P->tiles_per_seg;
    shape.n_peak_work = (int32_t)P->peak_work.size();
    shape.ld_wav = P->ld_wav;
    shape.mel_total_rows = row0;
    shape.mel_max_count = max_count;
    shape.n_chunks = n_chunks;
    shape.chunks_host = chunks_out;
    shape.n_tile_events = (int32_t)P->tile_events.size();
    shape.n_fx = (int32_t)P->fx.size();
    shape.sample_rate = P->sample_rate;
    size_t off[7], fixed = 0;
    int rc = adtfe_plan_blob_layout(&shape, off, &fixed);
    if (rc != ADTFE_OK) return rc;
    // mel_total_rows == 0 (every batch too short for a frame) drops the rows section: the ragged form needs rows
    const size_t need = (fixed + 4 * P->tile_events.size() + 15) & ~(size_t)15;
    *blob_bytes_out = need;
    *shape_out = shape;
    if (fx_offset_out) *fx_offset_out = off[5];
    if (!blob_host || blob_capacity < need) {
        adtfe::set_error("adtfe_planner_pack_batches: blob of %zu B needed, %zu B given", need, blob_capacity);
        return ADTFE_ERR_WORKSPACE;
    }
    char* h = (char*)blob_host;
    if (!P->events.empty()) memcpy(h + off[0], P->events.data(), P->events.size() * sizeof(adtfe_event));
    if (n_seg) memcpy(h + off[1], P->segments.data(), (size_t)n_seg * sizeof(adtfe_segment));
    memcpy(h + off[2], P->tile_ptr.data(), P->tile_ptr.size() * 4);
    if (!P->peak_work.empty()) memcpy(h + off[3], P->peak_work.data(), P->peak_work.size() * sizeof(adtfe_peak_item));
    if (row0 > 0) memcpy(h + off[4], rows.data(), rows.size() * sizeof(adtfe_mel_row));
    if (!P->fx.empty()) memcpy(h + off[5], P->fx.data(), P->fx.size() * sizeof(adtfe_fx));
    if (!P->tile_events.empty()) memcpy(h + off[6], P->tile_events.data(), P->tile_events.size() * 4);
    return ADTFE_OK;
}
