/* TEST INFRASTRUCTURE ONLY - CPU restatement of the reference's FX chain (oracle/__init__.py).
 *
 * The reference applies `pedalboard` plugins between the instrument sum and the global normalisation
 * (/root/reference/modules/synthetiser.py:30-87 BoardChain, :121-137 _add_fx, :154).  pedalboard (unpinned,
 * requirements.txt:10) is NOT installed in this image and its DSP is third-party C++ (JUCE 6/7, wrapped by
 * pedalboard's JucePlugin<>), absent from /root/reference - so this file restates the PUBLISHED algorithms of the
 * three JUCE processors pedalboard wraps, sample by sample, in float32, in the order JUCE evaluates them:
 *
 *   Reverb      juce::Reverb::processMono (juce_audio_basics/utilities/juce_Reverb.h): Freeverb - input * 0.015,
 *               8 parallel lowpass-feedback combs (tunings 1116..1617 @ 44.1 kHz scaled by (int)sr / 44100),
 *               4 series allpasses (556, 441, 341, 225), damp = damping * 0.4, feedback = room * 0.28 + 0.7,
 *               out = reverb * (0.5 * 3 * wet * (1 + width)) + in * (2 * dry); the parameter smoothers start at
 *               their targets (setParameters precedes prepare), so all gains are constants.
 *   Compressor  juce::dsp::Compressor<float>::processSample: peak BallisticsFilter (cte = exp(-2 pi 1000 / (sr t_ms)),
 *               attack when |x| > y, else release), gain = env < thr ? 1 : pow(env / thr, 1 / ratio - 1).
 *   Limiter     juce::dsp::Limiter<float>: compressor (-10 dB, 4:1, 2 ms, 200 ms) -> compressor (threshold, 1000:1,
 *               0.001 ms, release 100 ms) -> * 10^(10 * (1 - 1/4) / 40) * 10^(-threshold / 20) -> clip to [-1, 1].
 *
 * PARITY UNPINNED: there is no pedalboard here to check this restatement against (SURVEY §8c "FX: no oracle
 * available"); the GPU kernels are checked against THIS file, and the call order / RNG alignment of the reference
 * against a recording stand-in for pedalboard (tests/test_fx.py).
 *
 * JUCE_UNDENORMALISE (x += 0.1f; x -= 0.1f on Intel builds) is applied where Freeverb has it.
 *   gcc -O2 -ffp-contract=off -shared -fPIC -o oracle/_build/libfxoracle.so oracle/fx_oracle.c -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define UNDENORM(x) do { volatile float t_ = (x) + 0.1f; (x) = t_ - 0.1f; } while (0)

typedef struct { float* buf; int size, idx; float last; } comb_t;
typedef struct { float* buf; int size, idx; } allpass_t;

static const int kComb[8] = {1116, 1188, 1277, 1356, 1422, 1491, 1557, 1617};
static const int kAllpass[4] = {556, 441, 341, 225};

/* x[n] in place.  Returns 0, or -1 when out of memory. */
int fx_reverb(float* x, int64_t n, int sample_rate, float room_size, float damping, float wet_level, float dry_level,
              float width) {
    comb_t comb[8];
    allpass_t ap[4];
    const int isr = (int)sample_rate;
    for (int j = 0; j < 8; ++j) {
        comb[j].size = (isr * kComb[j]) / 44100; comb[j].idx = 0; comb[j].last = 0.0f;
        comb[j].buf = (float*)calloc((size_t)comb[j].size, sizeof(float));
        if (!comb[j].buf) return -1;
    }
    for (int j = 0; j < 4; ++j) {
        ap[j].size = (isr * kAllpass[j]) / 44100; ap[j].idx = 0;
        ap[j].buf = (float*)calloc((size_t)ap[j].size, sizeof(float));
        if (!ap[j].buf) return -1;
    }
    const float wet = wet_level * 3.0f;
    const float dry = dry_level * 2.0f;
    const float wet1 = 0.5f * wet * (1.0f + width);
    const float gain = 0.015f;
    const float damp = damping * 0.4f, feedback = room_size * 0.28f + 0.7f;
    for (int64_t i = 0; i < n; ++i) {
        const float input = x[i] * gain;
        float output = 0.0f;
        for (int j = 0; j < 8; ++j) {
            comb_t* c = &comb[j];
            const float out = c->buf[c->idx];
            c->last = (out * (1.0f - damp)) + (c->last * damp);
            UNDENORM(c->last);
            float temp = input + (c->last * feedback);
            UNDENORM(temp);
            c->buf[c->idx] = temp;
            c->idx = (c->idx + 1) % c->size;
            output += out;
        }
        for (int j = 0; j < 4; ++j) {
            allpass_t* a = &ap[j];
            const float buffered = a->buf[a->idx];
            float temp = output + (buffered * 0.5f);
            UNDENORM(temp);
            a->buf[a->idx] = temp;
            a->idx = (a->idx + 1) % a->size;
            output = buffered - output;
        }
        x[i] = output * wet1 + x[i] * dry;
    }
    for (int j = 0; j < 8; ++j) free(comb[j].buf);
    for (int j = 0; j < 4; ++j) free(ap[j].buf);
    return 0;
}

typedef struct { float threshold, threshold_inv, ratio_inv, cte_at, cte_rl, yold; } comp_t;

static float limited_cte(double exp_factor, float time_ms) {
    return time_ms < 1.0e-3f ? 0.0f : (float)exp(exp_factor / time_ms);
}

static void comp_init(comp_t* c, int sample_rate, float threshold_db, float ratio, float attack_ms, float release_ms) {
    const double exp_factor = -2.0 * 3.14159265358979323846 * 1000.0 / (double)sample_rate;
    c->threshold = threshold_db > -200.0f ? powf(10.0f, threshold_db * 0.05f) : 0.0f;
    c->threshold_inv = 1.0f / c->threshold;
    c->ratio_inv = 1.0f / ratio;
    c->cte_at = limited_cte(exp_factor, attack_ms);
    c->cte_rl = limited_cte(exp_factor, release_ms);
    c->yold = 0.0f;
}

static float comp_sample(comp_t* c, float in) {
    const float rect = fabsf(in);
    const float cte = rect > c->yold ? c->cte_at : c->cte_rl;
    const float env = rect + cte * (c->yold - rect);
    c->yold = env;
    const float g = env < c->threshold ? 1.0f : powf(env * c->threshold_inv, c->ratio_inv - 1.0f);
    return g * in;
}

void fx_compressor(float* x, int64_t n, int sample_rate, float threshold_db, float ratio, float attack_ms,
                   float release_ms) {
    comp_t c;
    comp_init(&c, sample_rate, threshold_db, ratio, attack_ms, release_ms);
    for (int64_t i = 0; i < n; ++i) x[i] = comp_sample(&c, x[i]);
}

void fx_limiter(float* x, int64_t n, int sample_rate, float threshold_db, float release_ms) {
    comp_t first, second;
    comp_init(&first, sample_rate, -10.0f, 4.0f, 2.0f, 200.0f);
    comp_init(&second, sample_rate, threshold_db, 1000.0f, 0.001f, release_ms);
    const float ratio_inv = (float)(1.0 / 4.0);
    float gain = (float)pow(10.0, 10.0 * (1.0 - ratio_inv) / 40.0);
    gain *= threshold_db < 100.0f ? powf(10.0f, -threshold_db * 0.05f) : 0.0f;   /* decibelsToGain(-thr, -100) */
    /* stage by stage over the block, as juce::dsp::Limiter::process does */
    for (int64_t i = 0; i < n; ++i) x[i] = comp_sample(&first, x[i]);
    for (int64_t i = 0; i < n; ++i) x[i] = comp_sample(&second, x[i]);
    for (int64_t i = 0; i < n; ++i) {
        float v = x[i] * gain;
        v = v < -1.0f ? -1.0f : (v > 1.0f ? 1.0f : v);   /* FloatVectorOperations::clip keeps NaN */
        x[i] = v;
    }
}
