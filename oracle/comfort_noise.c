/* TEST INFRASTRUCTURE ONLY (see lyra_oracle.h).
 *
 * CPU restatement of the reference's comfort-noise generator:
 *   ComfortNoiseGenerator::Create          lyra/comfort_noise_generator.cc:37-62
 *   RunConditioning = FftFromFeatures + InvertFft   lyra/comfort_noise_generator.cc:74-119
 *   RunModel (slice of the reconstructed hop)       lyra/comfort_noise_generator.cc:79-84
 *   GenerativeModel FIFO / partial hops             lyra/generative_model_interface.h:45-134 (lo_gen_* below)
 *
 * The arithmetic lives in the un-vendored dependency com_google_audio_dsp = mchinen/multichannel-audio-tools@14a45c5
 * (reference WORKSPACE:68-78): audio_dsp::MelFilterbank::EstimateInverse and audio_dsp::InverseSpectrogram.  Neither
 * source is available offline and the reference holds NO golden vector for them (its generator draws random phases from an
 * unseeded absl::BitGen, comfort_noise_generator.cc:103), so this file restates their published *algorithms* with two
 * explicit design choices, and is pinned by the one criterion the reference itself tests
 * (comfort_noise_generator_test.cc:100-138: log-mel of the generated noise within LSD < 0.7 of the log-mel that
 * conditioned it, after 10 hops) - "PARITY UNPINNED" beyond that statistical bound:
 *   1. mel inverse: every spectrum bin takes the triangular-weighted mix of its two mel channels, each channel first
 *      divided by the sum of its filter's weights (so a flat spectrum maps to itself), squared like the forward path's
 *      input (the filterbank works on magnitudes, the spectrogram on squared magnitudes);
 *   2. inverse spectrogram: exact inverse real FFT of the fft_length-point spectrum, weighted by a periodic Hann window
 *      of fft_length points and overlap-added at the hop; the window carries the constant that makes white noise keep
 *      its power through (reference extractor: Hann(window) analysis -> this synthesis): sqrt(N * hop / (sum wa^2 * sum ws^2)).
 *      (Measured against the reference's criterion: LSD 0.49 +- 0.03; gains between 1.5 and 3.4 times the plain Hann
 *      overlap-add pass it, 1.0 does not.)
 * Random phases: the reference is non-deterministic.  Here bin i of hop h of a generator seeded with `seed` gets phase
 * 2 pi p / 1024 with p = the top 10 bits of splitmix64(seed ^ h * 0xD1B54A32D192ED03 ^ i * 0x9E3779B97F4A7C15): counter
 * based, so the CUDA kernel reproduces it bit for bit, and quantised to the FFT's own twiddle angles so no sin / cos is
 * evaluated at run time.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "lyra_oracle.h"

/* ------------------------------------------------------------ GenerativeModel base (generative_model_interface.h:45-134) ---- */

void lo_gen_init(lo_gen* g, int num_samples_per_hop, int num_features, lo_gen_conditioning_fn fn, void* self) {
  memset(g, 0, sizeof(*g));
  g->hop = num_samples_per_hop;
  g->nf = num_features;
  g->run_conditioning = fn;
  g->self = self;
  g->cap = 8;
  g->queue = (float*)malloc(sizeof(float) * (size_t)g->cap * (size_t)num_features);
  g->hop_samples = (int16_t*)calloc((size_t)num_samples_per_hop, sizeof(int16_t));
}

void lo_gen_free(lo_gen* g) { free(g->queue); free(g->hop_samples); g->queue = NULL; g->hop_samples = NULL; }

int lo_gen_add_features(lo_gen* g, const float* features, int n) {
  g->calls_add++;
  if (n != g->nf) return -1;                                              /* :51-55 */
  if (g->count == g->cap) {
    float* q = (float*)malloc(sizeof(float) * (size_t)g->cap * 2 * (size_t)g->nf);
    for (int i = 0; i < g->count; ++i) memcpy(q + (size_t)i * g->nf, g->queue + (size_t)((g->head + i) % g->cap) * g->nf, sizeof(float) * (size_t)g->nf);
    free(g->queue);
    g->queue = q; g->head = 0; g->cap *= 2;
  }
  memcpy(g->queue + (size_t)((g->head + g->count) % g->cap) * g->nf, features, sizeof(float) * (size_t)g->nf);
  g->count++;
  return 0;
}

int lo_gen_num_samples_available(const lo_gen* g) { return g->count * g->hop - g->next_sample_in_hop; }   /* :101-103 */

int lo_gen_generate_samples(lo_gen* g, int num_samples, int16_t* out) {
  g->calls_generate++;
  g->last_generate = num_samples;
  if (num_samples < 0) return -1;                                         /* :62-65 */
  if (num_samples == 0) return 0;                                         /* :67-69 */
  if (lo_gen_num_samples_available(g) == 0) return -1;                    /* :70-74 */
  if (g->next_sample_in_hop == 0) {                                       /* :75-79 */
    if (g->run_conditioning(g->self, g->queue + (size_t)g->head * g->nf, g->hop_samples) != 0) return -1;
  }
  if (num_samples > g->hop - g->next_sample_in_hop) return -1;            /* :80-87 */
  memcpy(out, g->hop_samples + g->next_sample_in_hop, sizeof(int16_t) * (size_t)num_samples);   /* RunModel */
  g->next_sample_in_hop += num_samples;
  if (g->next_sample_in_hop == g->hop) {                                  /* :92-95 */
    g->next_sample_in_hop = 0;
    g->head = (g->head + 1) % g->cap;
    g->count--;
  }
  return num_samples;
}

/* ------------------------------------------------------------ comfort noise ---- */

struct lo_cng {
  int sample_rate, hop, window, nmel, fft, bins;
  int start_index, end_index;
  double* weights;      /* [bins] triangular weight of the bin in its lower channel */
  int* band;            /* [bins] lower channel (-1: below the first centre, -2: outside the limits) */
  double* norm;         /* [nmel] sum of the filter's weights */
  double* synth;        /* [fft] synthesis window incl. the power-preserving constant */
  double* work;         /* [fft] overlap-add buffer (InverseSpectrogram's working output) */
  double* re; double* im;
  double* sq;           /* [bins] squared-magnitude estimate */
  uint64_t seed, hops;
  lo_gen gen;
};

static double freq_to_mel(double f) { return 1127.0 * log1p(f / 700.0); }

uint32_t lo_cng_phase_index(uint64_t seed, uint64_t hop, int bin) {
  uint64_t x = seed ^ (hop * 0xD1B54A32D192ED03ull) ^ ((uint64_t)bin * 0x9E3779B97F4A7C15ull);
  x += 0x9E3779B97F4A7C15ull;                                             /* splitmix64 */
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  x ^= x >> 31;
  return (uint32_t)(x >> 54);
}

/* in-place radix-2 decimation-in-time transform with the conjugates of the forward twiddles (unscaled inverse DFT);
 * the twiddle expressions are those of logmel.c's forward transform, so the CUDA kernel's table (built from the same
 * expressions) gives the same bits */
static void ifft_inplace(double* re, double* im, int n) {
  for (int i = 1, j = 0; i < n; ++i) {
    int bit = n >> 1;
    for (; j & bit; bit >>= 1) j ^= bit;
    j ^= bit;
    if (i < j) { double t = re[i]; re[i] = re[j]; re[j] = t; t = im[i]; im[i] = im[j]; im[j] = t; }
  }
  for (int len = 2; len <= n; len <<= 1) {
    const double ang = -2.0 * M_PI / (double)len;
    for (int i = 0; i < n; i += len)
      for (int k = 0; k < len / 2; ++k) {
        const double wr = cos(ang * k), wi = -sin(ang * k);
        const int a = i + k, b = i + k + len / 2;
        const double xr = re[b] * wr - im[b] * wi, xi = re[b] * wi + im[b] * wr;
        re[b] = re[a] - xr; im[b] = im[a] - xi;
        re[a] += xr; im[a] += xi;
      }
  }
}

static int cng_conditioning(void* self, const float* features, int16_t* hop_out);

lo_cng* lo_cng_create(int sample_rate_hz, int hop, int window, int num_mel_bins, uint64_t seed) {
  if (window < hop || hop <= 0 || num_mel_bins <= 0) return NULL;
  lo_cng* c = (lo_cng*)calloc(1, sizeof(*c));
  c->sample_rate = sample_rate_hz; c->hop = hop; c->window = window; c->nmel = num_mel_bins; c->seed = seed;
  c->fft = 1;
  while (c->fft < window) c->fft <<= 1;                       /* NextPowerOfTwo(window) (.cc:40-41) */
  c->bins = c->fft / 2 + 1;
  c->weights = (double*)calloc((size_t)c->bins, sizeof(double));
  c->band = (int*)malloc(sizeof(int) * (size_t)c->bins);
  c->norm = (double*)calloc((size_t)num_mel_bins, sizeof(double));
  c->synth = (double*)malloc(sizeof(double) * (size_t)c->fft);
  c->work = (double*)calloc((size_t)c->fft, sizeof(double));
  c->re = (double*)malloc(sizeof(double) * (size_t)c->fft);
  c->im = (double*)malloc(sizeof(double) * (size_t)c->fft);
  c->sq = (double*)calloc((size_t)c->bins, sizeof(double));
  /* MelFilterbank::Initialize(bins, fs, nmel, 0, 0.495 fs) - the tables of logmel.c */
  const double lower = 0.0, upper = 0.495 * sample_rate_hz;
  const double mel_low = freq_to_mel(lower), mel_hi = freq_to_mel(upper);
  const double mel_spacing = (mel_hi - mel_low) / (double)(num_mel_bins + 1);
  double* center = (double*)malloc(sizeof(double) * (size_t)(num_mel_bins + 1));
  for (int i = 0; i < num_mel_bins + 1; ++i) center[i] = mel_low + mel_spacing * (i + 1);
  const double hz_per_sbin = 0.5 * sample_rate_hz / (double)(c->bins - 1);
  c->start_index = (int)(1.5 + lower / hz_per_sbin);
  c->end_index = (int)(upper / hz_per_sbin);
  int channel = 0;
  for (int i = 0; i < c->bins; ++i) {
    const double melf = freq_to_mel(i * hz_per_sbin);
    if (i < c->start_index || i > c->end_index) { c->band[i] = -2; continue; }
    while (channel < num_mel_bins && center[channel] < melf) ++channel;
    c->band[i] = channel - 1;
    const int ch = channel - 1;
    c->weights[i] = ch >= 0 ? (center[ch + 1] - melf) / (center[ch + 1] - center[ch]) : (center[0] - melf) / (center[0] - mel_low);
  }
  free(center);
  for (int i = c->start_index; i <= c->end_index; ++i) {
    const int ch = c->band[i];
    if (ch >= 0) c->norm[ch] += c->weights[i];
    if (ch + 1 < num_mel_bins) c->norm[ch + 1] += 1.0 - c->weights[i];
  }
  /* synthesis window */
  double swa = 0.0, sws = 0.0;
  for (int i = 0; i < window; ++i) { const double w = 0.5 - 0.5 * cos(2.0 * M_PI * i / (double)window); swa += w * w; }
  for (int i = 0; i < c->fft; ++i) { c->synth[i] = 0.5 - 0.5 * cos(2.0 * M_PI * i / (double)c->fft); sws += c->synth[i] * c->synth[i]; }
  const double gain = sqrt((double)c->fft * (double)hop / (swa * sws));
  for (int i = 0; i < c->fft; ++i) c->synth[i] *= gain;
  lo_gen_init(&c->gen, hop, num_mel_bins, cng_conditioning, c);
  return c;
}

void lo_cng_free(lo_cng* c) {
  if (!c) return;
  lo_gen_free(&c->gen);
  free(c->weights); free(c->band); free(c->norm); free(c->synth); free(c->work); free(c->re); free(c->im); free(c->sq);
  free(c);
}

lo_gen* lo_cng_gen(lo_cng* c) { return &c->gen; }
double lo_cng_synthesis_gain(const lo_cng* c) { return c->synth[c->fft / 2]; }     /* the Hann window is 1 at its centre */
const double* lo_cng_norm(const lo_cng* c) { return c->norm; }
const double* lo_cng_synth(const lo_cng* c) { return c->synth; }

/* one hop from explicit phase indices (phase[bins], 0..1023); phase == NULL: the generator's own counter-based draw */
int lo_cng_condition(lo_cng* c, const float* log_mel_features, const uint32_t* phase, int16_t* hop_out) {
  const int N = c->fft;
  /* FftFromFeatures (.cc:87-96): mel = exp(feature * 10) in float, widened; then the mel inverse (squared magnitudes) */
  for (int i = 0; i < c->bins; ++i) c->sq[i] = 0.0;
  for (int i = c->start_index; i <= c->end_index; ++i) {
    const int ch = c->band[i];
    double v = 0.0;
    if (ch >= 0) v += (double)(float)exp((double)(log_mel_features[ch] * 10.0f)) * c->weights[i] / c->norm[ch];
    if (ch + 1 < c->nmel) v += (double)(float)exp((double)(log_mel_features[ch + 1] * 10.0f)) * (1.0 - c->weights[i]) / c->norm[ch + 1];
    c->sq[i] = v * v;
  }
  /* InvertFft (.cc:98-119): magnitude * e^{j phase}, Hermitian spectrum of a real signal (DC and Nyquist real) */
  for (int i = 0; i < N; ++i) { c->re[i] = 0.0; c->im[i] = 0.0; }
  for (int i = 0; i < c->bins; ++i) {
    const double mag = sqrt(c->sq[i]);
    const uint32_t p = phase ? (phase[i] & 1023u) : lo_cng_phase_index(c->seed, c->hops, i);
    /* unit vector at angle 2 pi p / 1024: the FFT twiddle (cos, sin)(-2 pi k / 1024), k = p mod 512, conjugated and negated for p >= 512 */
    const double ang = -2.0 * M_PI * (double)(p & 511u) / (double)1024;
    double cr = cos(ang), ci = -sin(ang);
    if (p >= 512u) { cr = -cr; ci = -ci; }
    const double xr = mag * cr, xi = mag * ci;
    if (i == 0 || i == N / 2) { c->re[i] = xr; }
    else { c->re[i] = xr; c->im[i] = xi; c->re[N - i] = xr; c->im[N - i] = -xi; }
  }
  ifft_inplace(c->re, c->im, N);
  /* window, scale (1/N of the inverse DFT), overlap-add; emit one hop; shift */
  for (int i = 0; i < N; ++i) c->work[i] += (c->re[i] / (double)N) * c->synth[i];
  for (int i = 0; i < c->hop; ++i) {
    double v = c->work[i];                                   /* ClipToInt16Scalar<double> (dsp_utils.h:53-60): clamp, then truncate */
    v = v > -32768.0 ? v : -32768.0;
    v = v < 32767.0 ? v : 32767.0;
    hop_out[i] = (int16_t)v;
  }
  memmove(c->work, c->work + c->hop, sizeof(double) * (size_t)(N - c->hop));
  for (int i = N - c->hop; i < N; ++i) c->work[i] = 0.0;
  c->hops++;
  return 0;
}

static int cng_conditioning(void* self, const float* features, int16_t* hop_out) {
  return lo_cng_condition((lo_cng*)self, features, NULL, hop_out);
}
