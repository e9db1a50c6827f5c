/* TEST INFRASTRUCTURE ONLY (see lyra_oracle.h).
 *
 * CPU restatement of the reference's sample-rate converters:
 *   Resampler::{Create, Resample, Reset, samples_until_steady_state}   lyra/resampler.cc:31-83
 *   BufferedResampler::FilterAndBuffer and helpers                      lyra/buffered_resampler.cc:63-147
 * The filter itself lives in the un-vendored dependency com_google_audio_dsp = mchinen/multichannel-audio-tools@14a45c5
 * (reference WORKSPACE:68-78): audio_dsp::QResampler<float>, a polyphase FIR resampler for rational rate ratios whose
 * published design is restated here - Kaiser-windowed sinc kernel
 *     h(x) = 2 fc sinc(2 fc x) I0(beta sqrt(1 - (x / radius)^2)) / I0(beta),   |x| <= radius (input samples),
 *     radius = filter_radius_factor * max(1, in / out), fc = cutoff_proportion * 0.5 / max(1, in / out),
 * one filter of 2 ceil(radius) + 1 taps per output phase, library defaults cutoff_proportion 0.9 and kaiser_beta 5.658,
 * Lyra's radius of 17 input samples (resampler.cc:33-38), started "fully primed" (resampler.cc:60): the delay line begins
 * as zeros and the first output is aligned with input sample -radius, i.e. a constant delay of 17 input samples.
 * The reference holds no golden vector for the filter ("PARITY UNPINNED" at the coefficient level); this restatement is
 * pinned by the reference's own resampler tests (all-zero input, output sizes at 8/16/32/48 kHz, up-then-down similarity
 * with the documented delay of 17 + floor(17 / 2) samples, clipping; lyra/resampler_test.cc:33-102) and the buffered
 * resampler's leftover bookkeeping (lyra/buffered_resampler_test.cc:86-240), re-run in tests/test_oracle_resampler.py.
 * Canonical arithmetic (what the CUDA kernel reproduces bit for bit): coefficients computed in double and stored as float;
 * each output is a float sum over the taps in ascending input order, one separately rounded multiply and add per tap.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "lyra_oracle.h"

static int gcd_i(int a, int b) { while (b) { const int t = a % b; a = b; b = t; } return a; }

/* modified Bessel function of the first kind, order 0 (power series; converges fast for the small arguments used here) */
double lo_bessel_i0(double x) {
  double sum = 1.0, term = 1.0;
  const double q = x * x / 4.0;
  for (int k = 1; k < 64; ++k) {
    term *= q / ((double)k * (double)k);
    sum += term;
    if (term < 1e-17 * sum) break;
  }
  return sum;
}

/* coeffs[phase * taps + k] multiplies delay-line sample k (oldest first); returns the number of taps */
int lo_resampler_design(int input_rate, int output_rate, int* num, int* den, float* coeffs, int capacity) {
  const int g = gcd_i(input_rate, output_rate);
  *num = input_rate / g;                     /* factor = input / output = num / den */
  *den = output_rate / g;
  const double factor = (double)*num / (double)*den;
  const double radius_factor = 17.0 * (output_rate < input_rate ? (double)((float)output_rate / (float)input_rate) : 1.0);   /* resampler.cc:36-38 */
  const double radius = radius_factor * (factor > 1.0 ? factor : 1.0);
  const double cutoff = 0.9 * 0.5 / (factor > 1.0 ? factor : 1.0);
  const double beta = 5.658;
  const int rc = (int)ceil(radius - 1e-4);      /* 17 for every supported pair (the float product 17 * (out / in) * (in / out) may exceed 17 by an ulp) */
  const int taps = 2 * rc + 1;
  if (*den * taps > capacity) return -1;
  const double i0b = lo_bessel_i0(beta);
  for (int p = 0; p < *den; ++p) {
    const double offset = (double)p / (double)*den;
    /* output at input position t = i + offset - rc reads delay-line samples m = i - 2 rc .. i: x = t - m = (rc - j) + offset - ... */
    for (int j = 0; j < taps; ++j) {
      const double x = (double)(rc - j) + offset;      /* distance from the output position (i - rc + offset) to sample i - 2 rc + j */
      double h = 0.0;
      if (fabs(x) <= radius) {
        const double z = 2.0 * cutoff * x;
        const double sinc = fabs(z) < 1e-12 ? 1.0 : sin(M_PI * z) / (M_PI * z);
        const double r = x / radius;
        h = 2.0 * cutoff * sinc * lo_bessel_i0(beta * sqrt(1.0 - r * r > 0.0 ? 1.0 - r * r : 0.0)) / i0b;
      }
      coeffs[p * taps + j] = (float)h;
    }
  }
  return taps;
}

struct lo_resampler {
  int in_rate, out_rate, num, den, taps, rc;
  float coeffs[3 * 35];
  float delay[34];          /* the last taps - 1 input samples (zeros after Reset: fully primed) */
  int phase;                /* position of the next output inside the current input sample, in units of 1 / den */
  int skip;                 /* input samples still to be consumed before the next output (phase bookkeeping across calls) */
};

lo_resampler* lo_resampler_create(int input_rate, int output_rate) {
  static const int ok[] = {8000, 16000, 32000, 48000};
  int a = 0, b = 0;
  for (int i = 0; i < 4; ++i) { a |= ok[i] == input_rate; b |= ok[i] == output_rate; }
  if (!a || !b) return NULL;
  lo_resampler* r = (lo_resampler*)calloc(1, sizeof(*r));
  r->in_rate = input_rate; r->out_rate = output_rate;
  r->taps = lo_resampler_design(input_rate, output_rate, &r->num, &r->den, r->coeffs, 3 * 35);
  if (r->taps != 35) { free(r); return NULL; }
  r->rc = 17;
  return r;
}
void lo_resampler_free(lo_resampler* r) { free(r); }
void lo_resampler_reset(lo_resampler* r) { memset(r->delay, 0, sizeof(r->delay)); r->phase = 0; r->skip = 0; }
int lo_resampler_input_rate(const lo_resampler* r) { return r->in_rate; }
int lo_resampler_output_rate(const lo_resampler* r) { return r->out_rate; }
int lo_resampler_samples_until_steady_state(const lo_resampler* r) {       /* resampler.cc:74-83 */
  const float ratio = (float)r->den / (float)r->num;
  return (int)(2.f * 17.f * ratio);
}
void lo_resampler_coeffs(const lo_resampler* r, float* out) { memcpy(out, r->coeffs, sizeof(float) * (size_t)(r->den * r->taps)); }

/* Resampler::Resample (resampler.cc:52-57): int16 -> float, polyphase FIR, ClipToInt16.  Returns the number of outputs. */
int lo_resampler_resample(lo_resampler* r, const int16_t* in, int n, int16_t* out, int capacity) {
  const int T = r->taps;
  int produced = 0;
  /* work on the concatenation [delay (T - 1 samples) | in (n samples)]; index i below is the newest sample an output reads */
  for (int i = 0; i < n; ++i) {
    if (r->skip > 0) { r->skip--; continue; }
    /* every output whose window ends at input sample i: phases p, p + num, ... while they stay inside this sample */
    while (r->phase < r->den) {
      const float* c = r->coeffs + r->phase * T;
      float acc = 0.0f;
      for (int j = 0; j < T; ++j) {
        const int m = i - (T - 1) + j;                       /* input index of delay-line sample j */
        const float x = m >= 0 ? (float)in[m] : r->delay[(T - 1) + m];
        acc = acc + c[j] * x;
      }
      if (produced >= capacity) return -1;
      float v = acc;                                          /* ClipToInt16Scalar<float> (dsp_utils.h:53-60) */
      v = v > -32768.0f ? v : -32768.0f;
      v = v < 32767.0f ? v : 32767.0f;
      out[produced++] = (int16_t)v;
      r->phase += r->num;
    }
    r->phase -= r->den;
    r->skip = r->phase / r->den;                              /* whole input samples to step over (down-sampling) */
    r->phase %= r->den;
  }
  /* new delay line = the last T - 1 samples of [delay | in] */
  float nd[34];
  for (int j = 0; j < T - 1; ++j) {
    const int m = n - (T - 1) + j;
    nd[j] = m >= 0 ? (float)in[m] : r->delay[(T - 1) + m];
  }
  memcpy(r->delay, nd, sizeof(nd));
  return produced;
}

/* ------------------------------------------------------------ BufferedResampler (buffered_resampler.cc:63-147) ---- */

struct lo_buffered_resampler {
  lo_resampler* r;          /* internal -> external */
  int16_t leftover[8];
  int n_leftover;
};

lo_buffered_resampler* lo_buffered_resampler_create(int internal_rate, int external_rate) {
  lo_resampler* r = lo_resampler_create(internal_rate, external_rate);
  if (!r) return NULL;
  lo_buffered_resampler* b = (lo_buffered_resampler*)calloc(1, sizeof(*b));
  b->r = r;
  return b;
}
void lo_buffered_resampler_free(lo_buffered_resampler* b) { if (b) { lo_resampler_free(b->r); free(b); } }
int lo_buffered_resampler_leftover(const lo_buffered_resampler* b) { return b->n_leftover; }

int lo_buffered_resampler_internal_samples(const lo_buffered_resampler* b, int num_external_requested) {      /* :93-106 */
  if (num_external_requested <= b->n_leftover) return 0;
  const int needed = num_external_requested - b->n_leftover;
  const float ratio = (float)b->r->out_rate / (float)b->r->in_rate;
  return (int)ceilf((float)needed / ratio);
}

/* generator(user, n, out) must write exactly n internal-rate samples and return 0 */
int lo_buffered_resampler_filter_and_buffer(lo_buffered_resampler* b, int (*generator)(void*, int, int16_t*), void* user,
                                            int num_external_requested, int16_t* out) {
  const int n_int = lo_buffered_resampler_internal_samples(b, num_external_requested);
  const int used = b->n_leftover < num_external_requested ? b->n_leftover : num_external_requested;          /* :108-119 */
  memcpy(out, b->leftover, sizeof(int16_t) * (size_t)used);
  memmove(b->leftover, b->leftover + used, sizeof(int16_t) * (size_t)(b->n_leftover - used));
  b->n_leftover -= used;
  int16_t* internal = (int16_t*)malloc(sizeof(int16_t) * (size_t)(n_int + 1));
  if (generator(user, n_int, internal) != 0) { free(internal); return -1; }
  int16_t* external = internal;
  int n_ext = n_int;
  int16_t* tmp = NULL;
  if (b->r->in_rate != b->r->out_rate) {                                                                      /* :121-129 */
    tmp = (int16_t*)malloc(sizeof(int16_t) * (size_t)(n_int * 3 + 8));
    n_ext = lo_resampler_resample(b->r, internal, n_int, tmp, n_int * 3 + 8);
    external = tmp;
  }
  const int to_copy = num_external_requested - used;                                                          /* :131-147 */
  if (n_ext < to_copy) { free(internal); free(tmp); return -1; }
  memcpy(out + used, external, sizeof(int16_t) * (size_t)to_copy);
  for (int i = to_copy; i < n_ext && b->n_leftover < 8; ++i) b->leftover[b->n_leftover++] = external[i];
  free(internal);
  free(tmp);
  return num_external_requested;
}
