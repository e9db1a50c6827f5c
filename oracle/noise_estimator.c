/* TEST INFRASTRUCTURE ONLY (see lyra_oracle.h).
 *
 * CPU restatement of the reference's minimum-statistics noise estimator that runs on the decoder output
 * (SURVEY.md section 8 row f1):
 *   NoiseEstimator::Create               lyra/noise_estimator.cc:99-120   (constants)
 *   NoiseEstimator::ReceiveSamples       lyra/noise_estimator.cc:144-172  (full hops only here)
 *   NoiseEstimator::UpdateNoiseEstimate  lyra/noise_estimator.cc:174-205
 *   SmoothingFactor / UpdateMinAndTemp   lyra/noise_estimator.cc:37-95
 *   ComputeBounds / ComputeIsNoise / DecayBounds  lyra/noise_estimator.cc:207-245
 * Arithmetic follows the C++ expression types literally: float unless an operand is double (std::log of an
 * integer is double, so the bound is evaluated in double and rounded to float on assignment); every float
 * operation rounds separately (the oracle is built with -ffp-contract=off, like a default x86-64 build of
 * the reference).  std::exp(float) is evaluated as (float)exp((double)x): glibc's expf and exp are both
 * correctly rounded in all but astronomically rare cases, and this form is what the CUDA kernel can
 * reproduce bit for bit.
 * Parity pinning: the reference's own tests for this class are statistical (random comfort noise, loose
 * thresholds; lyra/noise_estimator_test.cc:131-197); tests/test_oracle_golden.py re-runs their
 * NoiseIdentification and FiveSecondsSilence properties against this restatement. */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "lyra_oracle.h"

struct lo_noise {
  int16_t past[2048];       /* past_samples_hop_ (noise_estimator.cc:144-160), hop <= 2048 */
  int next_sample_in_hop;
  int nf, hop, hops_per_update;
  float max_smoothing, bound_decay;
  float *smoothed, *sq_smoothed, *tmp_min, *est, *bound;
  int has_smoothed, is_noise, hops_received;
  lo_logmel* lm;
};

static float expf_via_double(float x) { return (float)exp((double)x); }

lo_noise* lo_noise_create(int sample_rate_hz, int hop, int window, int num_features) {
  lo_noise* e = (lo_noise*)calloc(1, sizeof(lo_noise));
  if (!e) return NULL;
  e->lm = lo_logmel_create(sample_rate_hz, hop, window, num_features);
  if (!e->lm) { free(e); return NULL; }
  /* noise_estimator.cc:99-120 */
  const float secs_per_hop = (float)hop / sample_rate_hz;
  e->nf = num_features;
  e->hop = hop;
  e->hops_per_update = (int)roundf(1.f / secs_per_hop);
  e->max_smoothing = powf(0.5f, secs_per_hop / 0.7f);
  e->bound_decay = powf(0.5f, secs_per_hop / 1.f);
  e->smoothed = (float*)calloc((size_t)num_features * 5, sizeof(float));
  e->sq_smoothed = e->smoothed + num_features;
  e->tmp_min = e->sq_smoothed + num_features;
  e->est = e->tmp_min + num_features;
  e->bound = e->est + num_features;
  e->is_noise = 1;
  return e;
}

void lo_noise_free(lo_noise* e) {
  if (!e) return;
  lo_logmel_free(e->lm);
  free(e->smoothed);
  free(e);
}

void lo_noise_set_constants(lo_noise* e, int hops_per_update, float max_smoothing, float bound_decay) {
  e->hops_per_update = hops_per_update;
  e->max_smoothing = max_smoothing;
  e->bound_decay = bound_decay;
}

static float average(const float* v, int n) {
  float s = 0.f;
  for (int i = 0; i < n; ++i) s = s + v[i];
  return s / (float)n;
}

/* noise_estimator.cc:231-241 */
int lo_noise_compute_is_noise(const lo_noise* e, const float* cur) {
  for (int i = 0; i < e->nf; ++i)
    if (fabsf(cur[i] - e->est[i]) > e->bound[i]) return 0;
  return 1;
}

/* noise_estimator.cc:174-229 */
void lo_noise_update(lo_noise* e, const float* cur) {
  const int n = e->nf;
  if (!e->has_smoothed) {
    for (int i = 0; i < n; ++i) { e->smoothed[i] = cur[i]; e->sq_smoothed[i] = cur[i] * cur[i]; e->tmp_min[i] = cur[i]; }
    e->has_smoothed = 1;
  }
  const float kPowDiff = 0.3f;
  const float q = (average(e->smoothed, n) - average(cur, n)) / kPowDiff;
  const float correction = expf_via_double(-(q * q));
  for (int i = 0; i < n; ++i) {
    const float r = (e->smoothed[i] - e->est[i]) / kPowDiff;
    const float sf = e->max_smoothing * correction * expf_via_double(-(r * r));
    const float a = sf * e->smoothed[i];
    const float b = (1.f - sf) * cur[i];
    const float c = sf * e->sq_smoothed[i];
    const float d = (1.f - sf) * (cur[i] * cur[i]);
    e->smoothed[i] = a + b;
    e->sq_smoothed[i] = c + d;
  }
  if (e->hops_received == 0) {
    for (int i = 0; i < n; ++i) {
      e->est[i] = e->smoothed[i] < e->tmp_min[i] ? e->smoothed[i] : e->tmp_min[i];   /* std::min(a, b): b < a ? b : a */
      e->tmp_min[i] = e->smoothed[i];
    }
  } else {
    for (int i = 0; i < n; ++i) {
      e->est[i] = e->smoothed[i] < e->est[i] ? e->smoothed[i] : e->est[i];
      e->tmp_min[i] = e->smoothed[i] < e->tmp_min[i] ? e->smoothed[i] : e->tmp_min[i];
    }
  }
  const double log_n = log((double)n);
  for (int i = 0; i < n; ++i) {
    const float t = e->sq_smoothed[i] - e->smoothed[i] * e->smoothed[i];
    const float var = t > 0.f ? t : 0.f;                               /* std::max<float>(0.f, t) */
    e->bound[i] = (float)((double)0.9f * sqrt((double)var * log_n));
  }
  e->hops_received = (e->hops_received + 1) % e->hops_per_update;
}

/* noise_estimator.cc:144-172 for one full hop */
int lo_noise_receive_samples(lo_noise* e, const int16_t* hop, float* logmel_out) {
  float* cur = (float*)malloc(sizeof(float) * (size_t)e->nf);
  if (!cur) return -1;
  if (lo_logmel_extract(e->lm, hop, e->hop, cur)) { free(cur); return -1; }
  e->is_noise = lo_noise_compute_is_noise(e, cur);
  if (e->is_noise) {
    for (int i = 0; i < e->nf; ++i) e->bound[i] = e->bound[i] * e->bound_decay;
  } else {
    lo_noise_update(e, cur);
  }
  if (logmel_out) memcpy(logmel_out, cur, sizeof(float) * (size_t)e->nf);
  free(cur);
  return 0;
}

/* noise_estimator.cc:144-172 with the partial-hop buffering: samples accumulate until a hop is complete; a call may
 * not straddle a hop boundary */
int lo_noise_receive_partial(lo_noise* e, const int16_t* samples, int n) {
  if (n < 0 || e->hop > 2048 || n + e->next_sample_in_hop > e->hop) return -1;
  memcpy(e->past + e->next_sample_in_hop, samples, sizeof(int16_t) * (size_t)n);
  e->next_sample_in_hop += n;
  if (e->next_sample_in_hop == e->hop) {
    e->next_sample_in_hop = 0;
    return lo_noise_receive_samples(e, e->past, NULL);
  }
  return 0;
}

int lo_noise_is_noise(const lo_noise* e) { return e->is_noise; }
void lo_noise_estimate(const lo_noise* e, float* out) { memcpy(out, e->est, sizeof(float) * (size_t)e->nf); }
void lo_noise_bound(const lo_noise* e, float* out) { memcpy(out, e->bound, sizeof(float) * (size_t)e->nf); }
