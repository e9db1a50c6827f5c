/*
 * ORACLE — TEST INFRASTRUCTURE ONLY (see lyra_oracle.h).
 *
 * Log-mel spectrogram extractor, restating LogMelSpectrogramExtractorImpl
 * (lyra/log_mel_spectrogram_extractor_impl.cc:37-126).  The arithmetic lives in the un-vendored
 * dependency com_google_audio_dsp = mchinen/multichannel-audio-tools@14a45c5 (reference
 * WORKSPACE:68-78): audio_dsp::Spectrogram (periodic Hann window, zero-padded power-of-two real
 * FFT in double, squared magnitude) and audio_dsp::MelFilterbank (the TensorFlow MFCC mel
 * filterbank: triangular weights in mel space applied to sqrt(power)).  Their published algorithms
 * are restated here; the restatement is pinned by the reference's golden vectors
 * (lyra/log_mel_spectrogram_extractor_impl_test.cc:37-59) in tests/test_oracle_golden.py.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "lyra_oracle.h"

struct lo_logmel {
  int sample_rate, hop, window, nmel, fft, bins;
  double* win;       /* [window] periodic Hann */
  double* queue;     /* last `window` samples (the spectrogram's input queue) */
  double* re;        /* [fft] scratch */
  double* im;
  double* power;     /* [bins] */
  double* weights;   /* [bins] */
  int* band;         /* [bins] */
  int start_index, end_index;
  double* mel;       /* [nmel] scratch */
};

static double freq_to_mel(double f) { return 1127.0 * log1p(f / 700.0); }

static void fft_inplace(double* re, double* im, int n) {
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
        const double wr = cos(ang * k), wi = sin(ang * k);
        const int a = i + k, b = i + k + len / 2;
        const double xr = re[b] * wr - im[b] * wi, xi = re[b] * wi + im[b] * wr;
        re[b] = re[a] - xr; im[b] = im[a] - xi;
        re[a] += xr; im[a] += xi;
      }
  }
}

lo_logmel* lo_logmel_create(int sample_rate_hz, int hop, int window, int num_mel_bins) {
  if (window < hop || hop <= 0 || num_mel_bins <= 0) return NULL;   /* .cc:57-62 */
  lo_logmel* m = (lo_logmel*)calloc(1, sizeof(*m));
  m->sample_rate = sample_rate_hz; m->hop = hop; m->window = window; m->nmel = num_mel_bins;
  m->fft = 1;
  while (m->fft < window) m->fft <<= 1;             /* audio_dsp::NextPowerOfTwo (.cc:79-80) */
  m->bins = m->fft / 2 + 1;
  m->win = (double*)malloc(sizeof(double) * (size_t)window);
  for (int i = 0; i < window; ++i) m->win[i] = 0.5 - 0.5 * cos(2.0 * M_PI * i / (double)window);
  /* Create() pushes one all-zero window through the spectrogram (.cc:69-77): queue starts as zeros */
  m->queue = (double*)calloc((size_t)window, sizeof(double));
  m->re = (double*)malloc(sizeof(double) * (size_t)m->fft);
  m->im = (double*)malloc(sizeof(double) * (size_t)m->fft);
  m->power = (double*)malloc(sizeof(double) * (size_t)m->bins);
  m->weights = (double*)calloc((size_t)m->bins, sizeof(double));
  m->band = (int*)malloc(sizeof(int) * (size_t)m->bins);
  m->mel = (double*)malloc(sizeof(double) * (size_t)num_mel_bins);

  /* MelFilterbank::Initialize(bins, sample_rate, num_mel, lower = 0.0, upper = 0.495 * fs) (.cc:39-40,84-90) */
  const double lower = 0.0, upper = 0.495 * sample_rate_hz;
  const double mel_low = freq_to_mel(lower), mel_hi = freq_to_mel(upper);
  const double mel_spacing = (mel_hi - mel_low) / (double)(num_mel_bins + 1);
  double* center = (double*)malloc(sizeof(double) * (size_t)(num_mel_bins + 1));
  for (int i = 0; i < num_mel_bins + 1; ++i) center[i] = mel_low + mel_spacing * (i + 1);
  const double hz_per_sbin = 0.5 * sample_rate_hz / (double)(m->bins - 1);
  m->start_index = (int)(1.5 + lower / hz_per_sbin);
  m->end_index = (int)(upper / hz_per_sbin);
  int channel = 0;
  for (int i = 0; i < m->bins; ++i) {
    const double melf = freq_to_mel(i * hz_per_sbin);
    if (i < m->start_index || i > m->end_index) {
      m->band[i] = -2;
    } else {
      while (channel < num_mel_bins && center[channel] < melf) ++channel;
      m->band[i] = channel - 1;
    }
  }
  for (int i = 0; i < m->bins; ++i) {
    channel = m->band[i];
    if (i < m->start_index || i > m->end_index) {
      m->weights[i] = 0.0;
    } else if (channel >= 0) {
      m->weights[i] = (center[channel + 1] - freq_to_mel(i * hz_per_sbin)) / (center[channel + 1] - center[channel]);
    } else {
      m->weights[i] = (center[0] - freq_to_mel(i * hz_per_sbin)) / (center[0] - mel_low);
    }
  }
  free(center);
  return m;
}

void lo_logmel_free(lo_logmel* m) {
  if (!m) return;
  free(m->win); free(m->queue); free(m->re); free(m->im); free(m->power);
  free(m->weights); free(m->band); free(m->mel);
  free(m);
}

int lo_logmel_extract(lo_logmel* m, const int16_t* audio, int n, float* out) {
  if (n != m->hop) return -1;                                     /* .cc:98-102 */
  /* slide the queue by one hop; int16 samples enter as doubles without scaling (.cc:104) */
  memmove(m->queue, m->queue + m->hop, sizeof(double) * (size_t)(m->window - m->hop));
  for (int i = 0; i < m->hop; ++i) m->queue[m->window - m->hop + i] = (double)audio[i];
  for (int i = 0; i < m->fft; ++i) {
    m->re[i] = i < m->window ? m->queue[i] * m->win[i] : 0.0;
    m->im[i] = 0.0;
  }
  fft_inplace(m->re, m->im, m->fft);
  for (int i = 0; i < m->bins; ++i) m->power[i] = m->re[i] * m->re[i] + m->im[i] * m->im[i];
  /* MelFilterbank::Compute */
  for (int c = 0; c < m->nmel; ++c) m->mel[c] = 0.0;
  for (int i = m->start_index; i <= m->end_index; ++i) {
    const double spec_val = sqrt(m->power[i]);
    const double weighted = spec_val * m->weights[i];
    int channel = m->band[i];
    if (channel >= 0) m->mel[channel] += weighted;
    ++channel;
    if (channel < m->nmel) m->mel[channel] += spec_val - weighted;
  }
  /* cast to float, floor at 500, log, / 10 (.cc:118-123) */
  for (int c = 0; c < m->nmel; ++c) {
    float v = (float)m->mel[c];
    v = v > 500.0f ? v : 500.0f;
    /* std::log(float): evaluated in double and rounded once, i.e. the correctly rounded logf; libm's logf (<= 0.82 ulp in
     * glibc) may differ from it by one ulp on rare inputs, which is inside the reference test's own tolerance */
    out[c] = (float)log((double)v) / 10.0f;
  }
  return 0;
}
