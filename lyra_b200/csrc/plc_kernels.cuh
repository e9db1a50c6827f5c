// Decoder packet-loss path as batched device steps:
//   PlcPlanKernel        the per-stream concealment / fade bookkeeping of LyraDecoder::SetEncodedPacket +
//                        DecodeSamplesInternal for one whole hop per tick          (lyra/lyra_decoder.cc:186-196, 228-283, 303)
//   ComfortNoiseKernel   ComfortNoiseGenerator::RunConditioning + RunModel(hop)   (lyra/comfort_noise_generator.cc:74-119)
//   PlcMixKernel         LyraDecoder::MaybeOverlapAndInsert                        (lyra/lyra_decoder.cc:342-373)
// Arithmetic mirrors oracle/comfort_noise.c and oracle/lyra_decoder.c operation by operation (separately rounded f64 / f32
// operations, host-computed tables for every transcendental), so the results are bit-identical to the oracle.
#pragma once

#include "aux_kernels.cuh"

namespace lyra_b200 {

// per-stream decoder control state: {concealment_progress, fade_progress, fade_direction (-1 from / +1 to comfort noise), -}
constexpr int kPlcConcealSamples = 1280, kPlcFadeSamples = 640;

// What one tick does to one stream (whole-hop regime: one optional packet, then 320 samples):
//   received packet:  concealment_progress <- 0 (a whole fake hop has always been played out), features queued
//   then the decode step of lyra_decoder.cc:249-283 with num_samples_to_generate = 320.
// plan[slot] bits: 1 = run the generative model, 2 = run the comfort-noise generator, 4 = the hop comes from a received packet
// (feed the noise estimator); fade0[slot] = fade progress before the hop, dir[slot] = fade direction of the hop.
__global__ void __launch_bounds__(256)
PlcPlanKernel(const int* __restrict__ stream_ids, int n, const uint8_t* __restrict__ received, int* __restrict__ state,
              uint8_t* __restrict__ plan, int* __restrict__ fade0, int* __restrict__ dir_out, uint8_t* __restrict__ skip_model,
              uint8_t* __restrict__ feed_mask, uint8_t* __restrict__ is_comfort_noise) {
  const int slot = (int)(blockIdx.x * blockDim.x + threadIdx.x);
  if (slot >= n) return;
  const int stream = stream_ids ? stream_ids[slot] : slot;
  int* st = state + (size_t)stream * 4;
  int cp = st[0], fp = st[1], dir = st[2];
  const bool rec = received == nullptr || received[slot] != 0;
  if (rec && cp > 0) cp = 0;                                  // SetEncodedPacket :186-196 (nothing left of a fake hop at a hop boundary)
  const bool is_packet_received = rec && cp == 0;             // :249-251 (the model queue holds exactly this tick's packet)
  if (is_packet_received) dir = -1;                           // :253-256
  else if (cp == kPlcConcealSamples) dir = +1;                // :257-260
  else cp += 320;                                             // :261-265
  int gen = 1, cng = 1;
  int next_fp = fp + dir * 320;                               // :269-270
  if (dir == +1 && fp == kPlcFadeSamples) { next_fp = kPlcFadeSamples; gen = 0; }       // :271-276
  else if (dir == -1 && fp == 0) { next_fp = 0; cng = 0; }                               // :277-282
  plan[slot] = (uint8_t)(gen | (cng << 1) | (is_packet_received ? 4 : 0));
  fade0[slot] = fp;
  dir_out[slot] = dir;
  skip_model[slot] = gen ? 0 : 1;
  feed_mask[slot] = is_packet_received ? 1 : 0;
  st[0] = cp; st[1] = next_fp; st[2] = dir;                   // :303
  if (is_comfort_noise) is_comfort_noise[slot] = next_fp == kPlcFadeSamples ? 1 : 0;     // :381-383
}

// splitmix64 of (seed, hop, bin): the phase index (0..1023) of oracle/comfort_noise.c lo_cng_phase_index
__device__ __forceinline__ uint32_t CngPhaseIndex(unsigned long long seed, unsigned long long hop, int bin) {
  unsigned long long x = seed ^ (hop * 0xD1B54A32D192ED03ull) ^ ((unsigned long long)bin * 0x9E3779B97F4A7C15ull);
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  x ^= x >> 31;
  return (uint32_t)(x >> 54);
}

// three consecutive inverse-transform stages: FftButterflies3 with conjugated twiddles
template <int STRIDE>
__device__ __forceinline__ void IfftButterflies3(double (&xr)[8], double (&xi)[8], int r, const double2* __restrict__ tw) {
#pragma unroll
  for (int s = 0; s < 3; ++s) {
    const int half = STRIDE << s;
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (!(j & (1 << s))) {
        const int k = r + (j & ((1 << s) - 1)) * STRIDE;
        const double2 w = tw[half - 1 + k];
        FftButterfly(xr[j], xi[j], xr[j + (1 << s)], xi[j + (1 << s)], w.x, -w.y);
      }
  }
}
template <int STRIDE>
__device__ __forceinline__ void IfftStages3(double* re, double* im, int base, int r, const double2* __restrict__ tw) {
  double xr[8], xi[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { xr[j] = re[FftIdx(base + j * STRIDE)]; xi[j] = im[FftIdx(base + j * STRIDE)]; }
  IfftButterflies3<STRIDE>(xr, xi, r, tw);
#pragma unroll
  for (int j = 0; j < 8; ++j) { re[FftIdx(base + j * STRIDE)] = xr[j]; im[FftIdx(base + j * STRIDE)] = xi[j]; }
}

// One block of 128 threads per stream.  features: [n][160] conditioning log-mel vectors (by slot), or nullptr = the stream's
// current noise estimate (the first 160 floats of its noise-estimator state, LyraDecoder::RunComfortNoiseGenerator,
// lyra/lyra_decoder.cc:328-340).  plan: nullptr = every slot runs; otherwise only slots with bit 2.
// work: [max_streams][1024] f64 overlap-add buffers; hops: [max_streams] hop counters (phase draw).
constexpr int kCngThreads = 128;
__global__ void __launch_bounds__(kCngThreads)
ComfortNoiseKernel(const uint8_t* __restrict__ blob, CngParams P, const int* __restrict__ stream_ids, int n,
                   const float* __restrict__ features, const float* __restrict__ noise_state, int noise_units,
                   const uint8_t* __restrict__ plan, double* __restrict__ work, unsigned long long* __restrict__ hops,
                   unsigned long long seed, int16_t* __restrict__ out, int slot_base) {
  unsigned char* smem = LYRA_DYN_SMEM();
  double* re = reinterpret_cast<double*>(smem);
  double* im = re + kLogMelFftPadded;
  double* xr0 = im + kLogMelFftPadded;           // spectrum in natural order: real | imag, padded like the FFT buffers
  double* xi0 = xr0 + kLogMelFftPadded;
  double* mel = xi0 + kLogMelFftPadded;          // [160]
  const int slot = slot_base + (int)blockIdx.x;
  if (slot >= n) return;
  if (plan && !(plan[slot] & 2)) return;
  const int stream = stream_ids ? stream_ids[slot] : slot;
  const int tid = (int)threadIdx.x;
  constexpr int NT = kCngThreads, N = kLogMelFft;
  const double* wts = BlobPtr<double>(blob, P.weights);
  const int* band = BlobPtr<int>(blob, P.band);
  const double* norm = BlobPtr<double>(blob, P.norm);
  const double* synth = BlobPtr<double>(blob, P.synth);
  const double2* tw = BlobPtr<double2>(blob, P.twiddle);
  const float* f = features ? features + (size_t)slot * P.num_mel : noise_state + (size_t)stream * noise_units;
  const unsigned long long hop_index = hops[stream];
  const unsigned long long sseed = seed + (unsigned long long)stream;
  // mel[c] = (double)(float)exp(feature * 10)   (FftFromFeatures, .cc:87-96; exp of a float evaluated in double, rounded once)
  for (int c = tid; c < P.num_mel; c += NT) mel[c] = (double)(float)exp((double)__fmul_rn(f[c], 10.0f));
  for (int i = tid; i < N; i += NT) { xr0[FftIdx(i)] = 0.0; xi0[FftIdx(i)] = 0.0; }
  __syncthreads();
  // mel inverse -> magnitude -> unit vector at the drawn phase -> Hermitian spectrum
  for (int i = tid; i <= N / 2; i += NT) {
    double v = 0.0;
    if (i >= P.start_index && i <= P.end_index) {
      const int ch = band[i];
      if (ch >= 0) v = __dadd_rn(v, __ddiv_rn(__dmul_rn(mel[ch], wts[i]), norm[ch]));
      if (ch + 1 < P.num_mel) v = __dadd_rn(v, __ddiv_rn(__dmul_rn(mel[ch + 1], __dsub_rn(1.0, wts[i])), norm[ch + 1]));
    }
    const double mag = __dsqrt_rn(__dmul_rn(v, v));
    const uint32_t p = CngPhaseIndex(sseed, hop_index, i);
    const double2 w = tw[511 + (p & 511u)];                  // last stage's table: (cos, sin)(-2 pi k / 1024)
    double cr = w.x, ci = -w.y;
    if (p >= 512u) { cr = -cr; ci = -ci; }
    const double vr = __dmul_rn(mag, cr), vi = __dmul_rn(mag, ci);
    if (i == 0 || i == N / 2) { xr0[FftIdx(i)] = vr; }
    else { xr0[FftIdx(i)] = vr; xi0[FftIdx(i)] = vi; xr0[FftIdx(N - i)] = vr; xi0[FftIdx(N - i)] = -vi; }
  }
  __syncthreads();
  {                                                                            // stages 2, 4, 8 (bit reversal folded into the gather)
    double xr[8], xi[8];
    const int b7 = (int)(__brev((unsigned)tid) >> 25);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int j3 = ((j & 1) << 2) | (j & 2) | ((j & 4) >> 2);
      const int i = b7 + 128 * j3;
      xr[j] = xr0[FftIdx(i)];
      xi[j] = xi0[FftIdx(i)];
    }
    IfftButterflies3<1>(xr, xi, 0, tw);
#pragma unroll
    for (int j = 0; j < 8; ++j) { re[FftIdx(8 * tid + j)] = xr[j]; im[FftIdx(8 * tid + j)] = xi[j]; }
  }
  __syncthreads();
  IfftStages3<8>(re, im, 64 * (tid / 8) + tid % 8, tid % 8, tw);               // stages 16, 32, 64
  __syncthreads();
  IfftStages3<64>(re, im, 512 * (tid / 64) + tid % 64, tid % 64, tw);          // stages 128, 256, 512
  __syncthreads();
#pragma unroll
  for (int m = 0; m < 4; ++m) {                                                // stage 1024
    const int a = tid + NT * m, ia = FftIdx(a), ib = FftIdx(a + 512);
    const double2 w = tw[511 + a];
    FftButterfly(re[ia], im[ia], re[ib], im[ib], w.x, -w.y);
  }
  __syncthreads();
  // 1/N, synthesis window, overlap-add; emit one hop (ClipToInt16: clamp, truncate); shift the buffer by one hop
  double* wk = work + (size_t)stream * N;
  for (int i = tid; i < N; i += NT) {
    const double v = __dadd_rn(wk[i], __dmul_rn(__ddiv_rn(re[FftIdx(i)], (double)N), synth[i]));
    xr0[FftIdx(i)] = v;                                                        // staged: the shift below crosses threads
  }
  __syncthreads();
  for (int i = tid; i < N; i += NT) wk[i] = i + P.hop < N ? xr0[FftIdx(i + P.hop)] : 0.0;
  for (int i = tid; i < P.hop; i += NT) {
    double v = xr0[FftIdx(i)];
    v = v > -32768.0 ? v : -32768.0;
    v = v < 32767.0 ? v : 32767.0;
    out[(size_t)slot * P.hop + i] = (int16_t)v;
  }
  if (tid == 0) hops[stream] = hop_index + 1;
}

// out = model hop, comfort-noise hop, or their raised-cosine cross-fade (lyra_decoder.cc:342-373); 320 threads per slot
__global__ void __launch_bounds__(320)
PlcMixKernel(const uint8_t* __restrict__ blob, CngParams P, int n, const uint8_t* __restrict__ plan, const int* __restrict__ fade0,
             const int* __restrict__ dir, const int16_t* __restrict__ model_pcm, const int16_t* __restrict__ cng_pcm,
             int16_t* __restrict__ out) {
  const int slot = (int)blockIdx.x, i = (int)threadIdx.x;
  if (slot >= n) return;
  const int pl = plan[slot];
  const size_t o = (size_t)slot * 320 + i;
  if (!(pl & 2)) { out[o] = model_pcm[o]; return; }
  if (!(pl & 1)) { out[o] = cng_pcm[o]; return; }
  const float* fade = BlobPtr<float>(blob, P.fade);
  const float w = fade[fade0[slot] + dir[slot] * i];
  const float v = __fadd_rn(__fmul_rn((float)model_pcm[o], w), __fmul_rn((float)cng_pcm[o], __fsub_rn(1.0f, w)));
  out[o] = (int16_t)v;
}

// read-only view of the noise estimators (NoiseEstimator::noise_estimate / is_noise, lyra/noise_estimator.h:55-62)
__global__ void __launch_bounds__(192)
NoiseReadKernel(const int* __restrict__ stream_ids, int n, const float* __restrict__ state, int nf, float* __restrict__ estimate_out,
                uint8_t* __restrict__ is_noise_out) {
  const int slot = (int)blockIdx.x, i = (int)threadIdx.x;
  if (slot >= n) return;
  const int stream = stream_ids ? stream_ids[slot] : slot;
  const float* st = state + (size_t)stream * NoiseStateUnits(nf);
  if (estimate_out && i < nf) estimate_out[(size_t)slot * nf + i] = st[i];
  if (is_noise_out && i == 0) is_noise_out[slot] = reinterpret_cast<const int*>(st + 5 * nf)[2] ? 0 : 1;
}

// Resampler::Resample (lyra/resampler.cc:52-57) for n streams, `n_in` input samples each: int16 -> float, polyphase FIR over the
// 35-tap delay line, ClipToInt16.  Arithmetic and tap order are oracle/resampler.c's (one separately rounded multiply and add per
// tap, ascending input order).  Per-stream state: delay[34] (the last 34 input samples), pos = position of the next output relative
// to the next input sample, in units of 1 / den input samples; rate = the external rate the state belongs to (a call at another
// rate starts from the fully-primed state, like a fresh Resampler).  counts[slot] = outputs produced (they differ by at most one
// between streams when down-sampling from different phases); out rows are `out_stride` samples apart.
__global__ void __launch_bounds__(128)
ResampleKernel(const uint8_t* __restrict__ blob, ResamplerParams P, int pair, int rate, const int* __restrict__ stream_ids, int n,
               const int16_t* __restrict__ in, int n_in, int16_t* __restrict__ out, int out_stride, int* __restrict__ counts,
               int16_t* __restrict__ delay_state, int* __restrict__ pos_state) {
  unsigned char* smem = LYRA_DYN_SMEM();
  float* x = reinterpret_cast<float*>(smem);              // [34 + n_in]: delay line followed by the new samples
  const int slot = (int)blockIdx.x;
  if (slot >= n) return;
  const int stream = stream_ids ? stream_ids[slot] : slot;
  constexpr int T = kResamplerTaps;
  const int tid = (int)threadIdx.x, NT = (int)blockDim.x;
  int16_t* dl = delay_state + (size_t)stream * (T - 1);
  int* ps = pos_state + (size_t)stream * 2;              // {pos, rate}
  const bool fresh = ps[1] != rate;
  const int a0 = fresh ? 0 : ps[0];
  for (int i = tid; i < T - 1; i += NT) x[i] = fresh ? 0.0f : (float)dl[i];
  for (int i = tid; i < n_in; i += NT) x[T - 1 + i] = (float)in[(size_t)slot * n_in + i];
  __syncthreads();
  const int num = P.num[pair], den = P.den[pair];
  const float* coeffs = BlobPtr<float>(blob, P.coeffs[pair]);
  const int total = n_in * den;
  const int count = a0 < total ? (total - a0 + num - 1) / num : 0;
  for (int j = tid; j < count && j < out_stride; j += NT) {
    const int a = a0 + j * num, i = a / den, ph = a % den;
    const float* c = coeffs + ph * T;
    const float* xs = x + i;                              // delay-line sample 0 of the window that ends at input sample i
    float acc = 0.0f;
#pragma unroll 5
    for (int t = 0; t < T; ++t) acc = __fadd_rn(acc, __fmul_rn(c[t], xs[t]));
    float v = acc;
    v = v > -32768.0f ? v : -32768.0f;
    v = v < 32767.0f ? v : 32767.0f;
    out[(size_t)slot * out_stride + j] = (int16_t)v;
  }
  __syncthreads();
  for (int i = tid; i < T - 1; i += NT) dl[i] = (int16_t)x[n_in + i];     // the last 34 samples of [delay | in]
  if (tid == 0) { ps[0] = a0 + count * num - total; ps[1] = rate; counts[slot] = count; }
}

}  // namespace lyra_b200
