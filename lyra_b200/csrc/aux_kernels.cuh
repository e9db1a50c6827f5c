// RVQ encode/decode + bit packing, state reset and the log-mel front end.
#pragma once

#include "kernel_prims.cuh"

namespace lyra_b200 {

// ------------------------------------------------------------------------------------------------
// Residual vector quantizer, encode side + Packet<184>::PackQuantized (0 header bits).
// Replaces ResidualVectorQuantizer::Quantize (lyra/residual_vector_quantizer.cc:77-110; the "encode"
// subgraph of quantizer.tflite) and Packet::Pack (lyra/packet.h:91-122).
// 16 lanes per stream (one per codeword), two streams per warp.  Per stage, exactly the graph's ops:
//   d[c] = sum_j (r[j] - cb[c][j])^2 (ascending j, each op rounded), argmin with lowest-index ties,
//   q = cb[best]; t = q - r; u = r + t; r = r - u.
// Only the first `nq` stages are evaluated: later stages never influence earlier indices.
constexpr int kRvqThreads = 128;
constexpr int kRvqSlotsPerBlock = kRvqThreads / 16;

__global__ void __launch_bounds__(kRvqThreads)
RvqEncodeKernel(const uint8_t* __restrict__ blob, RvqParams P, const float* __restrict__ features, int n, int nq,
                uint8_t* __restrict__ packets, int packet_bytes, int* __restrict__ indices_out, const uint8_t* __restrict__ skip) {
  unsigned char* smem = LYRA_DYN_SMEM();
  float* cbs = reinterpret_cast<float*>(smem);                                  // [2][64][16] stage codebooks (double buffer, 16-byte aligned)
  float* rs = cbs + 2 * 1024;                                                   // [slots][64] residuals
  int* idxs = reinterpret_cast<int*>(rs + kRvqSlotsPerBlock * 64);              // [slots][48]
  const int tid = (int)threadIdx.x, grp = tid / 16, c = tid % 16;
  const int slot = (int)blockIdx.x * kRvqSlotsPerBlock + grp;
  const bool valid = slot < n;
  const bool skipped = valid && skip != nullptr && skip[slot];     // DTX: the hop was noise, its packet is empty (bytes zeroed)
  float* r = rs + grp * 64;
  int* idx = idxs + grp * 48;
  const float* cbt = BlobPtr<float>(blob, P.codebooks_t);
  // per-stage codebooks (4 KB each) arrive by bulk asynchronous copy (TMA) into a double buffer; one mbarrier per buffer
  LYRA_STATIC_SMEM(LyraMbar, full, 2);
  if (tid == 0) {
    lyra_mbar_init(&full[0], 1);
    lyra_mbar_init(&full[1], 1);
    lyra_mbar_fence_init();
    lyra_bulk_g2s(cbs, cbt, 4096u, &full[0]);
  }
  for (int jj = 0; jj < 4; ++jj) r[c + 16 * jj] = valid ? features[(size_t)slot * 64 + c + 16 * jj] : 0.0f;
  for (int s = 0; s < nq; ++s) {
    __syncthreads();                       // everyone finished stage s-1: its buffer is free again (and the barriers are initialised)
    if (tid == 0 && s + 1 < nq) lyra_bulk_g2s(cbs + ((s + 1) & 1) * 1024, cbt + (size_t)(s + 1) * 1024, 4096u, &full[(s + 1) & 1]);
    lyra_mbar_wait(&full[s & 1], (unsigned)((s >> 1) & 1));     // codebook s has landed
    const float* cs = cbs + (s & 1) * 1024 + c;
    float d = 0.0f;
#pragma unroll 16
    for (int j = 0; j < 64; ++j) {
      const float df = __fsub_rn(r[j], cs[j * 16]);
      d = __fadd_rn(d, __fmul_rn(df, df));
    }
    int best = c;
#pragma unroll
    for (int m = 8; m >= 1; m >>= 1) {
      const float od = __shfl_xor_sync(0xffffffffu, d, m, 16);
      const int oi = __shfl_xor_sync(0xffffffffu, best, m, 16);
      if (od < d || (od == d && oi < best)) { d = od; best = oi; }
    }
    __syncwarp();
    const float* q = cbs + (s & 1) * 1024 + best;      // q[j] = codebook[best][j] at q[j * 16]
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const int j = c + 16 * jj;
      const float rj = r[j];
      const float t = __fsub_rn(q[j * 16], rj);
      const float u = __fadd_rn(rj, t);
      r[j] = __fsub_rn(rj, u);
    }
    if (c == 0) idx[s] = best;
  }
  __syncthreads();
  if (valid) {
    // first quantizer in the most significant bits (residual_vector_quantizer.cc:101-109), bytes MSB-first
    for (int b = c; b < packet_bytes; b += 16) {
      const int hi = idx[2 * b], lo = 2 * b + 1 < nq ? idx[2 * b + 1] : 0;
      packets[(size_t)slot * packet_bytes + b] = skipped ? (uint8_t)0 : (uint8_t)((hi << 4) | lo);
    }
    if (indices_out)
      for (int s = c; s < P.num_stages; s += 16) indices_out[(size_t)slot * P.num_stages + s] = s < nq ? idx[s] : -1;
  }
}

// ------------------------------------------------------------------------------------------------
// RVQ decode: Packet::UnpackPacket (lyra/packet.h:62-71,126-146) + DecodeToLossyFeatures
// (lyra/residual_vector_quantizer.cc:112-168, "decode" subgraph): left-to-right sum over all 46 stages,
// unused stages contribute codebook[0] * 0.  A stream whose packet was not received gets 64 zero
// features (ZeroFeatureEstimator, lyra/lyra_decoder.cc:317-326).
__global__ void __launch_bounds__(256)
RvqDecodeKernel(const uint8_t* __restrict__ blob, RvqParams P, const uint8_t* __restrict__ packets, int packet_bytes,
                const uint8_t* __restrict__ received, int n, int nq, float* __restrict__ features) {
  const int gid = (int)(blockIdx.x * blockDim.x + threadIdx.x);
  const int slot = gid / 64, j = gid % 64;
  if (slot >= n) return;
  float out = 0.0f;
  if (received == nullptr || received[slot]) {
    const float* cb = BlobPtr<float>(blob, P.codebooks);
    const uint8_t* pk = packets + (size_t)slot * packet_bytes;
    for (int k = 0; k < P.num_stages; ++k) {
      float t;
      if (k < nq) {
        const int byte = pk[k >> 1];
        const int idx = (k & 1) ? (byte & 15) : (byte >> 4);
        t = __fmul_rn(cb[((size_t)k * 16 + idx) * 64 + j], 1.0f);
      } else {
        t = __fmul_rn(cb[((size_t)k * 16) * 64 + j], 0.0f);
      }
      out = k == 0 ? t : __fadd_rn(out, t);
    }
  }
  features[(size_t)slot * 64 + j] = out;
}

// ------------------------------------------------------------------------------------------------
// Reset the streaming state of selected streams to the reference's CALL_ONCE initial values
// (all-zero resource variables; int8 rings hold the zero point of their tensor).
__global__ void __launch_bounds__(256)
ResetStateKernel(uint32_t* __restrict__ state, const uint32_t* __restrict__ init, int units, int S,
                 const int* __restrict__ streams, int nstreams, int* __restrict__ n18) {
  const int k = (int)blockIdx.x;
  if (k >= nstreams) return;
  const int stream = streams ? streams[k] : k;
  const int tile = stream / S, lane = stream % S;
  uint32_t* st = state + (size_t)tile * units * S + lane;
  for (int u = (int)threadIdx.x; u < units; u += (int)blockDim.x) st[(size_t)u * S] = init[u];
  if (threadIdx.x == 0) n18[stream] = 0;
}

// ------------------------------------------------------------------------------------------------
// Log-mel spectrogram (LogMelSpectrogramExtractorImpl::Extract, lyra/log_mel_spectrogram_extractor_impl.cc:96-126).
// One block of 128 threads per stream: periodic-Hann window over [previous hop, current hop], zero-padded 1024-point
// radix-2 decimation-in-time FFT in double — the same butterflies, operand order and host-computed twiddles as the
// oracle, so the spectrum is bit-identical to it — |X|, triangular mel weights accumulated in bin order, float cast,
// log(max(x, 500)) / 10 with the logarithm taken in double and rounded once (equal to a correctly rounded logf).
// The ten stages run as 3 + 3 + 3 + 1: a thread keeps 8 points in registers across three consecutive stages (points
// base + j * STRIDE, the closed set of three stages whose half-lengths are STRIDE, 2 STRIDE, 4 STRIDE), so the block
// synchronises four times instead of ten.
// prev: [max_streams][window - hop] int16 carried samples (zero after reset).
constexpr int kLogMelThreads = 128;
constexpr int kLogMelFft = 1024;

// butterfly of stage `len` on (a, b = a + len/2) with twiddle w = tw[k * (fft / len)]:  x = b * w;  b = a - x;  a = a + x
__device__ __forceinline__ void FftButterfly(double& ar, double& ai, double& br, double& bi, double wr, double wi) {
  const double xr = __dsub_rn(__dmul_rn(br, wr), __dmul_rn(bi, wi));
  const double xi = __dadd_rn(__dmul_rn(br, wi), __dmul_rn(bi, wr));
  br = __dsub_rn(ar, xr); bi = __dsub_rn(ai, xi);
  ar = __dadd_rn(ar, xr); ai = __dadd_rn(ai, xi);
}

// shared-memory index of FFT point i: one pad element per 8 keeps the stride-8 and stride-64 register-blocked passes
// (nearly) free of bank conflicts
__device__ __forceinline__ int FftIdx(int i) { return i + (i >> 3); }
constexpr int kLogMelFftPadded = kLogMelFft + kLogMelFft / 8;

// three consecutive stages (half-lengths STRIDE, 2 STRIDE, 4 STRIDE) on the 8 points base + j * STRIDE; r = base mod STRIDE.
// tw: per-stage twiddle tables, the stage of half-length h starts at complex entry h - 1.
template <int STRIDE>
__device__ __forceinline__ void FftButterflies3(double (&xr)[8], double (&xi)[8], int r, const double2* __restrict__ tw) {
#pragma unroll
  for (int s = 0; s < 3; ++s) {
    const int half = STRIDE << s;
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (!(j & (1 << s))) {
        const int k = r + (j & ((1 << s) - 1)) * STRIDE;
        const double2 w = tw[half - 1 + k];
        FftButterfly(xr[j], xi[j], xr[j + (1 << s)], xi[j + (1 << s)], w.x, w.y);
      }
  }
}

template <int STRIDE>
__device__ __forceinline__ void FftStages3(double* re, double* im, int base, int r, const double2* __restrict__ tw) {
  double xr[8], xi[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { xr[j] = re[FftIdx(base + j * STRIDE)]; xi[j] = im[FftIdx(base + j * STRIDE)]; }
  FftButterflies3<STRIDE>(xr, xi, r, tw);
#pragma unroll
  for (int j = 0; j < 8; ++j) { re[FftIdx(base + j * STRIDE)] = xr[j]; im[FftIdx(base + j * STRIDE)] = xi[j]; }
}

__global__ void __launch_bounds__(kLogMelThreads)
LogMelKernel(const uint8_t* __restrict__ blob, LogMelParams P, const int* __restrict__ stream_ids, int n,
             const int16_t* __restrict__ pcm, int16_t* __restrict__ prev, float* __restrict__ out,
             const uint8_t* __restrict__ mask, int slot_base) {
  unsigned char* smem = LYRA_DYN_SMEM();
  double* re = reinterpret_cast<double*>(smem);
  double* im = re + kLogMelFftPadded;
  double* mag = im + kLogMelFftPadded;       // [fft/2 + 1]
  double* xw = mag + kLogMelFft / 2 + 1;     // [window_len, padded like the FFT buffers] windowed samples in natural order
  const int slot = slot_base + (int)blockIdx.x;     // I/O arrays are indexed by slot; a sub-batch starts at slot_base
  if (slot >= n) return;
  if (mask && !mask[slot]) return;     // this stream's extractor is not fed this hop (its carried samples stay)
  const int stream = stream_ids ? stream_ids[slot] : slot;
  const int tid = (int)threadIdx.x;
  constexpr int NT = kLogMelThreads;
  const int carry = P.window_len - P.hop;
  const double* win = BlobPtr<double>(blob, P.window);
  const double2* tw = BlobPtr<double2>(blob, P.twiddle);
  int16_t* pv = prev + (size_t)stream * carry;
  const int16_t* cur = pcm + (size_t)slot * P.hop;
  // windowed samples in natural order (the zero padding up to 1024 points is implicit)
  for (int i = tid; i < P.window_len; i += NT) {
    const int16_t smp = i < carry ? pv[i] : cur[i - carry];
    xw[FftIdx(i)] = __dmul_rn((double)smp, win[i]);
  }
  __syncthreads();
  // the carried samples become the tail of (previous carry, current hop); staged through `mag`
  {
    int16_t* stage = reinterpret_cast<int16_t*>(mag);
    for (int i = tid; i < carry; i += NT) {
      const int src = i + P.hop;
      stage[i] = src < carry ? pv[src] : cur[src - carry];
    }
    __syncthreads();
    for (int i = tid; i < carry; i += NT) pv[i] = stage[i];
  }
  {                                                                            // stages 2, 4, 8 on points 8 tid .. 8 tid + 7
    double xr[8], xi[8];
    const int b7 = (int)(__brev((unsigned)tid) >> 25);                         // bit-reversed position 8t + j <-> natural index brev7(t) + 128 brev3(j)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int j3 = ((j & 1) << 2) | (j & 2) | ((j & 4) >> 2);
      const int i = b7 + 128 * j3;
      xr[j] = i < P.window_len ? xw[FftIdx(i)] : 0.0;
      xi[j] = 0.0;
    }
    FftButterflies3<1>(xr, xi, 0, tw);
#pragma unroll
    for (int j = 0; j < 8; ++j) { re[FftIdx(8 * tid + j)] = xr[j]; im[FftIdx(8 * tid + j)] = xi[j]; }
  }
  __syncthreads();
  FftStages3<8>(re, im, 64 * (tid / 8) + tid % 8, tid % 8, tw);                // stages 16, 32, 64
  __syncthreads();
  FftStages3<64>(re, im, 512 * (tid / 64) + tid % 64, tid % 64, tw);           // stages 128, 256, 512
  __syncthreads();
#pragma unroll
  for (int m = 0; m < 4; ++m) {                                                // stage 1024
    const int a = tid + NT * m, ia = FftIdx(a), ib = FftIdx(a + 512);
    const double2 w = tw[511 + a];
    FftButterfly(re[ia], im[ia], re[ib], im[ib], w.x, w.y);
  }
  __syncthreads();
  const int bins = kLogMelFft / 2 + 1;
  for (int i = tid; i < bins; i += NT)
    mag[i] = __dsqrt_rn(__dadd_rn(__dmul_rn(re[FftIdx(i)], re[FftIdx(i)]), __dmul_rn(im[FftIdx(i)], im[FftIdx(i)])));
  __syncthreads();
  // each mel channel: bins of band ch-1 contribute (v - v*w), bins of band ch contribute v*w, in bin order
  const double* wts = BlobPtr<double>(blob, P.weights);
  const int* band = BlobPtr<int>(blob, P.band);
  const int* range = BlobPtr<int>(blob, P.range);
  for (int ch = tid; ch < P.num_mel; ch += NT) {
    double acc = 0.0;
    const int lo = range[2 * ch], hi = range[2 * ch + 1];
    for (int i = lo; i <= hi; ++i) {
      if (band[i] == ch) acc = __dadd_rn(acc, __dmul_rn(mag[i], wts[i]));
      else acc = __dadd_rn(acc, __dsub_rn(mag[i], __dmul_rn(mag[i], wts[i])));
    }
    float v = (float)acc;
    v = v > 500.0f ? v : 500.0f;
    out[(size_t)slot * P.num_mel + ch] = __fdiv_rn((float)log((double)v), 10.0f);
  }
}

// ------------------------------------------------------------------------------------------------
// Minimum-statistics noise estimator on the decoder output:
// NoiseEstimator::ReceiveSamples for whole hops, after the log-mel kernel has produced this hop's
// 160-bin spectrum (lyra/noise_estimator.cc:144-245; SmoothingFactor / UpdateMinAndTemp :37-95).
// One block per stream, one thread per mel bin.  Per-stream state: est | bound | smoothed | squared-smoothed |
// tmp-min (nf floats each) followed by 4 ints {has_smoothed, hops_received, last_hop_was_not_noise, -} — all-zero
// is the reference's freshly constructed object (noise estimate and bound 0, is_noise() true).
// Arithmetic follows the C++ expression types: float ops rounded one by one, the bound in double
// (std::log of an integer is double), std::exp(float) as (float)exp((double)x) like the oracle.
// Streams whose mask byte is 0 are left untouched (LyraDecoder only feeds hops decoded from a received
// packet, lyra/lyra_decoder.cc:306-311) but still report their current is_noise / noise_estimate.
struct NoiseParams { int nf, hops_per_update; float max_smoothing, bound_decay; double log_nf; };
constexpr int kNoiseThreads = 192;
__host__ __device__ constexpr int NoiseStateUnits(int nf) { return 5 * nf + 4; }

__global__ void __launch_bounds__(kNoiseThreads)
NoiseEstimatorKernel(NoiseParams P, const int* __restrict__ stream_ids, int n, const float* __restrict__ mel,
                     const uint8_t* __restrict__ mask, float* __restrict__ state, uint8_t* __restrict__ is_noise_out,
                     float* __restrict__ estimate_out, int slot_base) {
  unsigned char* smem = LYRA_DYN_SMEM();
  float* cur = reinterpret_cast<float*>(smem);       // [nf]
  float* sm = cur + P.nf;                             // [nf] smoothed power before this update
  float* red = sm + P.nf;                             // [0] smoothing correction
  int* flag = reinterpret_cast<int*>(red + 1);
  const int slot = slot_base + (int)blockIdx.x;
  if (slot >= n) return;
  const int stream = stream_ids ? stream_ids[slot] : slot;
  const int i = (int)threadIdx.x, nf = P.nf;
  float* st = state + (size_t)stream * NoiseStateUnits(nf);
  float* est = st;
  float* bound = st + nf;
  float* smoothed = st + 2 * nf;
  float* sq = st + 3 * nf;
  float* tmp_min = st + 4 * nf;
  int* meta = reinterpret_cast<int*>(st + 5 * nf);
  const bool feed = mask == nullptr || mask[slot] != 0;
  if (!feed) {
    if (is_noise_out && i == 0) is_noise_out[slot] = meta[2] ? 0 : 1;
    if (estimate_out && i < nf) estimate_out[(size_t)slot * nf + i] = est[i];
    return;
  }
  const int has = meta[0], hops = meta[1];
  if (i == 0) *flag = 0;
  __syncthreads();
  float c = 0.0f, e = 0.0f, b = 0.0f;
  if (i < nf) {
    c = mel[(size_t)slot * nf + i];
    e = est[i];
    b = bound[i];
    cur[i] = c;
    if (fabsf(__fsub_rn(c, e)) > b) *flag = 1;       // ComputeIsNoise: any bin outside estimate +- bound
  }
  __syncthreads();
  const bool not_noise = *flag != 0;
  if (!not_noise) {
    if (i < nf) bound[i] = __fmul_rn(b, P.bound_decay);         // DecayBounds
  } else {
    float s = 0.0f, q = 0.0f, tm = 0.0f;
    if (i < nf) {
      s = has ? smoothed[i] : c;
      q = has ? sq[i] : __fmul_rn(c, c);
      tm = has ? tmp_min[i] : c;
      sm[i] = s;
    }
    __syncthreads();
    if (i == 0) {
      float a0 = 0.0f, a1 = 0.0f;                      // Average(): std::accumulate in index order, then / size
      for (int k = 0; k < nf; ++k) { a0 = __fadd_rn(a0, sm[k]); a1 = __fadd_rn(a1, cur[k]); }
      const float d = __fdiv_rn(__fsub_rn(__fdiv_rn(a0, (float)nf), __fdiv_rn(a1, (float)nf)), 0.3f);
      red[0] = (float)exp((double)(-__fmul_rn(d, d)));
    }
    __syncthreads();
    if (i < nf) {
      const float r = __fdiv_rn(__fsub_rn(s, e), 0.3f);
      const float sf = __fmul_rn(__fmul_rn(P.max_smoothing, red[0]), (float)exp((double)(-__fmul_rn(r, r))));
      const float om = __fsub_rn(1.0f, sf);
      const float ns = __fadd_rn(__fmul_rn(sf, s), __fmul_rn(om, c));
      const float nq = __fadd_rn(__fmul_rn(sf, q), __fmul_rn(om, __fmul_rn(c, c)));
      float ne, nt;
      if (hops == 0) { ne = ns < tm ? ns : tm; nt = ns; }          // UpdateMinAndTemp
      else { ne = ns < e ? ns : e; nt = ns < tm ? ns : tm; }
      const float t = __fsub_rn(nq, __fmul_rn(ns, ns));
      const float var = t > 0.0f ? t : 0.0f;
      const float nb = (float)__dmul_rn((double)0.9f, __dsqrt_rn(__dmul_rn((double)var, P.log_nf)));   // ComputeBounds
      smoothed[i] = ns; sq[i] = nq; tmp_min[i] = nt; est[i] = ne; bound[i] = nb;
      e = ne;
    }
    if (i == 0) { meta[0] = 1; meta[1] = (hops + 1) % P.hops_per_update; }
  }
  if (i == 0) {
    meta[2] = not_noise ? 1 : 0;
    if (is_noise_out) is_noise_out[slot] = not_noise ? 0 : 1;
  }
  if (estimate_out && i < nf) estimate_out[(size_t)slot * nf + i] = e;
}

}  // namespace lyra_b200
