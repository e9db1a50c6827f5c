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
                uint8_t* __restrict__ packets, int packet_bytes, int* __restrict__ indices_out) {
  unsigned char* smem = LYRA_DYN_SMEM();
  float* cbs = reinterpret_cast<float*>(smem);                                  // [2][64][16] stage codebooks (double buffer)
  float* rs = cbs + 2 * 1024;                                                   // [slots][64] residuals
  int* idxs = reinterpret_cast<int*>(rs + kRvqSlotsPerBlock * 64);              // [slots][48]
  const int tid = (int)threadIdx.x, grp = tid / 16, c = tid % 16;
  const int slot = (int)blockIdx.x * kRvqSlotsPerBlock + grp;
  const bool valid = slot < n;
  float* r = rs + grp * 64;
  int* idx = idxs + grp * 48;
  const float* cbt = BlobPtr<float>(blob, P.codebooks_t);
  // stage 0 codebook -> buffer 0 (4 KB = 256 x 16 B, two packets per thread)
  for (int i = tid; i < 256; i += kRvqThreads) lyra_cp_async16(cbs + 4 * i, cbt + 4 * i);
  lyra_cp_async_commit();
  for (int jj = 0; jj < 4; ++jj) r[c + 16 * jj] = valid ? features[(size_t)slot * 64 + c + 16 * jj] : 0.0f;
  for (int s = 0; s < nq; ++s) {
    lyra_cp_async_wait<0>();
    __syncthreads();                       // codebook s landed; everyone finished stage s-1 (its buffer is free again)
    if (s + 1 < nq) {
      float* dst = cbs + ((s + 1) & 1) * 1024;
      const float* src = cbt + (size_t)(s + 1) * 1024;
      for (int i = tid; i < 256; i += kRvqThreads) lyra_cp_async16(dst + 4 * i, src + 4 * i);
    }
    lyra_cp_async_commit();
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
      packets[(size_t)slot * packet_bytes + b] = (uint8_t)((hi << 4) | lo);
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
// One block per stream: periodic-Hann window over [previous hop, current hop], zero-padded radix-2
// FFT in double (same butterfly order and host-computed twiddles as the oracle), |X|, triangular mel
// weights accumulated in bin order, float cast, log(max(x, 500)) / 10.
// prev: [max_streams][window - hop] int16 carried samples (zero after reset).
__global__ void __launch_bounds__(256)
LogMelKernel(const uint8_t* __restrict__ blob, LogMelParams P, const int* __restrict__ stream_ids, int n,
             const int16_t* __restrict__ pcm, int16_t* __restrict__ prev, float* __restrict__ out) {
  unsigned char* smem = LYRA_DYN_SMEM();
  double* re = reinterpret_cast<double*>(smem);
  double* im = re + P.fft;
  double* mag = im + P.fft;            // [fft/2 + 1]
  const int slot = (int)blockIdx.x;
  if (slot >= n) return;
  const int stream = stream_ids ? stream_ids[slot] : slot;
  const int tid = (int)threadIdx.x, NT = (int)blockDim.x;
  const int carry = P.window_len - P.hop;
  const double* win = BlobPtr<double>(blob, P.window);
  const double* tw = BlobPtr<double>(blob, P.twiddle);
  int16_t* pv = prev + (size_t)stream * carry;
  const int16_t* cur = pcm + (size_t)slot * P.hop;
  int bits = 0;
  while ((1 << bits) < P.fft) ++bits;
  // windowed, zero-padded frame written in bit-reversed order
  for (int i = tid; i < P.fft; i += NT) {
    double v = 0.0;
    if (i < P.window_len) {
      const int16_t smp = i < carry ? pv[i] : cur[i - carry];
      v = __dmul_rn((double)smp, win[i]);
    }
    unsigned rev = 0;
    for (int b = 0; b < bits; ++b) rev |= ((unsigned)(i >> b) & 1u) << (bits - 1 - b);
    re[rev] = v;
    im[rev] = 0.0;
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
    __syncthreads();
  }
  for (int len = 2; len <= P.fft; len <<= 1) {
    const int half = len >> 1, step = P.fft / len;
    for (int i = tid; i < P.fft / 2; i += NT) {
      const int blk = i / half, k = i % half;
      const int a = blk * len + k, b = a + half;
      const double wr = tw[2 * (k * step)], wi = tw[2 * (k * step) + 1];
      const double xr = __dsub_rn(__dmul_rn(re[b], wr), __dmul_rn(im[b], wi));
      const double xi = __dadd_rn(__dmul_rn(re[b], wi), __dmul_rn(im[b], wr));
      const double ar = re[a], ai = im[a];
      re[b] = __dsub_rn(ar, xr); im[b] = __dsub_rn(ai, xi);
      re[a] = __dadd_rn(ar, xr); im[a] = __dadd_rn(ai, xi);
    }
    __syncthreads();
  }
  const int bins = P.fft / 2 + 1;
  for (int i = tid; i < bins; i += NT)
    mag[i] = __dsqrt_rn(__dadd_rn(__dmul_rn(re[i], re[i]), __dmul_rn(im[i], im[i])));
  __syncthreads();
  // each mel channel: bins of band ch-1 contribute (v - v*w), bins of band ch contribute v*w, in bin order
  const double* wts = BlobPtr<double>(blob, P.weights);
  const int* band = BlobPtr<int>(blob, P.band);
  for (int ch = tid; ch < P.num_mel; ch += NT) {
    double acc = 0.0;
    for (int i = P.start_index; i <= P.end_index; ++i) {
      const int bd = band[i];
      if (bd == ch) acc = __dadd_rn(acc, __dmul_rn(mag[i], wts[i]));
      else if (bd == ch - 1) acc = __dadd_rn(acc, __dsub_rn(mag[i], __dmul_rn(mag[i], wts[i])));
    }
    float v = (float)acc;
    v = v > 500.0f ? v : 500.0f;
    out[(size_t)slot * P.num_mel + ch] = __fdiv_rn(logf(v), 10.0f);
  }
}

}  // namespace lyra_b200
