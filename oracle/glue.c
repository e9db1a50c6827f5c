/*
 * ORACLE — TEST INFRASTRUCTURE ONLY (see lyra_oracle.h).
 *
 * The C++ glue around the three models, restated in C:
 *   Packet<184>::{PackQuantized,UnpackPacket,PacketSize}     lyra/packet.h:56-146
 *   Int16ToUnitScalar / UnitToInt16Scalar / ClipToInt16Scalar lyra/dsp_utils.h:53-60,79-88,104-108
 *   LogSpectralDistance                                       lyra/dsp_utils.cc:27-41
 *   LyraEncoder::Encode (16 kHz, no DTX)                      lyra/lyra_encoder.cc:113-156
 *   LyraDecoder::SetEncodedPacket + DecodeSamples(hop)        lyra/lyra_decoder.cc:172-226,317-326
 *   lyra_benchmark stage split                                lyra/lyra_benchmark_lib.cc:85-160
 */
#define _GNU_SOURCE
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "lyra_oracle.h"

/* ------------------------------------------------------------------ packet ---- */

int lo_packet_size(int num_header_bits, int num_quantized_bits) {
  return (int)ceilf((float)(num_quantized_bits + num_header_bits) / 8.0f);      /* packet.h:73-76 */
}

int lo_packet_pack(const char* bits, int nh, int nq, uint8_t* bytes) {
  /* Packet<MaxNumPacketBits>::Create rejects header + payload > MaxNumPacketBits (packet.h:41-48); Lyra
     instantiates 184 (lyra_components.cc:33,57-60), packet_test.cc up to 204: the cap here is the buffer's */
  if (nh + nq > 256 || nh < 0 || nq < 0) return -1;
  const int total = nh + nq, nb = lo_packet_size(nh, nq);
  memset(bytes, 0, (size_t)nb);
  /* header bits are all zero today (SetHeader, packet.h:160-170); payload follows MSB-first; the
     unused low bits of the last byte stay zero (packet.h:106-117) */
  for (int i = 0; i < nq; ++i)
    if (bits[i] == '1') { const int pos = nh + i; bytes[pos >> 3] |= (uint8_t)(0x80u >> (pos & 7)); }
  (void)total;
  return nb;
}

int lo_packet_unpack(const uint8_t* bytes, int nbytes, int nh, int nq, char* bits_out) {
  if (nbytes != lo_packet_size(nh, nq)) return -1;                              /* packet.h:64-67 */
  for (int i = 0; i < nq; ++i) { const int pos = nh + i; bits_out[i] = (bytes[pos >> 3] >> (7 - (pos & 7))) & 1 ? '1' : '0'; }
  bits_out[nq] = 0;
  return nq;
}

/* ------------------------------------------------------------------ dsp ---- */

float lo_int16_to_unit(int16_t v) { return -(float)v / (float)(-32768); }
int16_t lo_unit_to_int16(float v) {
  float s = v * 32768.0f;
  s = s > -32768.0f ? s : -32768.0f;      /* std::max(value, int16 min) */
  s = s < 32767.0f ? s : 32767.0f;        /* std::min(value, int16 max) */
  return (int16_t)s;                      /* implicit float -> int16_t conversion: truncation */
}
float lo_log_spectral_distance(const float* a, const float* b, int n) {
  float acc = 0.0f;
  for (int i = 0; i < n; ++i) { const float d = a[i] - b[i]; acc += d * d; }
  return 10.0f * sqrtf(acc / (float)n);
}

/* ------------------------------------------------------------------ one-stream codec ---- */

struct lo_codec {
  lo_net* enc;
  lo_net* dec;
  lo_rvq* rvq;
};

lo_codec* lo_codec_create(const char* model_dir) {
  char path[2048];
  lo_codec* c = (lo_codec*)calloc(1, sizeof(*c));
  snprintf(path, sizeof(path), "%s/soundstream_encoder.tflite", model_dir);
  c->enc = lo_net_create(path);
  snprintf(path, sizeof(path), "%s/lyragan.tflite", model_dir);
  c->dec = lo_net_create(path);
  snprintf(path, sizeof(path), "%s/quantizer.tflite", model_dir);
  c->rvq = lo_rvq_create(path);
  if (!c->enc || !c->dec || !c->rvq) { lo_codec_free(c); return NULL; }
  return c;
}
void lo_codec_free(lo_codec* c) {
  if (!c) return;
  if (c->enc) lo_net_free(c->enc);
  if (c->dec) lo_net_free(c->dec);
  if (c->rvq) lo_rvq_free(c->rvq);
  free(c);
}
int lo_codec_reset(lo_codec* c) { return lo_net_reset(c->enc) | lo_net_reset(c->dec); }
lo_net* lo_codec_encoder_net(lo_codec* c) { return c->enc; }
lo_net* lo_codec_decoder_net(lo_codec* c) { return c->dec; }

int lo_codec_encode(lo_codec* c, const int16_t* pcm, int num_bits, uint8_t* packet, float* features64, int32_t* indices46) {
  float in[320], feat[64];
  char bits[192];
  for (int i = 0; i < 320; ++i) in[i] = lo_int16_to_unit(pcm[i]);               /* soundstream_encoder.cc:55-57 */
  if (lo_net_invoke(c->enc, in, 320, feat, 64) != 0) return -1;
  if (features64) memcpy(features64, feat, sizeof(feat));
  if (lo_rvq_quantize_bits(c->rvq, feat, num_bits, bits) != 0) return -1;
  if (indices46) lo_rvq_encode(c->rvq, feat, num_bits / 4, indices46);
  return lo_packet_pack(bits, 0, num_bits, packet);
}

int lo_codec_decode(lo_codec* c, const uint8_t* packet, int num_bits, int16_t* pcm, float* lossy64, float* unit320) {
  float feat[64], out[320];
  if (packet) {
    char bits[192];
    if (lo_packet_unpack(packet, lo_packet_size(0, num_bits), 0, num_bits, bits) < 0) return -1;
    if (lo_rvq_decode_bits(c->rvq, bits, num_bits, feat) != 0) return -1;
  } else {
    for (int i = 0; i < 64; ++i) feat[i] = 0.0f;                                /* ZeroFeatureEstimator */
  }
  if (lossy64) memcpy(lossy64, feat, sizeof(feat));
  if (lo_net_invoke(c->dec, feat, 64, out, 320) != 0) return -1;
  if (unit320) memcpy(unit320, out, sizeof(out));
  for (int i = 0; i < 320; ++i) pcm[i] = lo_unit_to_int16(out[i]);              /* lyra_gan_model.cc:60-64 */
  return 0;
}

/* ------------------------------------------------------------------ CPU baseline ---- */

typedef struct {
  const char* model_dir;
  int streams, frames, num_bits;
  uint32_t seed;
  int* next_stream;
  pthread_mutex_t* mu;
  double stage_s[4];
  long frames_done;
  uint64_t checksum;
  int error;
} bench_arg;

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* the synthetic PCM of bench.py: xorshift32 per stream, uniform in [-8192, 8192) (0.25 full scale,
   cf. the reference benchmark's uniform random audio, lyra_benchmark_lib.cc:233-239) */
static inline uint32_t xs32(uint32_t* s) { uint32_t x = *s; x ^= x << 13; x ^= x >> 17; x ^= x << 5; *s = x; return x; }

static void* bench_worker(void* p) {
  bench_arg* a = (bench_arg*)p;
  lo_codec* c = lo_codec_create(a->model_dir);
  if (!c) { a->error = 1; return NULL; }
  for (;;) {
    pthread_mutex_lock(a->mu);
    const int s = (*a->next_stream)++;
    pthread_mutex_unlock(a->mu);
    if (s >= a->streams) break;
    lo_codec_reset(c);
    uint32_t rng = a->seed + (uint32_t)s;
    if (rng == 0) rng = 1;
    for (int f = 0; f < a->frames; ++f) {
      int16_t pcm[320], outpcm[320];
      float in[320], feat[64], lossy[64], out[320];
      char bits[192], bits2[192];
      uint8_t packet[24];
      for (int i = 0; i < 320; ++i) pcm[i] = (int16_t)((int)(xs32(&rng) & 0x3FFF) - 8192);
      double t0 = now_s();
      for (int i = 0; i < 320; ++i) in[i] = lo_int16_to_unit(pcm[i]);
      lo_net_invoke(c->enc, in, 320, feat, 64);
      double t1 = now_s();
      lo_rvq_quantize_bits(c->rvq, feat, a->num_bits, bits);
      const int nb = lo_packet_pack(bits, 0, a->num_bits, packet);
      double t2 = now_s();
      lo_packet_unpack(packet, nb, 0, a->num_bits, bits2);
      lo_rvq_decode_bits(c->rvq, bits2, a->num_bits, lossy);
      double t3 = now_s();
      lo_net_invoke(c->dec, lossy, 64, out, 320);
      for (int i = 0; i < 320; ++i) outpcm[i] = lo_unit_to_int16(out[i]);
      double t4 = now_s();
      a->stage_s[0] += t1 - t0; a->stage_s[1] += t2 - t1; a->stage_s[2] += t3 - t2; a->stage_s[3] += t4 - t3;
      a->frames_done++;
      for (int i = 0; i < nb; ++i) a->checksum = a->checksum * 1099511628211ull + packet[i];
      for (int i = 0; i < 320; i += 37) a->checksum = a->checksum * 1099511628211ull + (uint16_t)outpcm[i];
    }
  }
  lo_codec_free(c);
  return NULL;
}

double lo_cpu_bench(const char* model_dir, int streams, int frames, int num_bits, int threads,
                    uint32_t seed, double* stage_us, uint64_t* checksum) {
  if (threads < 1) threads = 1;
  if (threads > 512) threads = 512;
  pthread_t th[512];
  bench_arg* args = (bench_arg*)calloc((size_t)threads, sizeof(bench_arg));
  pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
  int next = 0;
  const double t0 = now_s();
  for (int i = 0; i < threads; ++i) {
    args[i].model_dir = model_dir; args[i].streams = streams; args[i].frames = frames;
    args[i].num_bits = num_bits; args[i].seed = seed; args[i].next_stream = &next; args[i].mu = &mu;
    pthread_create(&th[i], NULL, bench_worker, &args[i]);
  }
  for (int i = 0; i < threads; ++i) pthread_join(th[i], NULL);
  const double wall = now_s() - t0;
  double st[4] = {0, 0, 0, 0};
  long done = 0;
  uint64_t cs = 0;
  int err = 0;
  for (int i = 0; i < threads; ++i) {
    for (int k = 0; k < 4; ++k) st[k] += args[i].stage_s[k];
    done += args[i].frames_done;
    cs ^= args[i].checksum;
    err |= args[i].error;
  }
  free(args);
  if (err || done == 0) return -1.0;
  if (stage_us) for (int k = 0; k < 4; ++k) stage_us[k] = 1e6 * st[k] / (double)done;
  if (checksum) *checksum = cs;
  return wall;
}
