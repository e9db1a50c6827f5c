/* TEST INFRASTRUCTURE ONLY (see lyra_oracle.h).
 *
 * CPU restatement of the reference's decoder-side control flow at the internal rate (16 kHz; the external resampler is the
 * identity there, lyra/buffered_resampler.cc:117-123) and of the encoder's DTX branch:
 *   LyraDecoder::SetEncodedPacket          lyra/lyra_decoder.cc:172-209
 *   LyraDecoder::DecodeSamplesInternal     lyra/lyra_decoder.cc:228-315   (GetNumSamplesToGenerate :65-93)
 *   RunGenerativeModel / RunComfortNoiseGenerator   lyra/lyra_decoder.cc:317-340
 *   MaybeOverlapAndInsert (raised-cosine cross-fade) lyra/lyra_decoder.cc:342-373
 *   is_comfort_noise                       lyra/lyra_decoder.cc:381-383
 *   ZeroFeatureEstimator                   lyra/zero_feature_estimator.h:28-41
 *   LyraEncoder::Encode with enable_dtx    lyra/lyra_encoder.cc:113-156
 * The components are pluggable the way the reference's constructor takes interfaces: the real ones (LyraGAN net, comfort-noise
 * generator, RVQ, noise estimator of this oracle) or the fakes of the reference's own mock-driven tests (constant-valued
 * generative models, testing/mock_generative_model.h:33-53; fixed lossy features and noise estimate,
 * lyra_decoder_test.cc:143-151), with call counters standing in for gmock's expectations.  tests/test_oracle_decoder.py
 * re-runs the state-machine cases of lyra/lyra_decoder_test.cc against this file - that is what pins it.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "lyra_oracle.h"

#define HOP 320
#define NF 64
#define NMEL 160
#define CONCEAL_SAMPLES 1280     /* 0.08 s * 16000 (lyra_decoder.cc:42-51) */
#define FADE_SAMPLES 640         /* 0.04 s * 16000 (lyra_decoder.cc:55-63) */

struct lo_decoder {
  lo_gen* model;            /* generative_model_ */
  lo_gen* cng;              /* comfort_noise_generator_ */
  lo_gen fake_model, fake_cng;
  int16_t fake_model_value, fake_cng_value;
  int use_fakes;
  /* real components */
  lo_net* net;
  lo_rvq* rvq;
  lo_cng* cng_real;
  lo_noise* noise;
  lo_gen real_model;
  /* fakes of the other collaborators */
  float fake_features[NF], fake_noise[NMEL];
  /* state (lyra_decoder.h:146-160) */
  int concealment_progress, fade_progress, fade_direction;     /* direction: -1 = kFadeFromCNG, +1 = kFadeToCNG */
  /* call counters (stand-ins for gmock expectations) */
  int n_vq_decode, n_noise_receive, n_noise_estimate;
};

static int fake_conditioning_model(void* self, const float* f, int16_t* hop) {
  (void)f;
  lo_decoder* d = (lo_decoder*)self;
  for (int i = 0; i < HOP; ++i) hop[i] = d->fake_model_value;
  return 0;
}
static int fake_conditioning_cng(void* self, const float* f, int16_t* hop) {
  (void)f;
  lo_decoder* d = (lo_decoder*)self;
  for (int i = 0; i < HOP; ++i) hop[i] = d->fake_cng_value;
  return 0;
}
/* LyraGanModel::RunConditioning (lyra_gan_model.cc:53-58): the whole hop is generated at once; float -> int16 as
 * UnitToInt16 (dsp_utils.h:79-88) */
static int real_conditioning_model(void* self, const float* f, int16_t* hop) {
  lo_decoder* d = (lo_decoder*)self;
  float out[HOP];
  if (lo_net_invoke(d->net, f, NF, out, HOP) != 0) return -1;
  for (int i = 0; i < HOP; ++i) hop[i] = lo_unit_to_int16(out[i]);
  return 0;
}

static lo_decoder* decoder_alloc(void) {
  lo_decoder* d = (lo_decoder*)calloc(1, sizeof(*d));
  d->fade_direction = -1;                                       /* kFadeFromCNG (lyra_decoder.cc:166) */
  return d;
}

lo_decoder* lo_decoder_create(const char* model_dir, uint64_t cng_seed) {
  lo_decoder* d = decoder_alloc();
  char path[1024];
  snprintf(path, sizeof(path), "%s/lyragan.tflite", model_dir);
  d->net = lo_net_create(path);
  snprintf(path, sizeof(path), "%s/quantizer.tflite", model_dir);
  d->rvq = lo_rvq_create(path);
  d->cng_real = lo_cng_create(16000, HOP, 640, NMEL, cng_seed);
  d->noise = lo_noise_create(16000, HOP, 640, NMEL);
  if (!d->net || !d->rvq || !d->cng_real || !d->noise) { lo_decoder_free(d); return NULL; }
  lo_gen_init(&d->real_model, HOP, NF, real_conditioning_model, d);
  d->model = &d->real_model;
  d->cng = lo_cng_gen(d->cng_real);
  return d;
}

lo_decoder* lo_decoder_create_fake(int16_t model_value, int16_t cng_value) {
  lo_decoder* d = decoder_alloc();
  d->use_fakes = 1;
  d->fake_model_value = model_value;
  d->fake_cng_value = cng_value;
  lo_gen_init(&d->fake_model, HOP, NF, fake_conditioning_model, d);
  lo_gen_init(&d->fake_cng, HOP, NMEL, fake_conditioning_cng, d);
  d->model = &d->fake_model;
  d->cng = &d->fake_cng;
  for (int i = 0; i < NF; ++i) d->fake_features[i] = (float)i;              /* std::iota(mock_features_, 0) */
  for (int i = 0; i < NMEL; ++i) d->fake_noise[i] = 10.0f + (float)i;       /* std::iota(mock_noise_features_, 10.f) */
  return d;
}

void lo_decoder_free(lo_decoder* d) {
  if (!d) return;
  if (d->use_fakes) { lo_gen_free(&d->fake_model); lo_gen_free(&d->fake_cng); }
  else {
    if (d->real_model.queue) lo_gen_free(&d->real_model);
    lo_net_free(d->net); lo_rvq_free(d->rvq); lo_cng_free(d->cng_real); lo_noise_free(d->noise);
  }
  free(d);
}

static int packet_size_to_bits(int nbytes) { return nbytes == 8 ? 64 : nbytes == 15 ? 120 : nbytes == 23 ? 184 : -1; }   /* lyra_config.h:100-115 */

int lo_decoder_set_encoded_packet(lo_decoder* d, const uint8_t* encoded, int nbytes) {
  const int bits = packet_size_to_bits(nbytes);
  if (bits < 0) return -1;                                                   /* :173-178 */
  char str[185];
  if (lo_packet_unpack(encoded, nbytes, 0, bits, str) < 0) return -1;       /* :179-184 */
  /* finish playing out any concealment or comfort noise packet first (:186-196) */
  if (d->concealment_progress == CONCEAL_SAMPLES) d->concealment_progress = -lo_gen_num_samples_available(d->cng);
  else if (d->concealment_progress > 0) d->concealment_progress = -lo_gen_num_samples_available(d->model);
  float features[NF];
  d->n_vq_decode++;
  if (d->use_fakes) memcpy(features, d->fake_features, sizeof(features));
  else if (lo_rvq_decode_bits(d->rvq, str, bits, features) != 0) return -1; /* :198-202 */
  if (lo_gen_add_features(d->model, features, NF) != 0) return -1;          /* :203-206 */
  /* feature_estimator_->Update: ZeroFeatureEstimator ignores it (:207) */
  return 0;
}

/* GetNumSamplesToGenerate (lyra_decoder.cc:65-93) */
static int num_samples_to_generate(int requested, int so_far, int concealment_progress, int model_avail, int cng_avail) {
  int remaining;
  if (concealment_progress < 0) remaining = abs(concealment_progress);
  else if (concealment_progress < CONCEAL_SAMPLES) remaining = model_avail % HOP;
  else remaining = cng_avail;
  if (remaining == 0) remaining = HOP;
  const int left = requested - so_far;
  return left < remaining ? left : remaining;
}

int lo_decoder_decode_samples(lo_decoder* d, int num_samples, int16_t* out) {
  if (num_samples < 0) return -1;
  int16_t audio[HOP], noise_hop[HOP];
  int produced = 0;
  while (produced < num_samples) {                                           /* :232 */
    const int n = num_samples_to_generate(num_samples, produced, d->concealment_progress,
                                          lo_gen_num_samples_available(d->model), lo_gen_num_samples_available(d->cng));
    const int is_packet_received = lo_gen_num_samples_available(d->model) > 0 && d->concealment_progress == 0;   /* :249-251 */
    if (is_packet_received) d->fade_direction = -1;                         /* :253-256 */
    else if (d->concealment_progress == CONCEAL_SAMPLES) d->fade_direction = +1;                                /* :257-260 */
    else d->concealment_progress += n;                                      /* :261-265 */
    int cng_n = n, gen_n = n;
    int next_fade = d->fade_progress + d->fade_direction * n;               /* :269-270 */
    if (d->fade_direction == +1 && d->fade_progress == FADE_SAMPLES) { next_fade = FADE_SAMPLES; gen_n = 0; }   /* :271-276 */
    else if (d->fade_direction == -1 && d->fade_progress == 0) { next_fade = 0; cng_n = 0; }                    /* :277-282 */
    /* RunGenerativeModel (:317-326) */
    if (gen_n > 0 && lo_gen_num_samples_available(d->model) == 0) {
      float zeros[NF];
      memset(zeros, 0, sizeof(zeros));                                      /* ZeroFeatureEstimator::Estimate */
      if (lo_gen_add_features(d->model, zeros, NF) != 0) return -1;
    }
    if (lo_gen_generate_samples(d->model, gen_n, audio) < 0) return -1;
    /* RunComfortNoiseGenerator (:328-340) */
    if (cng_n > 0 && lo_gen_num_samples_available(d->cng) == 0) {
      float est[NMEL];
      d->n_noise_estimate++;
      if (d->use_fakes) memcpy(est, d->fake_noise, sizeof(est));
      else lo_noise_estimate(d->noise, est);
      if (lo_gen_add_features(d->cng, est, NMEL) != 0) return -1;
    }
    if (lo_gen_generate_samples(d->cng, cng_n, noise_hop) < 0) return -1;
    /* MaybeOverlapAndInsert (:342-373) */
    if (cng_n == 0) memcpy(out + produced, audio, sizeof(int16_t) * (size_t)gen_n);
    else if (gen_n == 0) memcpy(out + produced, noise_hop, sizeof(int16_t) * (size_t)cng_n);
    else {
      int fp = d->fade_progress;
      for (int i = 0; i < n; ++i) {
        /* float overlap_weight = (1.f + std::cos(fade_progress * M_PI / GetFadeDurationSamples())) / 2.f: the argument is
           double (M_PI), std::cos(double), (1.f + double) is double, / 2.f double, rounded to float on assignment */
        const float w = (float)((1.0 + cos((double)fp * M_PI / (double)FADE_SAMPLES)) / 2.0);
        /* int16 * float + int16 * (1.f - float): float arithmetic, push_back converts to int16 (truncation) */
        const float v = (float)audio[i] * w + (float)noise_hop[i] * (1.0f - w);
        out[produced + i] = (int16_t)v;
        fp += d->fade_direction;
      }
    }
    d->fade_progress = next_fade;                                           /* :303 */
    if (is_packet_received) {                                               /* :305-312 */
      d->n_noise_receive++;
      if (!d->use_fakes) {
        /* NoiseEstimator::ReceiveSamples buffers partial hops (noise_estimator.cc:144-160); whole hops are what the
           received-packet path produces when requests are hop aligned; partial ones are accumulated here */
        if (lo_noise_receive_partial(d->noise, audio, gen_n) != 0) return -1;
      }
    }
    produced += n;
  }
  return produced;
}

int lo_decoder_is_comfort_noise(const lo_decoder* d) { return d->fade_progress == FADE_SAMPLES; }
void lo_decoder_get_state(const lo_decoder* d, int* s3) { s3[0] = d->concealment_progress; s3[1] = d->fade_progress; s3[2] = d->fade_direction; }
void lo_decoder_set_state(lo_decoder* d, const int* s3) { d->concealment_progress = s3[0]; d->fade_progress = s3[1]; d->fade_direction = s3[2]; }
/* counters: {vq decode, model AddFeatures, model GenerateSamples, last model request, cng AddFeatures, cng GenerateSamples,
 *            last cng request, noise ReceiveSamples, noise noise_estimate} */
void lo_decoder_counters(const lo_decoder* d, int* c9) {
  c9[0] = d->n_vq_decode; c9[1] = d->model->calls_add; c9[2] = d->model->calls_generate; c9[3] = d->model->last_generate;
  c9[4] = d->cng->calls_add; c9[5] = d->cng->calls_generate; c9[6] = d->cng->last_generate;
  c9[7] = d->n_noise_receive; c9[8] = d->n_noise_estimate;
}
lo_noise* lo_decoder_noise(lo_decoder* d) { return d->noise; }
lo_cng* lo_decoder_cng(lo_decoder* d) { return d->cng_real; }

/* ------------------------------------------------------------ encoder with DTX (lyra_encoder.cc:113-156) ---- */

struct lo_encoder {
  lo_net* net;
  lo_rvq* rvq;
  lo_noise* noise;     /* NULL unless enable_dtx */
};

lo_encoder* lo_encoder_create(const char* model_dir, int enable_dtx) {
  lo_encoder* e = (lo_encoder*)calloc(1, sizeof(*e));
  char path[1024];
  snprintf(path, sizeof(path), "%s/soundstream_encoder.tflite", model_dir);
  e->net = lo_net_create(path);
  snprintf(path, sizeof(path), "%s/quantizer.tflite", model_dir);
  e->rvq = lo_rvq_create(path);
  if (enable_dtx) e->noise = lo_noise_create(16000, HOP, 640, NMEL);        /* lyra_encoder.cc:80-89 */
  if (!e->net || !e->rvq || (enable_dtx && !e->noise)) { lo_encoder_free(e); return NULL; }
  return e;
}

void lo_encoder_free(lo_encoder* e) {
  if (!e) return;
  lo_net_free(e->net); lo_rvq_free(e->rvq); lo_noise_free(e->noise);
  free(e);
}

/* returns the packet size in bytes (0 = the empty DTX packet), < 0 on error */
int lo_encoder_encode(lo_encoder* e, const int16_t* pcm, int n, int num_bits, uint8_t* packet) {
  if (n != HOP) return -1;                                                   /* :124-129 */
  if (e->noise) {
    if (lo_noise_receive_samples(e->noise, pcm, NULL) != 0) return -1;      /* :131-135 */
    if (lo_noise_is_noise(e->noise)) return 0;                              /* :137-140: Packet<0> packs to zero bytes */
  }
  float in[HOP], features[NF];
  for (int i = 0; i < HOP; ++i) in[i] = lo_int16_to_unit(pcm[i]);
  if (lo_net_invoke(e->net, in, HOP, features, NF) != 0) return -1;
  char bits[185];
  if (lo_rvq_quantize_bits(e->rvq, features, num_bits, bits) != 0) return -1;
  return lo_packet_pack(bits, 0, num_bits, packet);
}
