/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C) of the per-20 ms-frame hot path of google/lyra v1.3.2:
 *   SoundStreamEncoder::Extract           lyra/soundstream_encoder.cc:53-64
 *   ResidualVectorQuantizer::Quantize     lyra/residual_vector_quantizer.cc:77-110
 *   ResidualVectorQuantizer::DecodeToLossyFeatures   lyra/residual_vector_quantizer.cc:112-168
 *   Packet<184>::PackQuantized/UnpackPacket           lyra/packet.h:56-71,91-146
 *   LyraGanModel::RunConditioning/RunModel            lyra/lyra_gan_model.cc:53-64
 *   LogMelSpectrogramExtractorImpl::Extract           lyra/log_mel_spectrogram_extractor_impl.cc:96-126
 *   Int16ToUnitScalar / UnitToInt16Scalar             lyra/dsp_utils.h:53-60,79-88,104-108
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may
 * load this library, and only as the checker / reported CPU baseline — never as the product.
 * The product (lyra_b200/) has its own loader and CUDA kernels and fails loudly without a GPU.
 *
 * PARITY PINNING.  The reference binary cannot be built or imported offline (no bazel, TFLite,
 * XNNPACK, abseil, audio_dsp), so this oracle is pinned by the reference's own
 * fixtures instead (tests/test_oracle_golden.py):
 *   - log-mel known-answer vectors      lyra/log_mel_spectrogram_extractor_impl_test.cc:37-59  (tight)
 *   - packet byte layouts               lyra/packet_test.cc:93-300                            (exact)
 *   - RVQ round-trip distance < 1.11    lyra/residual_vector_quantizer_test.cc:41-54,104-111  (loose)
 *   - end-to-end LSD < 2.0 per hop      lyra/lyra_integration_test.cc:132-142                 (loose)
 * The SoundStream / LyraGAN numerics at the TFLite boundary are "parity unpinned" beyond the
 * integration bound: the reference's own unit tests check only shapes.
 */
#ifndef LYRA_ORACLE_H_
#define LYRA_ORACLE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- streaming conv nets (soundstream_encoder.tflite / lyragan.tflite), one stream per object ---- */
typedef struct lo_net lo_net;
lo_net* lo_net_create(const char* tflite_path);
void lo_net_free(lo_net* n);
int lo_net_reset(lo_net* n);                           /* all state variables back to zero */
int lo_net_invoke(lo_net* n, const float* in, int n_in, float* out, int n_out);
/* introspection for layer-by-layer GPU debugging */
int lo_net_num_tensors(const lo_net* n);
int lo_net_tensor_info(const lo_net* n, int idx, int* type, int* count, float* scale, int* zero_point);
int lo_net_read_tensor(const lo_net* n, int idx, void* dst, int max_bytes);
int lo_net_num_vars(const lo_net* n);
const char* lo_net_var_name(const lo_net* n, int v);
int lo_net_var_count(const lo_net* n, int v);
int lo_net_read_var(const lo_net* n, int v, float* dst, int max_count);

/* fixed-point helpers shared with the tests */
void lo_quantize_multiplier(double d, int32_t* qm, int* shift);
int32_t lo_mbqm(int32_t x, int32_t qm, int shift);

/* ---- residual vector quantizer (quantizer.tflite) ---- */
typedef struct lo_rvq lo_rvq;
lo_rvq* lo_rvq_create(const char* tflite_path);
void lo_rvq_free(lo_rvq* q);
int lo_rvq_num_stages(const lo_rvq* q);                /* 46 */
int lo_rvq_bits_per_stage(const lo_rvq* q);            /* 4  */
const float* lo_rvq_codebook(const lo_rvq* q, int stage); /* [16][64] */
/* indices[0..num_stages): stage index or -1 for stages >= num_quantizers (as the graph's output_0) */
int lo_rvq_encode(const lo_rvq* q, const float* features64, int num_quantizers, int32_t* indices);
/* indices may contain -1 (masked to a zero contribution) */
int lo_rvq_decode(const lo_rvq* q, const int32_t* indices, float* features64);
/* C++ glue of Quantize()/DecodeToLossyFeatures(): '0'/'1' strings, MSB-first, stage 0 first.
 * Return 0, or -1 for num_bits > 184 / not a multiple of bits-per-stage (reference returns nullopt). */
int lo_rvq_quantize_bits(const lo_rvq* q, const float* features64, int num_bits, char* bits_out);
int lo_rvq_decode_bits(const lo_rvq* q, const char* bits, int num_bits, float* features64);

/* ---- Packet<184> with 0 header bits (lyra/packet.h) ---- */
int lo_packet_size(int num_header_bits, int num_quantized_bits);
int lo_packet_pack(const char* bits, int num_header_bits, int num_quantized_bits, uint8_t* bytes);
int lo_packet_unpack(const uint8_t* bytes, int nbytes, int num_header_bits, int num_quantized_bits, char* bits_out);

/* ---- dsp_utils.h scalar conversions ---- */
float lo_int16_to_unit(int16_t v);
int16_t lo_unit_to_int16(float v);
float lo_log_spectral_distance(const float* a, const float* b, int n);

/* ---- log-mel extractor ---- */
typedef struct lo_logmel lo_logmel;
lo_logmel* lo_logmel_create(int sample_rate_hz, int hop, int window, int num_mel_bins);
void lo_logmel_free(lo_logmel* m);
int lo_logmel_extract(lo_logmel* m, const int16_t* audio, int n, float* out);

/* NoiseEstimator (lyra/noise_estimator.{h,cc}); see noise_estimator.c */
typedef struct lo_noise lo_noise;
lo_noise* lo_noise_create(int sample_rate_hz, int hop, int window, int num_features);
void lo_noise_free(lo_noise* e);
void lo_noise_set_constants(lo_noise* e, int hops_per_update, float max_smoothing, float bound_decay);   /* test peer ctor */
int lo_noise_receive_samples(lo_noise* e, const int16_t* hop, float* logmel_out /* may be NULL */);
void lo_noise_update(lo_noise* e, const float* current_power_db);          /* UpdateNoiseEstimate */
int lo_noise_compute_is_noise(const lo_noise* e, const float* current_power_db);
int lo_noise_is_noise(const lo_noise* e);
void lo_noise_estimate(const lo_noise* e, float* out);
void lo_noise_bound(const lo_noise* e, float* out);

int lo_noise_receive_partial(lo_noise* e, const int16_t* samples, int n);   /* ReceiveSamples with partial-hop buffering */

/* ---- GenerativeModel base: FIFO of feature vectors + partial-hop bookkeeping (lyra/generative_model_interface.h:45-134);
 *      run_conditioning(self, features, hop_out) produces one whole hop, generate slices it.  comfort_noise.c ---- */
typedef int (*lo_gen_conditioning_fn)(void* self, const float* features, int16_t* hop_out);
typedef struct lo_gen {
  int hop, nf, next_sample_in_hop;
  int head, count, cap;
  float* queue;
  int16_t* hop_samples;
  lo_gen_conditioning_fn run_conditioning;
  void* self;
  int calls_add, calls_generate, last_generate;      /* counters for the mock-style tests */
} lo_gen;
void lo_gen_init(lo_gen* g, int num_samples_per_hop, int num_features, lo_gen_conditioning_fn fn, void* self);
void lo_gen_free(lo_gen* g);
int lo_gen_add_features(lo_gen* g, const float* features, int n);           /* 0, or -1 for a wrong feature count */
int lo_gen_num_samples_available(const lo_gen* g);
int lo_gen_generate_samples(lo_gen* g, int num_samples, int16_t* out);      /* count, or -1 (reference: nullopt) */

/* ---- ComfortNoiseGenerator (lyra/comfort_noise_generator.{h,cc}); see comfort_noise.c for the restated audio_dsp pieces,
 *      the seeded phase policy and what pins it ---- */
typedef struct lo_cng lo_cng;
lo_cng* lo_cng_create(int sample_rate_hz, int hop, int window, int num_mel_bins, uint64_t seed);
void lo_cng_free(lo_cng* c);
lo_gen* lo_cng_gen(lo_cng* c);                                              /* AddFeatures / GenerateSamples / num_samples_available */
int lo_cng_condition(lo_cng* c, const float* log_mel_features, const uint32_t* phase /* [bins] or NULL */, int16_t* hop_out);
uint32_t lo_cng_phase_index(uint64_t seed, uint64_t hop, int bin);          /* 0..1023 */
double lo_cng_synthesis_gain(const lo_cng* c);
const double* lo_cng_norm(const lo_cng* c);
const double* lo_cng_synth(const lo_cng* c);

/* ---- LyraDecoder at 16 kHz: packet-loss concealment / comfort-noise / fade state machine (lyra/lyra_decoder.cc:172-383)
 *      and LyraEncoder::Encode with DTX (lyra/lyra_encoder.cc:113-156); lyra_decoder.c ---- */
typedef struct lo_decoder lo_decoder;
lo_decoder* lo_decoder_create(const char* model_dir, uint64_t cng_seed);                 /* real components */
lo_decoder* lo_decoder_create_fake(int16_t model_value, int16_t cng_value);               /* the reference tests' fakes */
void lo_decoder_free(lo_decoder* d);
int lo_decoder_set_encoded_packet(lo_decoder* d, const uint8_t* encoded, int nbytes);     /* 0 / -1 (reference: true / false) */
int lo_decoder_decode_samples(lo_decoder* d, int num_samples, int16_t* out);              /* count / -1 */
int lo_decoder_is_comfort_noise(const lo_decoder* d);
void lo_decoder_get_state(const lo_decoder* d, int* s3);   /* concealment_progress, fade_progress, fade_direction (-1 from / +1 to CNG) */
void lo_decoder_set_state(lo_decoder* d, const int* s3);   /* LyraDecoderPeer of lyra_decoder_test.cc:56-90 */
void lo_decoder_counters(const lo_decoder* d, int* c9);
lo_noise* lo_decoder_noise(lo_decoder* d);
lo_cng* lo_decoder_cng(lo_decoder* d);
typedef struct lo_encoder lo_encoder;
lo_encoder* lo_encoder_create(const char* model_dir, int enable_dtx);
void lo_encoder_free(lo_encoder* e);
int lo_encoder_encode(lo_encoder* e, const int16_t* pcm, int n, int num_bits, uint8_t* packet);   /* packet bytes (0 = DTX), < 0 error */

/* ---- Resampler / BufferedResampler (lyra/resampler.{h,cc}, lyra/buffered_resampler.{h,cc}); resampler.c ---- */
typedef struct lo_resampler lo_resampler;
lo_resampler* lo_resampler_create(int input_rate_hz, int output_rate_hz);                 /* rates in {8, 16, 32, 48} kHz */
void lo_resampler_free(lo_resampler* r);
void lo_resampler_reset(lo_resampler* r);
int lo_resampler_input_rate(const lo_resampler* r);
int lo_resampler_output_rate(const lo_resampler* r);
int lo_resampler_samples_until_steady_state(const lo_resampler* r);
int lo_resampler_resample(lo_resampler* r, const int16_t* in, int n, int16_t* out, int capacity);   /* number of outputs, -1 overflow */
int lo_resampler_design(int input_rate, int output_rate, int* num, int* den, float* coeffs, int capacity);   /* taps per phase */
void lo_resampler_coeffs(const lo_resampler* r, float* out);
double lo_bessel_i0(double x);
typedef struct lo_buffered_resampler lo_buffered_resampler;
lo_buffered_resampler* lo_buffered_resampler_create(int internal_rate_hz, int external_rate_hz);
void lo_buffered_resampler_free(lo_buffered_resampler* b);
int lo_buffered_resampler_leftover(const lo_buffered_resampler* b);
int lo_buffered_resampler_internal_samples(const lo_buffered_resampler* b, int num_external_requested);
int lo_buffered_resampler_filter_and_buffer(lo_buffered_resampler* b, int (*generator)(void*, int, int16_t*), void* user,
                                            int num_external_requested, int16_t* out);

/* ---- whole codec, one stream (LyraEncoder::Encode / LyraDecoder::{SetEncodedPacket,DecodeSamples}
 *      restricted to 16 kHz, no DTX, packets always received or concealed with zero features) ---- */
typedef struct lo_codec lo_codec;
lo_codec* lo_codec_create(const char* model_dir);
void lo_codec_free(lo_codec* c);
int lo_codec_reset(lo_codec* c);
/* pcm[320] -> packet (8/15/23 bytes for num_bits 64/120/184); also returns features and indices if non-NULL */
int lo_codec_encode(lo_codec* c, const int16_t* pcm, int num_bits, uint8_t* packet, float* features64, int32_t* indices46);
/* packet==NULL: packet lost -> LyraGAN is fed 64 zero features (lyra/lyra_decoder.cc:317-326) */
int lo_codec_decode(lo_codec* c, const uint8_t* packet, int num_bits, int16_t* pcm, float* lossy_features64, float* unit_out320);
lo_net* lo_codec_encoder_net(lo_codec* c);
lo_net* lo_codec_decoder_net(lo_codec* c);

/* ---- CPU baseline: `streams` independent streams x `frames` hops, one stream per thread at a time
 *      (mirrors TFLite num_threads = 1, lyra/tflite_model_wrapper.cc:51,68).
 *      stage_us[4] = mean microseconds per frame for {extract, quantize, dequantize, model_decode}
 *      (the split of lyra/lyra_benchmark_lib.cc:85-160).  Returns wall seconds, <0 on error. ---- */
double lo_cpu_bench(const char* model_dir, int streams, int frames, int num_bits, int threads,
                    uint32_t seed, double* stage_us, uint64_t* checksum);

#ifdef __cplusplus
}
#endif
#endif
