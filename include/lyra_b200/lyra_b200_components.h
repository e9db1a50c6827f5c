// C++ host side above the C ABI: drop-in counterparts of the reference's plugin classes for the hot path,
// with the reference's names, argument meaning and error behaviour (std::nullopt / false / nullptr, never
// an exception, never an abort).  Header-only; link liblyra_b200.so.
//
//   reference class (lyra/…)                                    here
//   SoundStreamEncoder          soundstream_encoder.{h,cc}       lyra_b200::SoundStreamEncoderB200
//   ResidualVectorQuantizer     residual_vector_quantizer.{h,cc} lyra_b200::ResidualVectorQuantizerB200
//   LyraGanModel                lyra_gan_model.{h,cc}            lyra_b200::LyraGanModelB200 (: GenerativeModel)
//   LogMelSpectrogramExtractorImpl  log_mel_spectrogram_extractor_impl.{h,cc}   lyra_b200::LogMelSpectrogramExtractorB200
//   NoiseEstimator              noise_estimator.{h,cc}           lyra_b200::NoiseEstimatorB200 (: NoiseEstimatorInterface)
//   Packet<184>                 packet.h                         lyra_b200::Packet184
//   ComfortNoiseGenerator       comfort_noise_generator.{h,cc}   lyra_b200::ComfortNoiseGeneratorB200 (: GenerativeModel)
//   Resampler / BufferedResampler  resampler.{h,cc} / buffered_resampler.{h,cc}   lyra_b200::ResamplerB200 / BufferedResamplerB200
//   LyraEncoder / LyraDecoder   lyra_encoder.{h,cc} / lyra_decoder.{h,cc}       lyra_b200::LyraEncoderB200 / LyraDecoderB200
//                                                                (8 / 16 / 32 / 48 kHz; DTX; packet-loss concealment, comfort
//                                                                 noise and the cross-fades between them, arbitrary request sizes)
//
// The reference's interface headers need abseil, which is not available in this build environment, so the
// three interfaces are restated below with std:: types (absl::Span<const T> -> pointer + size overloads on
// std::vector).  INTEGRATION.md shows the 1:1 binding a maintainer adds inside the reference tree, where the
// adapters derive from chromemedia::codec::{FeatureExtractorInterface, VectorQuantizerInterface, GenerativeModel}.
//
// Every object owns ONE stream id of a process-wide context (one context per device, created on first use with
// LYRA_B200_MAX_STREAMS, default 4096, stream slots).  Per-object calls go through the session's Coalescer: calls of the
// same kind that arrive from different threads while a launch is in flight are gathered into ONE batched C-ABI call
// (a thread-per-stream server gets n > 1 launches without changing its code); a lone caller runs at once with n = 1.
// Throughput users with their own batching call the batched C ABI directly (include/lyra_b200.h).
#pragma once

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <functional>
#include <cstdlib>
#include <memory>
#include <mutex>
#include <optional>
#include <queue>
#include <string>
#include <thread>
#include <vector>

#include "../lyra_b200.h"

namespace lyra_b200 {

// ---- the three plugin interfaces (lyra/feature_extractor_interface.h:32-39, lyra/vector_quantizer_interface.h:28-41,
//      lyra/generative_model_interface.h:32-42) ---------------------------------------------------------------------
class FeatureExtractorInterface {
 public:
  virtual ~FeatureExtractorInterface() {}
  virtual std::optional<std::vector<float>> Extract(const std::vector<int16_t>& audio) = 0;
};

class VectorQuantizerInterface {
 public:
  virtual ~VectorQuantizerInterface() {}
  virtual std::optional<std::string> Quantize(const std::vector<float>& features, int num_bits) const = 0;
  virtual std::optional<std::vector<float>> DecodeToLossyFeatures(const std::string& quantized_features) const = 0;
};

class GenerativeModelInterface {
 public:
  virtual ~GenerativeModelInterface() {}
  virtual bool AddFeatures(const std::vector<float>& features) = 0;
  virtual std::optional<std::vector<int16_t>> GenerateSamples(int num_samples) = 0;
  virtual int num_samples_available() const = 0;
};

// lyra/generative_model_interface.h:45-134: FIFO of feature vectors + partial-hop bookkeeping
class GenerativeModel : public GenerativeModelInterface {
 public:
  bool AddFeatures(const std::vector<float>& features) final {
    if ((int)features.size() != num_features_) return false;
    features_queue_.push(features);
    return true;
  }
  std::optional<std::vector<int16_t>> GenerateSamples(int num_samples) final {
    if (num_samples < 0) return std::nullopt;
    if (num_samples == 0) return std::vector<int16_t>(0);
    if (num_samples_available() == 0) return std::nullopt;
    if (next_sample_in_hop_ == 0 && !RunConditioning(features_queue_.front())) return std::nullopt;
    if (num_samples > num_samples_per_hop_ - next_sample_in_hop_) return std::nullopt;
    auto samples = RunModel(num_samples);
    if (samples.has_value()) {
      next_sample_in_hop_ += (int)samples->size();
      if (next_sample_in_hop_ == num_samples_per_hop_) { next_sample_in_hop_ = 0; features_queue_.pop(); }
    }
    return samples;
  }
  int num_samples_available() const final { return (int)features_queue_.size() * num_samples_per_hop_ - next_sample_in_hop_; }

 protected:
  GenerativeModel(int num_samples_per_hop, int num_features)
      : num_samples_per_hop_(num_samples_per_hop), num_features_(num_features), next_sample_in_hop_(0) {}
  virtual bool RunConditioning(const std::vector<float>& features) = 0;
  virtual std::optional<std::vector<int16_t>> RunModel(int num_samples) = 0;
  int next_sample_in_hop() const { return next_sample_in_hop_; }

 private:
  const int num_samples_per_hop_;
  const int num_features_;
  int next_sample_in_hop_;
  std::queue<std::vector<float>> features_queue_;
};


// ---- coalescing front for the per-object API (SURVEY.md section 7 "API impedance") ---------------------------------------
// The reference's objects are called once per stream and hop (e.g. lyra_encoder.cc:119-151, lyra_decoder.cc:228-315); the
// device kernels want many streams per launch.  Flat combining: a caller posts its request and, if no launch is being
// prepared, becomes the leader - it takes every queued request with the oldest one's key (kind, argument), gathers their
// inputs into contiguous rows, issues one batched C-ABI call and scatters the results; the other callers sleep until their
// request is marked done.  No helper thread, no added latency for a lone caller, results identical to n = 1 calls (the
// kernels treat streams independently; tests/cpp/test_components.cc compares threaded with serial runs bit for bit).
// LYRA_B200_COALESCE_US (default 0) lets a leader wait that long for company before it launches.
class Coalescer {
 public:
  enum Kind { kExtract, kQuantize, kDequantize, kGenerate, kCng, kLogMel, kNoiseUpdate };
  struct Call {
    int kind = 0, arg = 0, arg2 = 0;   // batch key: calls are merged only when all three agree
    int id = -1;                       // stream id; -1 for the stateless quantizer calls
    const void* in = nullptr;  size_t in_bytes = 0;     // one row in
    void* out = nullptr;       size_t out_bytes = 0;    // one row out
    uint8_t* flag = nullptr;                            // kNoiseUpdate: is_noise of the row
    int rc = -1;
    bool done = false;
  };
  struct Stats { uint64_t calls = 0, launches = 0, max_batch = 0; };

  Coalescer(lyra_b200_ctx* ctx, std::mutex* ctx_mutex, int max_batch) : ctx_(ctx), ctx_mu_(ctx_mutex), max_batch_(max_batch) {
    if (const char* e = std::getenv("LYRA_B200_COALESCE_US")) linger_us_ = std::atoi(e) > 0 ? std::atoi(e) : 0;
  }
  int Submit(Call& c) {
    std::unique_lock<std::mutex> lk(mu_);
    queue_.push_back(&c);
    while (!c.done) {
      if (leader_) { cv_.wait(lk); continue; }
      leader_ = true;
      if (const int us = linger_us_) { lk.unlock(); std::this_thread::sleep_for(std::chrono::microseconds(us)); lk.lock(); }
      std::vector<Call*> batch = Take();
      lk.unlock();
      Run(batch);
      lk.lock();
      for (Call* q : batch) q->done = true;
      ++stats_.launches;
      stats_.calls += batch.size();
      stats_.max_batch = std::max<uint64_t>(stats_.max_batch, batch.size());
      leader_ = false;
      cv_.notify_all();
    }
    return c.rc;
  }
  Stats stats() { std::lock_guard<std::mutex> lk(mu_); return stats_; }
  void set_linger_us(int us) { std::lock_guard<std::mutex> lk(mu_); linger_us_ = us > 0 ? us : 0; }

 private:
  // the oldest request decides the key; a stream id appears at most once per batch (a second call on the same stream
  // must see the first one's state update, so it waits for the next launch)
  std::vector<Call*> Take() {
    std::vector<Call*> batch, rest;
    const Call* head = queue_.front();
    if (seen_.size() < (size_t)max_batch_) seen_.assign((size_t)max_batch_, 0);      // stream ids are < max_streams = max_batch_
    for (Call* q : queue_) {
      bool take = q->kind == head->kind && q->arg == head->arg && q->arg2 == head->arg2 && (int)batch.size() < max_batch_;
      if (take && q->id >= 0 && q->id < max_batch_) {
        take = !seen_[(size_t)q->id];
        if (take) seen_[(size_t)q->id] = 1;
      }
      (take ? batch : rest).push_back(q);
    }
    for (const Call* b : batch) if (b->id >= 0 && b->id < max_batch_) seen_[(size_t)b->id] = 0;
    queue_.swap(rest);
    return batch;
  }
  void Run(const std::vector<Call*>& batch) {
    const int n = (int)batch.size();
    const Call& h = *batch[0];
    int rc;
    std::lock_guard<std::mutex> ctx_lock(*ctx_mu_);      // the context itself is single-caller (resampler, reset, getters use it too)
    if (n == 1) {
      rc = Issue(h, 1, &h.id, h.in, h.out, h.flag);
    } else {
      ids_.resize((size_t)n);
      in_.resize((size_t)n * h.in_bytes);
      out_.resize((size_t)n * h.out_bytes);
      flags_.assign((size_t)n, 0);
      for (int i = 0; i < n; ++i) {
        ids_[(size_t)i] = batch[(size_t)i]->id;
        std::copy_n((const uint8_t*)batch[(size_t)i]->in, h.in_bytes, in_.data() + (size_t)i * h.in_bytes);
      }
      rc = Issue(h, n, ids_.data(), in_.data(), out_.data(), flags_.data());
      if (rc == LYRA_B200_OK) for (int i = 0; i < n; ++i) {
        std::copy_n(out_.data() + (size_t)i * h.out_bytes, h.out_bytes, (uint8_t*)batch[(size_t)i]->out);
        if (batch[(size_t)i]->flag) *batch[(size_t)i]->flag = flags_[(size_t)i];
      }
    }
    for (Call* q : batch) q->rc = rc;
  }
  int Issue(const Call& h, int n, const int32_t* ids, const void* in, void* out, uint8_t* flags) {
    switch (h.kind) {
      case kExtract:     return lyra_b200_extract_features(ctx_, ids, n, (const int16_t*)in, (float*)out);
      case kQuantize:    return lyra_b200_quantize(ctx_, n, (const float*)in, h.arg, (uint8_t*)out, nullptr);
      case kDequantize:  return lyra_b200_dequantize(ctx_, n, (const uint8_t*)in, h.arg, (float*)out);
      case kGenerate:    return lyra_b200_generate(ctx_, ids, n, (const float*)in, (int16_t*)out);
      case kCng:         return lyra_b200_cng_generate(ctx_, ids, n, (const float*)in, (int16_t*)out);
      case kLogMel:      return lyra_b200_logmel(ctx_, h.arg, ids, n, (const int16_t*)in, h.arg2, (float*)out);
      case kNoiseUpdate: return lyra_b200_noise_update(ctx_, ids, n, (const int16_t*)in, nullptr, flags, nullptr);
    }
    return LYRA_B200_EINVAL;
  }

  lyra_b200_ctx* ctx_;
  std::mutex* ctx_mu_;
  int max_batch_, linger_us_ = 0;
  std::mutex mu_;
  std::condition_variable cv_;
  std::vector<Call*> queue_;
  bool leader_ = false;
  Stats stats_;
  std::vector<int32_t> ids_;           // leader-only scratch (one leader at a time)
  std::vector<uint8_t> in_, out_, flags_, seen_;
};

// ---- process-wide context + stream-id allocator ------------------------------------------------------------------
class Session {
 public:
  // nullptr when the context cannot be created (no GPU, bad model directory): factories then return nullptr,
  // like the reference's Create() functions (e.g. lyra/soundstream_encoder.cc:36-46).
  // One session per role: encoder-side objects (SoundStreamEncoder) share an encoder-only context, decoder-side objects
  // (LyraGanModel, NoiseEstimator) a decoder-only one, stateless ones (quantizer, log-mel) use the encoder's.  The two
  // contexts hold only their half of the streaming state and have their own mutex, so a LyraEncoder thread and a
  // LyraDecoder thread never wait for each other.
  static std::shared_ptr<Session> Get(const std::string& model_path, int role = LYRA_B200_ROLE_ENCODER, int device = 0) {
    static std::mutex mu;
    static std::weak_ptr<Session> cached[2];
    static std::string cached_path[2];
    static int cached_device[2] = {-1, -1};
    const int k = role == LYRA_B200_ROLE_DECODER ? 1 : 0;
    std::lock_guard<std::mutex> lock(mu);
    if (auto s = cached[k].lock()) if (cached_path[k] == model_path && cached_device[k] == device) return s;
    int max_streams = 4096;
    if (const char* e = std::getenv("LYRA_B200_MAX_STREAMS")) max_streams = std::atoi(e) > 0 ? std::atoi(e) : max_streams;
    lyra_b200_ctx* ctx = nullptr;
    if (lyra_b200_create_ex(model_path.c_str(), device, max_streams, k ? LYRA_B200_ROLE_DECODER : LYRA_B200_ROLE_ENCODER, &ctx) != LYRA_B200_OK)
      return nullptr;
    std::shared_ptr<Session> s(new Session(ctx, max_streams));
    cached[k] = s;
    cached_path[k] = model_path;
    cached_device[k] = device;
    return s;
  }
  ~Session() { lyra_b200_destroy(ctx_); }
  lyra_b200_ctx* ctx() const { return ctx_; }
  std::mutex& mutex() { return mu_; }   // calls on one context are serialised
  // one fixed-shape per-hop call of one object, merged with concurrent calls of the same kind (see Coalescer)
  int Call(int kind, int id, const void* in, size_t in_bytes, void* out, size_t out_bytes, int arg = 0, int arg2 = 0, uint8_t* flag = nullptr) {
    Coalescer::Call c;
    c.kind = kind; c.arg = arg; c.arg2 = arg2; c.id = id;
    c.in = in; c.in_bytes = in_bytes; c.out = out; c.out_bytes = out_bytes; c.flag = flag;
    return coalescer_.Submit(c);
  }
  Coalescer::Stats coalescer_stats() { return coalescer_.stats(); }
  void set_coalesce_linger_us(int us) { coalescer_.set_linger_us(us); }
  int Acquire() {
    std::lock_guard<std::mutex> lock(mu_);
    int id;
    if (!free_.empty()) { id = free_.back(); free_.pop_back(); }
    else if (next_ < max_streams_) id = next_++;
    else return -1;
    if (lyra_b200_reset(ctx_, &id, 1) != LYRA_B200_OK) { free_.push_back(id); return -1; }
    return id;
  }
  void Release(int id) { std::lock_guard<std::mutex> lock(mu_); free_.push_back(id); }

 private:
  Session(lyra_b200_ctx* c, int n) : ctx_(c), max_streams_(n), coalescer_(c, &mu_, n) {}
  lyra_b200_ctx* ctx_;
  int max_streams_, next_ = 0;
  std::vector<int> free_;
  std::mutex mu_;
  Coalescer coalescer_;
};

// ---- SoundStreamEncoder (lyra/soundstream_encoder.cc:36-64) --------------------------------------------------------
class SoundStreamEncoderB200 : public FeatureExtractorInterface {
 public:
  static std::unique_ptr<SoundStreamEncoderB200> Create(const std::string& model_path) {
    auto s = Session::Get(model_path);
    if (!s) return nullptr;
    const int id = s->Acquire();
    if (id < 0) return nullptr;
    return std::unique_ptr<SoundStreamEncoderB200>(new SoundStreamEncoderB200(std::move(s), id));
  }
  ~SoundStreamEncoderB200() override { session_->Release(id_); }
  std::optional<std::vector<float>> Extract(const std::vector<int16_t>& audio) override {
    if ((int)audio.size() != LYRA_B200_HOP) return std::nullopt;
    std::vector<float> out(LYRA_B200_NUM_FEATURES);
    if (session_->Call(Coalescer::kExtract, id_, audio.data(), audio.size() * sizeof(int16_t), out.data(), out.size() * sizeof(float)) != LYRA_B200_OK)
      return std::nullopt;
    return out;
  }

 private:
  SoundStreamEncoderB200(std::shared_ptr<Session> s, int id) : session_(std::move(s)), id_(id) {}
  std::shared_ptr<Session> session_;
  int id_;
};

// ---- Packet<184> with 0 header bits (lyra/packet.h:56-146, lyra/lyra_components.cc:57-60) --------------------------
struct Packet184 {
  static int PacketSize(int num_quantized_bits) { return (num_quantized_bits + 7) / 8; }
  static std::vector<uint8_t> PackQuantized(const std::string& bits) {
    std::vector<uint8_t> bytes((size_t)PacketSize((int)bits.size()), 0);
    for (size_t i = 0; i < bits.size(); ++i) if (bits[i] == '1') bytes[i >> 3] |= (uint8_t)(0x80u >> (i & 7));
    return bytes;
  }
  static std::optional<std::string> UnpackPacket(const std::vector<uint8_t>& packet, int num_quantized_bits) {
    if ((int)packet.size() != PacketSize(num_quantized_bits)) return std::nullopt;
    std::string bits((size_t)num_quantized_bits, '0');
    for (int i = 0; i < num_quantized_bits; ++i) if ((packet[(size_t)i >> 3] >> (7 - (i & 7))) & 1) bits[(size_t)i] = '1';
    return bits;
  }
};

// ---- ResidualVectorQuantizer (lyra/residual_vector_quantizer.cc:36-168) --------------------------------------------
class ResidualVectorQuantizerB200 : public VectorQuantizerInterface {
 public:
  // role: which process-wide context serves the (stateless) calls - a decoder's quantizer shares the decoder context, so a
  // decoder-only process never allocates encoder state and encoder / decoder threads never share a mutex
  static std::unique_ptr<ResidualVectorQuantizerB200> Create(const std::string& model_path, int role = LYRA_B200_ROLE_ENCODER) {
    auto s = Session::Get(model_path, role);
    if (!s) return nullptr;
    return std::unique_ptr<ResidualVectorQuantizerB200>(new ResidualVectorQuantizerB200(std::move(s)));
  }
  std::optional<std::string> Quantize(const std::vector<float>& features, int num_bits) const override {
    if ((int)features.size() != LYRA_B200_NUM_FEATURES) return std::nullopt;
    // invalid bit counts are refused by the library (EINVAL) before anything is written; the row size is the packet size
    if (num_bits <= 0 || num_bits > LYRA_B200_MAX_BITS) return std::nullopt;
    std::vector<uint8_t> packet((size_t)Packet184::PacketSize(num_bits));
    if (session_->Call(Coalescer::kQuantize, -1, features.data(), features.size() * sizeof(float), packet.data(), packet.size(), num_bits) != LYRA_B200_OK)
      return std::nullopt;
    return Packet184::UnpackPacket(packet, num_bits);
  }
  std::optional<std::vector<float>> DecodeToLossyFeatures(const std::string& quantized_features) const override {
    const int num_bits = (int)quantized_features.size();
    if (num_bits > LYRA_B200_MAX_BITS || num_bits % 4 != 0 || num_bits == 0) return std::nullopt;
    const std::vector<uint8_t> packet = Packet184::PackQuantized(quantized_features);
    std::vector<float> out(LYRA_B200_NUM_FEATURES);
    if (session_->Call(Coalescer::kDequantize, -1, packet.data(), packet.size(), out.data(), out.size() * sizeof(float), num_bits) != LYRA_B200_OK)
      return std::nullopt;
    return out;
  }

 private:
  explicit ResidualVectorQuantizerB200(std::shared_ptr<Session> s) : session_(std::move(s)) {}
  std::shared_ptr<Session> session_;
};

// ---- LyraGanModel (lyra/lyra_gan_model.cc:36-64) ------------------------------------------------------------------
class LyraGanModelB200 : public GenerativeModel {
 public:
  static std::unique_ptr<LyraGanModelB200> Create(const std::string& model_path, int num_features) {
    if (num_features != LYRA_B200_NUM_FEATURES) return nullptr;
    auto s = Session::Get(model_path, LYRA_B200_ROLE_DECODER);
    if (!s) return nullptr;
    const int id = s->Acquire();
    if (id < 0) return nullptr;
    return std::unique_ptr<LyraGanModelB200>(new LyraGanModelB200(std::move(s), id));
  }
  ~LyraGanModelB200() override { session_->Release(id_); }

 protected:
  bool RunConditioning(const std::vector<float>& features) override {
    // like the reference (lyra_gan_model.cc:53-58) the hop is generated in one go; RunModel slices it
    return session_->Call(Coalescer::kGenerate, id_, features.data(), features.size() * sizeof(float), hop_, sizeof(hop_)) == LYRA_B200_OK;
  }
  std::optional<std::vector<int16_t>> RunModel(int num_samples) override {
    return std::vector<int16_t>(hop_ + next_sample_in_hop(), hop_ + next_sample_in_hop() + num_samples);
  }

 private:
  LyraGanModelB200(std::shared_ptr<Session> s, int id) : GenerativeModel(LYRA_B200_HOP, LYRA_B200_NUM_FEATURES), session_(std::move(s)), id_(id) {}
  std::shared_ptr<Session> session_;
  int id_;
  int16_t hop_[LYRA_B200_HOP] = {0};
};

// ---- LogMelSpectrogramExtractorImpl at (16 kHz, hop 320, window 640) (lyra/log_mel_spectrogram_extractor_impl.cc:53-126)
class LogMelSpectrogramExtractorB200 : public FeatureExtractorInterface {
 public:
  // bank: which of the context's two independent extractor states this object advances (0 or 1)
  static std::unique_ptr<LogMelSpectrogramExtractorB200> Create(const std::string& model_path, int sample_rate_hz, int hop_length_samples,
                                                                int window_length_samples, int num_mel_bins, int bank = 0) {
    if (sample_rate_hz != 16000 || hop_length_samples != 320 || window_length_samples != 640) return nullptr;
    if (num_mel_bins != 160 && num_mel_bins != 64) return nullptr;
    auto s = Session::Get(model_path);
    if (!s) return nullptr;
    const int id = s->Acquire();
    if (id < 0) return nullptr;
    return std::unique_ptr<LogMelSpectrogramExtractorB200>(new LogMelSpectrogramExtractorB200(std::move(s), id, num_mel_bins, bank));
  }
  ~LogMelSpectrogramExtractorB200() override { session_->Release(id_); }
  std::optional<std::vector<float>> Extract(const std::vector<int16_t>& audio) override {
    if ((int)audio.size() != LYRA_B200_HOP) return std::nullopt;
    std::vector<float> out((size_t)num_mel_);
    if (session_->Call(Coalescer::kLogMel, id_, audio.data(), audio.size() * sizeof(int16_t), out.data(), out.size() * sizeof(float), bank_, num_mel_) != LYRA_B200_OK)
      return std::nullopt;
    return out;
  }

 private:
  LogMelSpectrogramExtractorB200(std::shared_ptr<Session> s, int id, int nmel, int bank)
      : session_(std::move(s)), id_(id), num_mel_(nmel), bank_(bank) {}
  std::shared_ptr<Session> session_;
  int id_, num_mel_, bank_;
};

// ---- lyra/noise_estimator_interface.h:28-38, implemented by lyra/noise_estimator.{h,cc} ---------------------------------
class NoiseEstimatorInterface {
 public:
  virtual ~NoiseEstimatorInterface() {}
  virtual bool ReceiveSamples(const std::vector<int16_t>& samples) = 0;
  virtual std::vector<float> noise_estimate() const = 0;
  virtual bool is_noise() const = 0;
};

class NoiseEstimatorB200 : public NoiseEstimatorInterface {
 public:
  // NoiseEstimator::Create(sample_rate_hz, num_samples_per_hop, num_samples_per_window, num_features), lyra/noise_estimator.cc:99-120
  static std::unique_ptr<NoiseEstimatorB200> Create(const std::string& model_path, int sample_rate_hz, int num_samples_per_hop,
                                                    int num_samples_per_window, int num_features) {
    if (sample_rate_hz != 16000 || num_samples_per_hop != 320 || num_samples_per_window != 640 || num_features != 160) return nullptr;
    auto s = Session::Get(model_path, LYRA_B200_ROLE_DECODER);
    if (!s) return nullptr;
    const int id = s->Acquire();
    if (id < 0) return nullptr;
    return std::unique_ptr<NoiseEstimatorB200>(new NoiseEstimatorB200(std::move(s), id));
  }
  ~NoiseEstimatorB200() override { session_->Release(id_); }
  // Samples are buffered until a hop is complete (their total may never straddle a hop boundary,
  // lyra/noise_estimator.cc:144-156); a full hop runs the log-mel + estimator kernels.
  bool ReceiveSamples(const std::vector<int16_t>& samples) override {
    if (samples.size() + hop_.size() > (size_t)LYRA_B200_HOP) return false;
    hop_.insert(hop_.end(), samples.begin(), samples.end());
    if ((int)hop_.size() < LYRA_B200_HOP) return true;
    uint8_t flag = 1;
    const int rc = session_->Call(Coalescer::kNoiseUpdate, id_, hop_.data(), hop_.size() * sizeof(int16_t), nullptr, 0, 0, 0, &flag);
    hop_.clear();
    if (rc != LYRA_B200_OK) return false;
    is_noise_ = flag != 0;
    return true;
  }
  std::vector<float> noise_estimate() const override {
    std::vector<float> out(160, 0.0f);
    std::lock_guard<std::mutex> lock(session_->mutex());
    lyra_b200_noise_estimate(session_->ctx(), &id_, 1, out.data(), nullptr);
    return out;
  }
  bool is_noise() const override { return is_noise_; }

 private:
  NoiseEstimatorB200(std::shared_ptr<Session> s, int id) : session_(std::move(s)), id_(id) {}
  std::shared_ptr<Session> session_;
  int id_;
  bool is_noise_ = true;
  std::vector<int16_t> hop_;
};

// ---- ComfortNoiseGenerator (lyra/comfort_noise_generator.{h,cc}) ---------------------------------------------------
class ComfortNoiseGeneratorB200 : public GenerativeModel {
 public:
  static std::unique_ptr<ComfortNoiseGeneratorB200> Create(const std::string& model_path, int sample_rate_hz, int num_samples_per_hop,
                                                           int window_length_samples, int num_mel_bins) {
    if (sample_rate_hz != 16000 || num_samples_per_hop != 320 || window_length_samples != 640 || num_mel_bins != 160) return nullptr;
    auto s = Session::Get(model_path, LYRA_B200_ROLE_DECODER);
    if (!s) return nullptr;
    const int id = s->Acquire();
    if (id < 0) return nullptr;
    return std::unique_ptr<ComfortNoiseGeneratorB200>(new ComfortNoiseGeneratorB200(std::move(s), id));
  }
  ~ComfortNoiseGeneratorB200() override { session_->Release(id_); }

 protected:
  bool RunConditioning(const std::vector<float>& features) override {       // FftFromFeatures + InvertFft, .cc:74-77
    return session_->Call(Coalescer::kCng, id_, features.data(), features.size() * sizeof(float), hop_, sizeof(hop_)) == LYRA_B200_OK;
  }
  std::optional<std::vector<int16_t>> RunModel(int num_samples) override {   // .cc:79-84
    return std::vector<int16_t>(hop_ + next_sample_in_hop(), hop_ + next_sample_in_hop() + num_samples);
  }

 private:
  ComfortNoiseGeneratorB200(std::shared_ptr<Session> s, int id) : GenerativeModel(LYRA_B200_HOP, 160), session_(std::move(s)), id_(id) {}
  std::shared_ptr<Session> session_;
  int id_;
  int16_t hop_[LYRA_B200_HOP] = {0};
};

// ---- Resampler (lyra/resampler_interface.h, lyra/resampler.{h,cc}) -------------------------------------------------------
inline bool IsSampleRateSupported(int hz) { return hz == 8000 || hz == 16000 || hz == 32000 || hz == 48000; }   // lyra_config.h:56,92-97
class ResamplerInterface {
 public:
  virtual ~ResamplerInterface() {}
  virtual std::vector<int16_t> Resample(const std::vector<int16_t>& audio) = 0;
  virtual void Reset() = 0;
  virtual int input_sample_rate_hz() const = 0;
  virtual int target_sample_rate_hz() const = 0;
  virtual int samples_until_steady_state() const = 0;
};

class ResamplerB200 : public ResamplerInterface {
 public:
  // one of the two rates must be the codec's 16 kHz (the only way the reference uses the class: lyra_encoder.cc:58-66,
  // lyra_decoder.cc:108-114); equal rates pass the samples through
  static std::unique_ptr<ResamplerB200> Create(const std::string& model_path, int input_sample_rate_hz, int target_sample_rate_hz) {
    if (!IsSampleRateSupported(input_sample_rate_hz) || !IsSampleRateSupported(target_sample_rate_hz)) return nullptr;
    if (input_sample_rate_hz != 16000 && target_sample_rate_hz != 16000) return nullptr;
    const bool to_internal = target_sample_rate_hz == 16000;
    auto s = Session::Get(model_path, to_internal ? LYRA_B200_ROLE_ENCODER : LYRA_B200_ROLE_DECODER);
    if (!s) return nullptr;
    const int id = s->Acquire();
    if (id < 0) return nullptr;
    return std::unique_ptr<ResamplerB200>(new ResamplerB200(std::move(s), id, input_sample_rate_hz, target_sample_rate_hz));
  }
  ~ResamplerB200() override { session_->Release(id_); }
  std::vector<int16_t> Resample(const std::vector<int16_t>& audio) override {
    if (in_ == out_ || audio.empty()) return audio;
    std::vector<int16_t> result;
    const bool to_internal = out_ == 16000;
    const int external = to_internal ? in_ : out_;
    std::lock_guard<std::mutex> lock(session_->mutex());
    const size_t chunk = out_ > in_ ? (size_t)(960 / (out_ / in_)) : 960;       // the device call takes / returns up to 960 samples per stream
    for (size_t pos = 0; pos < audio.size(); pos += chunk) {
      const int n = (int)std::min(chunk, audio.size() - pos);
      const int stride = (int)(((int64_t)n * out_ + in_ - 1) / in_) + 1;
      std::vector<int16_t> chunk((size_t)stride);
      int32_t count = 0;
      if (lyra_b200_resample(session_->ctx(), to_internal ? 1 : 0, &id_, 1, external, audio.data() + pos, n, chunk.data(), stride, &count) != LYRA_B200_OK)
        return std::vector<int16_t>();
      result.insert(result.end(), chunk.begin(), chunk.begin() + count);
    }
    return result;
  }
  void Reset() override { std::lock_guard<std::mutex> lock(session_->mutex()); lyra_b200_reset(session_->ctx(), &id_, 1); }
  int input_sample_rate_hz() const override { return in_; }
  int target_sample_rate_hz() const override { return out_; }
  int samples_until_steady_state() const override { return (int)(2.f * 17.f * ((float)out_ / (float)in_)); }   // resampler.cc:74-83

 private:
  ResamplerB200(std::shared_ptr<Session> s, int id, int in, int out) : session_(std::move(s)), id_(id), in_(in), out_(out) {}
  std::shared_ptr<Session> session_;
  int id_, in_, out_;
};

// ---- BufferedResampler (lyra/buffered_filter_interface.h, lyra/buffered_resampler.{h,cc}) --------------------------------
class BufferedResamplerB200 {
 public:
  static std::unique_ptr<BufferedResamplerB200> Create(const std::string& model_path, int internal_sample_rate, int external_sample_rate) {
    auto r = ResamplerB200::Create(model_path, internal_sample_rate, external_sample_rate);
    if (!r) return nullptr;
    return std::unique_ptr<BufferedResamplerB200>(new BufferedResamplerB200(std::move(r)));
  }
  explicit BufferedResamplerB200(std::unique_ptr<ResamplerInterface> resampler) : resampler_(std::move(resampler)) {}
  std::optional<std::vector<int16_t>> FilterAndBuffer(const std::function<std::optional<std::vector<int16_t>>(int)>& sample_generator,
                                                      int num_external_samples_requested) {                  // buffered_resampler.cc:63-91
    const int n_int = GetInternalNumSamplesToGenerate(num_external_samples_requested);
    std::vector<int16_t> samples((size_t)num_external_samples_requested);
    const int used = std::min((int)leftover_samples_.size(), num_external_samples_requested);                // :108-119
    std::copy(leftover_samples_.begin(), leftover_samples_.begin() + used, samples.begin());
    leftover_samples_.erase(leftover_samples_.begin(), leftover_samples_.begin() + used);
    auto internal = sample_generator(n_int);
    if (!internal.has_value() || (int)internal->size() != n_int) return std::nullopt;
    const std::vector<int16_t> external = resampler_->target_sample_rate_hz() == resampler_->input_sample_rate_hz()
                                              ? internal.value() : resampler_->Resample(internal.value());  // :121-129
    const int to_copy = num_external_samples_requested - used;                                               // :131-147
    if ((int)external.size() < to_copy) return std::nullopt;
    std::copy(external.begin(), external.begin() + to_copy, samples.begin() + used);
    leftover_samples_.insert(leftover_samples_.end(), external.begin() + to_copy, external.end());
    return samples;
  }
  int GetInternalNumSamplesToGenerate(int num_external_samples_requested) const {                            // :93-106
    if (num_external_samples_requested <= (int)leftover_samples_.size()) return 0;
    const int needed = num_external_samples_requested - (int)leftover_samples_.size();
    const float ratio = (float)resampler_->target_sample_rate_hz() / (float)resampler_->input_sample_rate_hz();
    return (int)std::ceil((float)needed / ratio);
  }

 private:
  std::vector<int16_t> leftover_samples_;
  std::unique_ptr<ResamplerInterface> resampler_;
};

// ---- factories with the reference's names (lyra/lyra_components.cc:42-60) ------------------------------------------
inline std::unique_ptr<VectorQuantizerInterface> CreateQuantizer(const std::string& model_path, int role = LYRA_B200_ROLE_ENCODER) {
  return ResidualVectorQuantizerB200::Create(model_path, role);
}
inline std::unique_ptr<GenerativeModelInterface> CreateGenerativeModel(int num_output_features, const std::string& model_path) {
  return LyraGanModelB200::Create(model_path, num_output_features);
}
inline std::unique_ptr<FeatureExtractorInterface> CreateFeatureExtractor(const std::string& model_path) { return SoundStreamEncoderB200::Create(model_path); }

// lyra/lyra_config.h:79-115
inline int GetPacketSize(int num_quantized_bits) { return (num_quantized_bits + 7) / 8; }
inline int BitrateToNumQuantizedBits(int bitrate) { return bitrate == 3200 ? 64 : bitrate == 6000 ? 120 : bitrate == 9200 ? 184 : -1; }
inline int PacketSizeToNumQuantizedBits(int packet_size) { return packet_size == 8 ? 64 : packet_size == 15 ? 120 : packet_size == 23 ? 184 : -1; }

// ---- LyraEncoder at 16 kHz, mono (lyra/lyra_encoder.h:44-122, lyra/lyra_encoder.cc:43-156), DTX included ------------------
class LyraEncoderB200 {
 public:
  static std::unique_ptr<LyraEncoderB200> Create(int sample_rate_hz, int num_channels, int bitrate, bool enable_dtx, const std::string& model_path) {
    if (!IsSampleRateSupported(sample_rate_hz) || num_channels != 1) return nullptr;   // AreParamsSupported, lyra_config.h:119-168
    const int bits = BitrateToNumQuantizedBits(bitrate);
    if (bits < 0) return nullptr;
    std::unique_ptr<ResamplerInterface> rs;
    if (sample_rate_hz != 16000) {                                       // lyra_encoder.cc:58-66
      rs = ResamplerB200::Create(model_path, sample_rate_hz, 16000);
      if (!rs) return nullptr;
    }
    auto fe = CreateFeatureExtractor(model_path);
    auto vq = CreateQuantizer(model_path);
    if (!fe || !vq) return nullptr;
    std::unique_ptr<NoiseEstimatorInterface> ne;
    if (enable_dtx) {                                                    // lyra_encoder.cc:80-89
      ne = NoiseEstimatorB200::Create(model_path, 16000, LYRA_B200_HOP, 640, 160);
      if (!ne) return nullptr;
    }
    return std::unique_ptr<LyraEncoderB200>(new LyraEncoderB200(std::move(rs), std::move(fe), std::move(vq), std::move(ne), sample_rate_hz, bits));
  }
  std::optional<std::vector<uint8_t>> Encode(const std::vector<int16_t>& audio_in) {
    if ((int)audio_in.size() != sample_rate_hz_ / 50) return std::nullopt;            // exactly one hop at the external rate
    const std::vector<int16_t> audio = resampler_ ? resampler_->Resample(audio_in) : audio_in;   // lyra_encoder.cc:118-122
    if ((int)audio.size() != LYRA_B200_HOP) return std::nullopt;                      // lyra_encoder.cc:124-129
    if (noise_estimator_) {                                                           // :131-141
      if (!noise_estimator_->ReceiveSamples(audio)) return std::nullopt;
      if (noise_estimator_->is_noise()) return std::vector<uint8_t>();               // the empty packet
    }
    auto features = feature_extractor_->Extract(audio);
    if (!features.has_value()) return std::nullopt;
    auto quantized = vector_quantizer_->Quantize(features.value(), num_quantized_bits_);
    if (!quantized.has_value()) return std::nullopt;
    return Packet184::PackQuantized(quantized.value());
  }
  bool set_bitrate(int bitrate) {
    const int bits = BitrateToNumQuantizedBits(bitrate);
    if (bits < 0) return false;
    num_quantized_bits_ = bits;
    return true;
  }
  int sample_rate_hz() const { return sample_rate_hz_; }
  int num_channels() const { return 1; }
  int bitrate() const { return GetPacketSize(num_quantized_bits_) * 8 * 50; }
  int frame_rate() const { return 50; }

 private:
  LyraEncoderB200(std::unique_ptr<ResamplerInterface> rs, std::unique_ptr<FeatureExtractorInterface> fe, std::unique_ptr<VectorQuantizerInterface> vq,
                  std::unique_ptr<NoiseEstimatorInterface> ne, int sample_rate_hz, int bits)
      : resampler_(std::move(rs)), feature_extractor_(std::move(fe)), vector_quantizer_(std::move(vq)), noise_estimator_(std::move(ne)),
        sample_rate_hz_(sample_rate_hz), num_quantized_bits_(bits) {}
  std::unique_ptr<ResamplerInterface> resampler_;
  std::unique_ptr<FeatureExtractorInterface> feature_extractor_;
  std::unique_ptr<VectorQuantizerInterface> vector_quantizer_;
  std::unique_ptr<NoiseEstimatorInterface> noise_estimator_;
  int sample_rate_hz_;
  int num_quantized_bits_;
};

// ---- LyraDecoder at 16 kHz (lyra/lyra_decoder.h:41-163, lyra/lyra_decoder.cc:95-383): packet-loss concealment, comfort noise and
//      the fades between them, for any number of requested samples.  The components are interfaces, as in the reference's
//      constructor (lyra_decoder.cc:148-170), so tests can plug fakes like lyra/lyra_decoder_test.cc does. -----------------------
class LyraDecoderB200 {
 public:
  enum FadeDirection { kFadeFromCNG = -1, kFadeToCNG = 1 };                          // lyra_decoder.h:105-108

  static std::unique_ptr<LyraDecoderB200> Create(int sample_rate_hz, int num_channels, const std::string& model_path) {
    if (!IsSampleRateSupported(sample_rate_hz) || num_channels != 1) return nullptr;
    auto rs = BufferedResamplerB200::Create(model_path, 16000, sample_rate_hz);        // lyra_decoder.cc:108-114
    if (!rs) return nullptr;
    auto gm = CreateGenerativeModel(LYRA_B200_NUM_FEATURES, model_path);
    auto cng = ComfortNoiseGeneratorB200::Create(model_path, 16000, LYRA_B200_HOP, 640, 160);
    auto ne = NoiseEstimatorB200::Create(model_path, 16000, LYRA_B200_HOP, 640, 160);
    auto vq = CreateQuantizer(model_path, LYRA_B200_ROLE_DECODER);
    if (!gm || !cng || !ne || !vq) return nullptr;
    return std::unique_ptr<LyraDecoderB200>(new LyraDecoderB200(std::move(gm), std::move(cng), std::move(vq), std::move(ne), std::move(rs), sample_rate_hz));
  }
  LyraDecoderB200(std::unique_ptr<GenerativeModelInterface> generative_model, std::unique_ptr<GenerativeModelInterface> comfort_noise_generator,
                  std::unique_ptr<VectorQuantizerInterface> vector_quantizer, std::unique_ptr<NoiseEstimatorInterface> noise_estimator,
                  std::unique_ptr<BufferedResamplerB200> resampler = nullptr, int external_sample_rate_hz = 16000)
      : generative_model_(std::move(generative_model)), comfort_noise_generator_(std::move(comfort_noise_generator)),
        vector_quantizer_(std::move(vector_quantizer)), noise_estimator_(std::move(noise_estimator)), resampler_(std::move(resampler)),
        external_sample_rate_hz_(external_sample_rate_hz) {}

  bool SetEncodedPacket(const std::vector<uint8_t>& encoded) {                       // lyra_decoder.cc:172-209
    const int bits = PacketSizeToNumQuantizedBits((int)encoded.size());
    if (bits < 0) return false;
    const auto unpacked = Packet184::UnpackPacket(encoded, bits);
    if (!unpacked.has_value()) return false;
    if (concealment_progress_ == kConcealmentSamples) concealment_progress_ = -comfort_noise_generator_->num_samples_available();
    else if (concealment_progress_ > 0) concealment_progress_ = -generative_model_->num_samples_available();
    auto features = vector_quantizer_->DecodeToLossyFeatures(unpacked.value());
    if (!features.has_value()) return false;
    return generative_model_->AddFeatures(features.value());                        // ZeroFeatureEstimator::Update is a no-op
  }

  std::optional<std::vector<int16_t>> DecodeSamples(int num_samples) {               // lyra_decoder.cc:211-226
    if (num_samples < 0) return std::nullopt;
    if (!resampler_) return DecodeSamplesInternal(num_samples);
    return resampler_->FilterAndBuffer([this](int n) { return DecodeSamplesInternal(n); }, num_samples);
  }

  std::optional<std::vector<int16_t>> DecodeSamplesInternal(int num_samples) {       // lyra_decoder.cc:228-315 (internal rate)
    if (num_samples < 0) return std::nullopt;
    std::vector<int16_t> result;
    result.reserve((size_t)num_samples);
    while ((int)result.size() < num_samples) {
      const int n = NumSamplesToGenerate(num_samples, (int)result.size());
      const bool is_packet_received = generative_model_->num_samples_available() > 0 && concealment_progress_ == 0;
      if (is_packet_received) fade_direction_ = kFadeFromCNG;
      else if (concealment_progress_ == kConcealmentSamples) fade_direction_ = kFadeToCNG;
      else concealment_progress_ += n;
      int cng_n = n, gen_n = n;
      int next_fade = fade_progress_ + fade_direction_ * n;
      if (fade_direction_ == kFadeToCNG && fade_progress_ == kFadeSamples) { next_fade = kFadeSamples; gen_n = 0; }
      else if (fade_direction_ == kFadeFromCNG && fade_progress_ == 0) { next_fade = 0; cng_n = 0; }
      if (gen_n > 0 && generative_model_->num_samples_available() == 0 &&                                      // RunGenerativeModel :317-326
          !generative_model_->AddFeatures(std::vector<float>(LYRA_B200_NUM_FEATURES, 0.0f))) return std::nullopt;
      auto audio = generative_model_->GenerateSamples(gen_n);
      if (!audio.has_value()) return std::nullopt;
      if (cng_n > 0 && comfort_noise_generator_->num_samples_available() == 0 &&                              // RunComfortNoiseGenerator :328-340
          !comfort_noise_generator_->AddFeatures(noise_estimator_->noise_estimate())) return std::nullopt;
      auto noise = comfort_noise_generator_->GenerateSamples(cng_n);
      if (!noise.has_value()) return std::nullopt;
      if (noise->empty()) result.insert(result.end(), audio->begin(), audio->end());                          // MaybeOverlapAndInsert :342-373
      else if (audio->empty()) result.insert(result.end(), noise->begin(), noise->end());
      else {
        if (audio->size() != noise->size()) return std::nullopt;
        int fp = fade_progress_;
        for (size_t i = 0; i < audio->size(); ++i) {
          const float w = (1.f + std::cos(fp * M_PI / kFadeSamples)) / 2.f;
          result.push_back((int16_t)((*audio)[i] * w + (*noise)[i] * (1.f - w)));
          fp += fade_direction_;
        }
      }
      fade_progress_ = next_fade;
      if (is_packet_received && !noise_estimator_->ReceiveSamples(audio.value())) return std::nullopt;
    }
    return result;
  }
  int sample_rate_hz() const { return external_sample_rate_hz_; }
  int num_channels() const { return 1; }
  int frame_rate() const { return 50; }
  bool is_comfort_noise() const { return fade_progress_ == kFadeSamples; }           // lyra_decoder.cc:381-383

  // test peer (lyra_decoder_test.cc:56-90)
  void SetStateForTest(int concealment_progress, int fade_progress, FadeDirection d) { concealment_progress_ = concealment_progress; fade_progress_ = fade_progress; fade_direction_ = d; }
  int concealment_progress() const { return concealment_progress_; }
  int fade_progress() const { return fade_progress_; }

 private:
  static constexpr int kConcealmentSamples = 1280, kFadeSamples = 640;               // 0.08 s, 0.04 s (lyra_decoder.cc:42-63)
  int NumSamplesToGenerate(int requested, int so_far) const {                        // lyra_decoder.cc:65-93
    int remaining;
    if (concealment_progress_ < 0) remaining = -concealment_progress_;
    else if (concealment_progress_ < kConcealmentSamples) remaining = generative_model_->num_samples_available() % LYRA_B200_HOP;
    else remaining = comfort_noise_generator_->num_samples_available();
    if (remaining == 0) remaining = LYRA_B200_HOP;
    return requested - so_far < remaining ? requested - so_far : remaining;
  }
  std::unique_ptr<GenerativeModelInterface> generative_model_, comfort_noise_generator_;
  std::unique_ptr<VectorQuantizerInterface> vector_quantizer_;
  std::unique_ptr<NoiseEstimatorInterface> noise_estimator_;
  std::unique_ptr<BufferedResamplerB200> resampler_;
  int external_sample_rate_hz_ = 16000;
  int concealment_progress_ = 0, fade_progress_ = 0;
  FadeDirection fade_direction_ = kFadeFromCNG;
};

}  // namespace lyra_b200
