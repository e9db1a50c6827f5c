// GPU tier: the C++ adapters of include/lyra_b200/lyra_b200_components.h (the reference's plugin surface) against the
// CPU oracle, in the style of the reference's own unit tests (soundstream_encoder_test.cc, residual_vector_quantizer_test.cc,
// lyra_gan_model_test.cc, lyra_integration_test.cc).  Links liblyra_b200.so (product) and liblyra_oracle.so (checker).
#include <cmath>
#include <cstdio>
#include <cstring>
#include <random>
#include <string>
#include <thread>
#include <vector>

#include "lyra_b200/lyra_b200_components.h"
#include "lyra_oracle.h"

static int g_fail = 0;
#define CHECK(cond) do { if (!(cond)) { std::printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #cond); ++g_fail; } } while (0)

int main(int argc, char** argv) {
  const std::string model = argc > 1 ? argv[1] : "lyra_b200/model_coeffs";
  using namespace lyra_b200;
  // Create() with a bad path returns nullptr (soundstream_encoder_test.cc:40-44, residual_vector_quantizer_test.cc:76-78)
  CHECK(SoundStreamEncoderB200::Create("invalid/model/path") == nullptr);
  CHECK(ResidualVectorQuantizerB200::Create("invalid/model/path") == nullptr);
  CHECK(LyraGanModelB200::Create("invalid/model/path", 64) == nullptr);

  auto enc = SoundStreamEncoderB200::Create(model);
  auto vq = ResidualVectorQuantizerB200::Create(model);
  auto gan = LyraGanModelB200::Create(model, 64);
  CHECK(enc && vq && gan);
  if (!enc || !vq || !gan) { std::printf("cannot create components: %s\n", lyra_b200_last_error(nullptr)); return 2; }
  lo_codec* ref = lo_codec_create(model.c_str());
  CHECK(ref != nullptr);

  std::mt19937 rng(5);
  std::uniform_int_distribution<int> d(-8192, 8191);
  // wrong sizes -> nullopt / false
  CHECK(!enc->Extract(std::vector<int16_t>(319)).has_value());
  CHECK(!vq->Quantize(std::vector<float>(64), 185).has_value());
  CHECK(!vq->Quantize(std::vector<float>(64), 62).has_value());
  CHECK(!vq->DecodeToLossyFeatures(std::string(185, '0')).has_value());
  CHECK(!vq->DecodeToLossyFeatures(std::string(62, '0')).has_value());
  CHECK(!gan->AddFeatures(std::vector<float>(63)));
  CHECK(!gan->GenerateSamples(10).has_value());          // nothing queued (lyra_gan_model_test.cc)

  for (int f = 0; f < 6; ++f) {
    std::vector<int16_t> pcm(320);
    for (auto& v : pcm) v = (int16_t)d(rng);
    auto feat = enc->Extract(pcm);
    CHECK(feat.has_value() && feat->size() == 64);
    uint8_t rpkt[24];
    float rfeat[64];
    int32_t ridx[46];
    const int bits = f % 3 == 0 ? 64 : f % 3 == 1 ? 120 : 184;
    CHECK(lo_codec_encode(ref, pcm.data(), bits, rpkt, rfeat, ridx) == GetPacketSize(bits));
    CHECK(std::memcmp(feat->data(), rfeat, sizeof(rfeat)) == 0);
    auto q = vq->Quantize(*feat, bits);
    CHECK(q.has_value() && (int)q->size() == bits);
    const std::vector<uint8_t> pkt = Packet184::PackQuantized(*q);
    CHECK((int)pkt.size() == GetPacketSize(bits) && std::memcmp(pkt.data(), rpkt, pkt.size()) == 0);
    auto lossy = vq->DecodeToLossyFeatures(*q);
    CHECK(lossy.has_value());
    int16_t rpcm[320];
    float rlossy[64];
    CHECK(lo_codec_decode(ref, rpkt, bits, rpcm, rlossy, nullptr) == 0);
    CHECK(std::memcmp(lossy->data(), rlossy, sizeof(rlossy)) == 0);
    // partial-hop generation: 100 + 220 samples of one hop (generative_model_interface.h:62-101)
    CHECK(gan->AddFeatures(*lossy));
    CHECK(gan->num_samples_available() == 320);
    auto a = gan->GenerateSamples(100);
    CHECK(a.has_value() && a->size() == 100);
    CHECK(!gan->GenerateSamples(221).has_value());        // more than the hop holds
    auto b = gan->GenerateSamples(220);
    CHECK(b.has_value() && b->size() == 220 && gan->num_samples_available() == 0);
    a->insert(a->end(), b->begin(), b->end());
    CHECK(std::memcmp(a->data(), rpcm, sizeof(rpcm)) == 0);
  }
  lo_codec_free(ref);

  // LyraEncoder / LyraDecoder counterparts: encode -> decode, one lost packet, vs a fresh oracle stream
  {
    auto e = LyraEncoderB200::Create(16000, 1, 6000, false, model);
    auto dcd = LyraDecoderB200::Create(16000, 1, model);
    CHECK(e && dcd);
    CHECK(LyraEncoderB200::Create(16000, 1, 5000, false, model) == nullptr);   // unsupported bitrate
    CHECK(LyraEncoderB200::Create(44100, 1, 6000, false, model) == nullptr);   // unsupported rate
    lo_codec* r = lo_codec_create(model.c_str());
    for (int f = 0; f < 5; ++f) {
      std::vector<int16_t> pcm(320);
      for (auto& v : pcm) v = (int16_t)d(rng);
      auto pkt = e->Encode(pcm);
      CHECK(pkt.has_value() && pkt->size() == 15);
      uint8_t rp[24];
      lo_codec_encode(r, pcm.data(), 120, rp, nullptr, nullptr);
      CHECK(std::memcmp(pkt->data(), rp, 15) == 0);
      const bool lost = f == 2;
      if (!lost) CHECK(dcd->SetEncodedPacket(*pkt));
      auto out = dcd->DecodeSamples(320);
      int16_t rpcm[320];
      lo_codec_decode(r, lost ? nullptr : rp, 120, rpcm, nullptr, nullptr);
      CHECK(out.has_value() && out->size() == 320 && std::memcmp(out->data(), rpcm, sizeof(rpcm)) == 0);
    }
    CHECK(!dcd->SetEncodedPacket(std::vector<uint8_t>(9)));                     // incomplete packet
    CHECK(e->set_bitrate(9200) && e->bitrate() == 9200 && !e->set_bitrate(1234));
    lo_codec_free(r);
  }
  // NoiseEstimator counterpart: samples in two pieces per hop, against the oracle's estimator
  {
    CHECK(NoiseEstimatorB200::Create(model, 16000, 320, 640, 64) == nullptr);   // only the decoder's configuration
    auto ne = NoiseEstimatorB200::Create(model, 16000, 320, 640, 160);
    CHECK(ne != nullptr);
    lo_noise* r = lo_noise_create(16000, 320, 640, 160);
    CHECK(ne->is_noise());
    for (int f = 0; f < 12; ++f) {
      std::vector<int16_t> pcm(320);
      for (auto& v : pcm) v = f % 4 == 3 ? 0 : (int16_t)(d(rng) >> (f % 3 == 0 ? 0 : 6));
      CHECK(ne->ReceiveSamples(std::vector<int16_t>(pcm.begin(), pcm.begin() + 100)));
      CHECK(!ne->ReceiveSamples(std::vector<int16_t>(221)));                      // would straddle the hop boundary
      CHECK(ne->ReceiveSamples(std::vector<int16_t>(pcm.begin() + 100, pcm.end())));
      CHECK(lo_noise_receive_samples(r, pcm.data(), nullptr) == 0);
      float want[160];
      lo_noise_estimate(r, want);
      const std::vector<float> got = ne->noise_estimate();
      CHECK(ne->is_noise() == (lo_noise_is_noise(r) != 0));
      CHECK(got.size() == 160 && std::memcmp(got.data(), want, sizeof(want)) == 0);
    }
    lo_noise_free(r);
  }
  // LyraDecoder's concealment / comfort-noise / fade state machine with the reference tests' fakes (constant-valued generative
  // models, fixed lossy features and noise estimate; lyra_decoder_test.cc:143-151, testing/mock_generative_model.h:33-53) against
  // the oracle's restatement driven by the same random script of packets and arbitrary request sizes
  {
    struct FakeModel : GenerativeModel {
      FakeModel(int16_t v, int nf) : GenerativeModel(320, nf), v_(v) {}
      bool RunConditioning(const std::vector<float>&) override { return true; }
      std::optional<std::vector<int16_t>> RunModel(int n) override { return std::vector<int16_t>((size_t)n, v_); }
      int16_t v_;
    };
    struct FakeVq : VectorQuantizerInterface {
      std::optional<std::string> Quantize(const std::vector<float>&, int) const override { return std::nullopt; }
      std::optional<std::vector<float>> DecodeToLossyFeatures(const std::string&) const override {
        std::vector<float> f(64);
        for (int i = 0; i < 64; ++i) f[(size_t)i] = (float)i;
        return f;
      }
    };
    struct FakeNoise : NoiseEstimatorInterface {
      bool ReceiveSamples(const std::vector<int16_t>&) override { return true; }
      std::vector<float> noise_estimate() const override { std::vector<float> f(160); for (int i = 0; i < 160; ++i) f[(size_t)i] = 10.f + (float)i; return f; }
      bool is_noise() const override { return false; }
    };
    LyraDecoderB200 dec(std::make_unique<FakeModel>((int16_t)-10000, 64), std::make_unique<FakeModel>((int16_t)10000, 160),
                        std::make_unique<FakeVq>(), std::make_unique<FakeNoise>());
    lo_decoder* r = lo_decoder_create_fake(-10000, 10000);
    std::uniform_int_distribution<int> coin(0, 99), len(0, 700);
    const std::vector<uint8_t> zeros(8, 0);
    int faded = 0;
    for (int step = 0; step < 400; ++step) {
      const int burst = (step / 40) % 2;                       // alternate good and bad stretches so every state is visited
      if (coin(rng) < (burst ? 10 : 85)) { CHECK(dec.SetEncodedPacket(zeros)); CHECK(lo_decoder_set_encoded_packet(r, zeros.data(), 8) == 0); }
      const int n = coin(rng) < 50 ? 320 : len(rng);
      auto out = dec.DecodeSamples(n);
      std::vector<int16_t> want((size_t)n + 1);
      CHECK(lo_decoder_decode_samples(r, n, want.data()) == n);
      CHECK(out.has_value() && (int)out->size() == n && std::memcmp(out->data(), want.data(), sizeof(int16_t) * (size_t)n) == 0);
      int s3[3];
      lo_decoder_get_state(r, s3);
      CHECK(dec.concealment_progress() == s3[0] && dec.fade_progress() == s3[1]);
      CHECK(dec.is_comfort_noise() == (lo_decoder_is_comfort_noise(r) != 0));
      for (int16_t v : *out) faded += v != -10000 && v != 10000;
    }
    CHECK(faded > 0);                                          // cross-faded samples were produced and matched
    CHECK(!dec.DecodeSamples(-1).has_value());
    lo_decoder_free(r);
  }
  // the real components behind LyraDecoder::Create: comfort noise generator basics (comfort_noise_generator_test.cc:42-98) and a
  // long outage that ends in comfort noise, then recovery
  {
    CHECK(ComfortNoiseGeneratorB200::Create(model, 16000, 320, 640, 64) == nullptr);
    auto cng = ComfortNoiseGeneratorB200::Create(model, 16000, 320, 640, 160);
    CHECK(cng != nullptr);
    CHECK(!cng->AddFeatures(std::vector<float>(159, 1.0f)) && !cng->GenerateSamples(320).has_value());
    CHECK(cng->AddFeatures(std::vector<float>(160, 0.0f)));
    CHECK(!cng->GenerateSamples(321).has_value() && !cng->GenerateSamples(-1).has_value());
    auto z = cng->GenerateSamples(320);
    CHECK(z.has_value() && z->size() == 320);
    if (z.has_value()) for (int16_t v : *z) CHECK(v == 0);     // features without energy give silence
    auto e = LyraEncoderB200::Create(16000, 1, 3200, false, model);
    auto dcd = LyraDecoderB200::Create(16000, 1, model);
    CHECK(e && dcd);
    bool reached_cn = false, back = false;
    for (int f = 0; f < 30 && e && dcd; ++f) {
      std::vector<int16_t> pcm(320);
      for (auto& v : pcm) v = (int16_t)d(rng);
      auto pkt = e->Encode(pcm);
      CHECK(pkt.has_value());
      if (f < 5 || f >= 18) CHECK(dcd->SetEncodedPacket(*pkt));
      auto out = dcd->DecodeSamples(f % 2 ? 320 : 123);        // odd request sizes keep working
      CHECK(out.has_value());
      if (f % 2 == 0) CHECK(dcd->DecodeSamples(197).has_value());
      reached_cn |= dcd->is_comfort_noise();
      back |= reached_cn && !dcd->is_comfort_noise() && dcd->fade_progress() == 0 && dcd->concealment_progress() == 0;
    }
    CHECK(reached_cn && back);
  }
  // DTX (lyra_encoder.cc:131-141): digital silence becomes empty packets from the second hop on, speech-level input is encoded
  {
    auto e = LyraEncoderB200::Create(16000, 1, 3200, true, model);
    CHECK(e != nullptr);
    for (int f = 0; f < 6 && e; ++f) {
      auto pkt = e->Encode(std::vector<int16_t>(320, 0));
      CHECK(pkt.has_value() && pkt->size() == (f == 0 ? 8u : 0u));
    }
    std::vector<int16_t> loud(320);
    for (auto& v : loud) v = (int16_t)d(rng);
    auto pkt = e ? e->Encode(loud) : std::nullopt;
    CHECK(pkt.has_value() && pkt->size() == 8);
  }
  // Resampler / BufferedResampler / the other sample rates of LyraEncoder and LyraDecoder (resampler_test.cc, buffered_resampler_test.cc,
  // lyra_encoder_test.cc / lyra_decoder_test.cc size checks)
  {
    CHECK(ResamplerB200::Create(model, 16000, 44100) == nullptr && ResamplerB200::Create(model, 8000, 48000) == nullptr);
    for (int rate : {8000, 32000, 48000}) {
      for (int dir = 0; dir < 2; ++dir) {
        const int a = dir ? 16000 : rate, b = dir ? rate : 16000;
        auto rs = ResamplerB200::Create(model, a, b);
        lo_resampler* r = lo_resampler_create(a, b);
        CHECK(rs != nullptr && r != nullptr);
        if (!rs || !r) continue;
        CHECK(rs->samples_until_steady_state() == lo_resampler_samples_until_steady_state(r));
        for (int n : {a / 50, 33, 1, a / 50 + 5}) {
          std::vector<int16_t> x((size_t)n);
          for (auto& v : x) v = (int16_t)(d(rng) * 3);
          const std::vector<int16_t> got = rs->Resample(x);
          std::vector<int16_t> want((size_t)n * 3 + 8);
          const int m = lo_resampler_resample(r, x.data(), n, want.data(), (int)want.size());
          CHECK((int)got.size() == m && std::memcmp(got.data(), want.data(), sizeof(int16_t) * (size_t)m) == 0);
        }
        lo_resampler_free(r);
      }
      auto e = LyraEncoderB200::Create(rate, 1, 3200, false, model);
      auto dcd = LyraDecoderB200::Create(rate, 1, model);
      CHECK(e && dcd && e->sample_rate_hz() == rate && dcd->sample_rate_hz() == rate);
      if (!e || !dcd) continue;
      const int hop = rate / 50;
      CHECK(!e->Encode(std::vector<int16_t>(320 + (rate == 16000))).has_value() || rate == 16000);   // a 16 kHz hop is the wrong size here
      for (int f = 0; f < 4; ++f) {
        std::vector<int16_t> pcm((size_t)hop);
        for (auto& v : pcm) v = (int16_t)d(rng);
        auto pkt = e->Encode(pcm);
        CHECK(pkt.has_value() && pkt->size() == 8);
        CHECK(dcd->SetEncodedPacket(*pkt));
        auto a1 = dcd->DecodeSamples(hop - 7);                     // odd request sizes exercise the leftover buffer
        auto a2 = dcd->DecodeSamples(7);
        CHECK(a1.has_value() && a2.has_value() && (int)a1->size() == hop - 7 && a2->size() == 7);
      }
    }
  }
  // coalescing front: one thread per stream, each with its own SoundStreamEncoder / quantizer / LyraGanModel objects, calling
  // concurrently; calls of a kind are merged into batched launches, results stay those of the per-stream oracle
  {
    const int kThreads = 6, kFrames = 3;
    auto es = Session::Get(model, LYRA_B200_ROLE_ENCODER);
    auto ds = Session::Get(model, LYRA_B200_ROLE_DECODER);
    CHECK(es && ds);
    const Coalescer::Stats e0 = es->coalescer_stats(), d0 = ds->coalescer_stats();
    es->set_coalesce_linger_us(2000);
    ds->set_coalesce_linger_us(2000);
    std::vector<int> fails((size_t)kThreads, 0);
    std::vector<std::thread> pool;
    for (int t = 0; t < kThreads; ++t) pool.emplace_back([&, t] {
      int& bad = fails[(size_t)t];
      auto e = SoundStreamEncoderB200::Create(model);
      auto q = ResidualVectorQuantizerB200::Create(model, t % 2 ? LYRA_B200_ROLE_DECODER : LYRA_B200_ROLE_ENCODER);
      auto g = LyraGanModelB200::Create(model, 64);
      lo_codec* r = lo_codec_create(model.c_str());
      if (!e || !q || !g || !r) { bad = 100; return; }
      std::mt19937 trng(100 + (unsigned)t);
      std::uniform_int_distribution<int> td(-8192, 8191);
      for (int f = 0; f < kFrames; ++f) {
        std::vector<int16_t> pcm(320);
        for (auto& v : pcm) v = (int16_t)td(trng);
        const int bits = t % 3 == 0 ? 64 : 120;                // two batch keys for the quantizer calls
        uint8_t rp[24];
        int16_t rpcm[320];
        lo_codec_encode(r, pcm.data(), bits, rp, nullptr, nullptr);
        lo_codec_decode(r, rp, bits, rpcm, nullptr, nullptr);
        auto feat = e->Extract(pcm);
        if (!feat) { ++bad; continue; }
        if (t == 5) bad += q->Quantize(*feat, 62).has_value();   // a refused call (its own batch key) fails alone, the others' launches go on
        auto str = q->Quantize(*feat, bits);
        if (!str) { ++bad; continue; }
        const std::vector<uint8_t> pkt = Packet184::PackQuantized(*str);
        bad += std::memcmp(pkt.data(), rp, pkt.size()) != 0;
        auto lossy = q->DecodeToLossyFeatures(*str);
        if (!lossy || !g->AddFeatures(*lossy)) { ++bad; continue; }
        auto out = g->GenerateSamples(320);
        bad += !out || std::memcmp(out->data(), rpcm, sizeof(rpcm)) != 0;
      }
      lo_codec_free(r);
    });
    for (auto& th : pool) th.join();
    for (int t = 0; t < kThreads; ++t) CHECK(fails[(size_t)t] == 0);
    es->set_coalesce_linger_us(0);
    ds->set_coalesce_linger_us(0);
    const Coalescer::Stats e1 = es->coalescer_stats(), d1 = ds->coalescer_stats();
    std::printf("coalescer: encoder context %llu calls in %llu launches (largest batch %llu), decoder context %llu in %llu (largest %llu)\n",
                (unsigned long long)(e1.calls - e0.calls), (unsigned long long)(e1.launches - e0.launches), (unsigned long long)e1.max_batch,
                (unsigned long long)(d1.calls - d0.calls), (unsigned long long)(d1.launches - d0.launches), (unsigned long long)d1.max_batch);
    CHECK(e1.max_batch >= 2 && d1.max_batch >= 2);             // concurrent calls did share launches
    CHECK(e1.launches - e0.launches < e1.calls - e0.calls);
  }
  std::printf(g_fail ? "FAILED (%d)\n" : "ALL OK\n", g_fail);
  return g_fail ? 1 : 0;
}
