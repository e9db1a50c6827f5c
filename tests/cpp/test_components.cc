// GPU tier: the C++ adapters of include/lyra_b200/lyra_b200_components.h (the reference's plugin surface) against the
// CPU oracle, in the style of the reference's own unit tests (soundstream_encoder_test.cc, residual_vector_quantizer_test.cc,
// lyra_gan_model_test.cc, lyra_integration_test.cc).  Links liblyra_b200.so (product) and liblyra_oracle.so (checker).
#include <cmath>
#include <cstdio>
#include <cstring>
#include <random>
#include <string>
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
  std::printf(g_fail ? "FAILED (%d)\n" : "ALL OK\n", g_fail);
  return g_fail ? 1 : 0;
}
