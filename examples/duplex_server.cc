// Example host program (C++17, links liblyra_b200.so): a full-duplex codec server loop over the C ABI.
//
// The streams are divided among G worker groups.  Each group owns an encoder-only and a decoder-only context
// (lyra_b200_create_ex) and two threads: the uplink thread encodes one 20 ms hop of every stream of the group per step
// (lyra_b200_encode: host PCM in, host packets out), the downlink thread decodes the packets of the same step
// (lyra_b200_decode: host packets in, host PCM out).  Calls on one context are serialised (one thread per context);
// different contexts run concurrently, so encode kernels of step i+1 overlap decode kernels of step i on the GPU.
// This is the C++ counterpart of bench.py's host-buffer pass and of what LyraEncoder / LyraDecoder pairs do per stream
// in the reference (lyra/lyra_encoder.h, lyra/lyra_decoder.h).
//
//   duplex_server <model_dir> <streams> <steps> [groups=2] [bits=64]
// prints frames/s and a checksum of the decoded audio of the last step (tests compare it with the oracle's).
// It is an illustration, not the benchmark: the input is synthesised inside the uplink thread and the host buffers are
// ordinary pageable vectors (bench.py measures the same loop with pinned buffers and precomputed input).
#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <thread>
#include <vector>

#include "lyra_b200.h"

namespace {

// single-producer single-consumer hand-over of step numbers with a bounded number of packet buffers in flight
class StepQueue {
 public:
  explicit StepQueue(int slots) : free_(slots) {}
  void AcquireSlot() { std::unique_lock<std::mutex> l(m_); cv_.wait(l, [&] { return free_ > 0; }); --free_; }
  void Publish() { { std::lock_guard<std::mutex> l(m_); ++ready_; } cv_.notify_all(); }
  void WaitReady() { std::unique_lock<std::mutex> l(m_); cv_.wait(l, [&] { return ready_ > 0; }); --ready_; }
  void ReleaseSlot() { { std::lock_guard<std::mutex> l(m_); ++free_; } cv_.notify_all(); }

 private:
  std::mutex m_;
  std::condition_variable cv_;
  int free_, ready_ = 0;
};

// synthetic input of (stream, step): a 32-bit LCG, samples uniform in [-8192, 8191] (a quarter of full scale)
void FillHop(int16_t* dst, int global_stream, int step) {
  uint32_t x = 2463534242u ^ (uint32_t)(global_stream * 7919 + step * 104729);
  for (int i = 0; i < LYRA_B200_HOP; ++i) {
    x = x * 1664525u + 1013904223u;
    dst[i] = (int16_t)((int)((x >> 16) & 16383u) - 8192);
  }
}

struct Group {
  lyra_b200_ctx* enc = nullptr;
  lyra_b200_ctx* dec = nullptr;
  int n = 0, first_stream = 0;
  std::vector<std::vector<int16_t>> pcm_in;     // [slot][n * 320]
  std::vector<std::vector<uint8_t>> packets;    // [slot][n * packet_bytes]
  std::vector<int16_t> pcm_out;                 // [n * 320]
  int failed = 0;
};

}  // namespace

int main(int argc, char** argv) {
  if (argc < 4) { std::fprintf(stderr, "usage: %s <model_dir> <streams> <steps> [groups=2] [bits=64]\n", argv[0]); return 2; }
  const char* model_dir = argv[1];
  const int streams = std::atoi(argv[2]), steps = std::atoi(argv[3]);
  int groups = argc > 4 ? std::atoi(argv[4]) : 2;
  const int bits = argc > 5 ? std::atoi(argv[5]) : 64;
  if (streams <= 0 || steps <= 0 || groups <= 0 || streams % groups) { std::fprintf(stderr, "streams must be a positive multiple of groups\n"); return 2; }
  const int kSlots = 4, pb = (bits + 7) / 8, per = streams / groups;

  std::vector<Group> g((size_t)groups);
  for (int k = 0; k < groups; ++k) {
    Group& x = g[(size_t)k];
    x.n = per;
    x.first_stream = k * per;
    if (lyra_b200_create_ex(model_dir, 0, per, LYRA_B200_ROLE_ENCODER, &x.enc) != LYRA_B200_OK ||
        lyra_b200_create_ex(model_dir, 0, per, LYRA_B200_ROLE_DECODER, &x.dec) != LYRA_B200_OK) {
      std::fprintf(stderr, "cannot create contexts: %s\n", lyra_b200_last_error(nullptr));
      return 1;
    }
    x.pcm_in.assign(kSlots, std::vector<int16_t>((size_t)per * LYRA_B200_HOP));
    x.packets.assign(kSlots, std::vector<uint8_t>((size_t)per * pb));
    x.pcm_out.assign((size_t)per * LYRA_B200_HOP, 0);
  }

  const auto t0 = std::chrono::steady_clock::now();
  std::vector<std::thread> threads;
  std::vector<StepQueue*> queues;
  for (int k = 0; k < groups; ++k) {
    Group* x = &g[(size_t)k];
    StepQueue* q = new StepQueue(kSlots);
    queues.push_back(q);
    threads.emplace_back([=] {                                            // uplink
      for (int i = 0; i < steps; ++i) {
        q->AcquireSlot();
        const int slot = i % kSlots;
        for (int s = 0; s < x->n; ++s) FillHop(x->pcm_in[(size_t)slot].data() + (size_t)s * LYRA_B200_HOP, x->first_stream + s, i);
        if (lyra_b200_encode(x->enc, nullptr, x->n, x->pcm_in[(size_t)slot].data(), bits, x->packets[(size_t)slot].data()) != LYRA_B200_OK) x->failed = 1;
        q->Publish();
      }
    });
    threads.emplace_back([=] {                                            // downlink
      for (int i = 0; i < steps; ++i) {
        q->WaitReady();
        const int slot = i % kSlots;
        if (lyra_b200_decode(x->dec, nullptr, x->n, x->packets[(size_t)slot].data(), nullptr, bits, x->pcm_out.data()) != LYRA_B200_OK) x->failed = 1;
        q->ReleaseSlot();
      }
    });
  }
  for (auto& t : threads) t.join();
  const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();

  long long checksum = 0;
  int failed = 0;
  for (auto& x : g) {
    for (int16_t v : x.pcm_out) checksum += v;
    failed |= x.failed;
    if (x.failed) std::fprintf(stderr, "a call failed: %s / %s\n", lyra_b200_last_error(x.enc), lyra_b200_last_error(x.dec));
    lyra_b200_destroy(x.enc);
    lyra_b200_destroy(x.dec);
  }
  for (auto* q : queues) delete q;
  std::printf("streams %d steps %d groups %d bits %d: %.0f frames/s, checksum %lld\n", streams, steps, groups, bits,
              (double)streams * steps / secs, checksum);
  return failed ? 1 : 0;
}
