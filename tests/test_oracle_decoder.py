"""CPU tier: the oracle's LyraDecoder state machine, comfort-noise generator and DTX encoder against the reference's own tests.

The state-machine cases restate lyra/lyra_decoder_test.cc (fake generative models that return -10000 / +10000, gmock
expectations replaced by the oracle's call counters); the generator cases restate lyra/comfort_noise_generator_test.cc."""
import numpy as np
import pytest

from conftest import MODEL_DIR

HOP, CONCEAL, FADE = 320, 1280, 640
GEN, COMFORT = -10000, 10000
FROM_CNG, TO_CNG = -1, 1
ZERO_PACKET = bytes(8)


def diff(a, b):
    return {k: b[k] - a[k] for k in a if not k.endswith("last_request")}


@pytest.fixture
def dec(oracle):
    return oracle.Decoder(fake=(GEN, COMFORT))


def test_normal_to_concealment_to_normal(dec):
    # lyra_decoder_test.cc:371-410 EntirePacketRequests_NormalToConcealmentToNormal
    c0 = dec.counters()
    assert dec.set_encoded_packet(ZERO_PACKET)
    out = dec.decode_samples(HOP)
    c1 = dec.counters()
    assert (out == GEN).all() and len(out) == HOP
    # ExpectSetEncodedPacket(1) + ExpectNormalDecoding: one quantizer call, one AddFeatures, model asked for the hop, the comfort
    # noise generator asked for 0 samples, no noise estimate read, the estimator fed once
    assert diff(c0, c1) == dict(vq_decode=1, model_add=1, model_generate=1, cng_add=0, cng_generate=1, noise_receive=1, noise_estimate=0)
    assert c1["model_last_request"] == HOP and c1["cng_last_request"] == 0
    out = dec.decode_samples(HOP)                                  # state 2: concealment with estimated (zero) features
    c2 = dec.counters()
    assert (out == GEN).all()
    assert diff(c1, c2) == dict(vq_decode=0, model_add=1, model_generate=1, cng_add=0, cng_generate=1, noise_receive=0, noise_estimate=0)
    assert dec.set_encoded_packet(ZERO_PACKET)
    out = dec.decode_samples(HOP)
    c3 = dec.counters()
    assert (out == GEN).all()
    assert diff(c2, c3) == dict(vq_decode=1, model_add=1, model_generate=1, cng_add=0, cng_generate=1, noise_receive=1, noise_estimate=0)


def test_concealment_to_comfort_noise(dec):
    # lyra_decoder_test.cc:414-478 TestFinishDecoding_ConcealmentToComfortNoise
    for i in range(CONCEAL // HOP):                                # state 2
        c = dec.counters()
        out = dec.decode_samples(HOP)
        assert (out == GEN).all()
        assert diff(c, dec.counters()) == dict(vq_decode=0, model_add=1, model_generate=1, cng_add=0, cng_generate=1, noise_receive=0, noise_estimate=0)
        assert dec.counters()["cng_last_request"] == 0
    prev_mean = GEN
    for i in range(FADE // HOP):                                   # state 3: fade to comfort noise
        c = dec.counters()
        out = dec.decode_samples(HOP)
        assert diff(c, dec.counters()) == dict(vq_decode=0, model_add=1, model_generate=1, cng_add=1, cng_generate=1, noise_receive=0, noise_estimate=1)
        assert ((out >= GEN) & (out <= COMFORT)).all() and (np.diff(out.astype(np.int32)) >= 0).all()
        assert out.mean() > prev_mean
        prev_mean = out.mean()
        assert not dec.is_comfort_noise() or i == FADE // HOP - 1
    assert dec.is_comfort_noise()
    for i in range(3):                                             # state 4: pure comfort noise
        c = dec.counters()
        out = dec.decode_samples(HOP)
        assert (out == COMFORT).all()
        assert diff(c, dec.counters()) == dict(vq_decode=0, model_add=0, model_generate=1, cng_add=1, cng_generate=1, noise_receive=0, noise_estimate=1)
        assert dec.counters()["model_last_request"] == 0


def test_comfort_noise_fade_to_normal(dec):
    # lyra_decoder_test.cc:482-552 TestFinishDecoding_ComfortNoiseFadetoNormal
    dec.state = (CONCEAL, FADE, TO_CNG)
    c = dec.counters()
    out = dec.decode_samples(100)                                  # partially decode a comfort-noise packet
    assert (out == COMFORT).all() and len(out) == 100
    assert diff(c, dec.counters()) == dict(vq_decode=0, model_add=0, model_generate=1, cng_add=1, cng_generate=1, noise_receive=0, noise_estimate=1)
    assert dec.set_encoded_packet(ZERO_PACKET)
    assert dec.state[0] == -(HOP - 100)                            # the rest of the fake packet is played out first
    c = dec.counters()
    out = dec.decode_samples(HOP - 100)
    assert (out == COMFORT).all()
    assert diff(c, dec.counters()) == dict(vq_decode=0, model_add=0, model_generate=1, cng_add=0, cng_generate=1, noise_receive=0, noise_estimate=0)
    prev = COMFORT
    for i in range(FADE // HOP):                                   # state 5: fade to normal decoding
        if i > 0:
            assert dec.set_encoded_packet(ZERO_PACKET)
        c = dec.counters()
        out = dec.decode_samples(HOP)
        d = diff(c, dec.counters())
        assert d["model_generate"] == 1 and d["cng_add"] == 1 and d["cng_generate"] == 1 and d["noise_estimate"] == 1 and d["noise_receive"] == 1
        assert ((out >= GEN) & (out <= COMFORT)).all() and (np.diff(out.astype(np.int32)) <= 0).all()
        assert out.mean() < prev
        prev = out.mean()
    assert dec.state == (0, 0, FROM_CNG)
    assert dec.set_encoded_packet(ZERO_PACKET)
    assert (dec.decode_samples(HOP) == GEN).all()                  # back in state 1


def test_multiple_hops_one_request(dec):
    # lyra_decoder_test.cc:554-580 MultipleHopsOneRequestNormalDecode
    for _ in range(4):
        assert dec.set_encoded_packet(ZERO_PACKET)
    c = dec.counters()
    out = dec.decode_samples(4 * HOP)
    assert len(out) == 4 * HOP and (out == GEN).all()
    d = diff(c, dec.counters())
    assert d["noise_receive"] == 4 and d["model_generate"] == 4 and d["model_add"] == 0


def test_hops_are_overlapped_correctly(dec):
    # lyra_decoder_test.cc:582-686 HopsAreOverlappedCorrectly at the internal rate (no resampler transient)
    out = dec.decode_samples(HOP)
    assert (out == GEN).all()                                      # state 2
    dec.state = (CONCEAL, 0, TO_CNG)                               # state 3
    out = dec.decode_samples(HOP).astype(np.int32)
    assert (out >= GEN).all() and (out <= COMFORT).all() and (np.diff(out) >= -2).all()
    assert out[0] == GEN                                           # weight (1 + cos 0) / 2 = 1 at fade progress 0
    dec.state = (CONCEAL, FADE, TO_CNG)                            # state 4
    assert (dec.decode_samples(HOP) == COMFORT).all()
    dec.state = (0, FADE, FROM_CNG)                                # state 5 needs a received packet
    assert dec.set_encoded_packet(ZERO_PACKET)
    out = dec.decode_samples(HOP).astype(np.int32)
    assert (out >= GEN).all() and (out <= COMFORT).all() and (np.diff(out) <= 2).all()


def test_arbitrary_num_samples(dec):
    # lyra_decoder_test.cc:688-760 ArbitraryNumSamples{NormalDecode,Concealment,ComfortNoise}: any request size is served
    for n in range(0, HOP, 37):
        dec.state = (0, 0, FROM_CNG)
        assert dec.set_encoded_packet(ZERO_PACKET)
        out = dec.decode_samples(n)
        assert len(out) == n and (out == GEN).all()
        rest = dec.decode_samples(HOP - n)                         # finish the hop so the next iteration starts aligned
        assert len(rest) == HOP - n
    d2 = type(dec)(fake=(GEN, COMFORT))
    total = 0
    for n in [1, 50, 319, 320, 321, 640, 7]:
        out = d2.decode_samples(n)
        assert len(out) == n
        total += n
    assert d2.decode_samples(-1) is None


def test_bad_packets_are_rejected(dec):
    # lyra_decoder_test.cc: packets of an unsupported size are refused (SetEncodedPacket :173-178)
    assert not dec.set_encoded_packet(bytes(7))
    assert not dec.set_encoded_packet(b"")
    assert dec.set_encoded_packet(bytes(15)) and dec.set_encoded_packet(bytes(23))


# ---------------------------------------------------------------- comfort noise generator ----

def test_cng_sample_requests(oracle):
    # comfort_noise_generator_test.cc:42-56 NumSamplesRequestedOutOfBounds
    g = oracle.ComfortNoiseGenerator()
    assert g.add_features(np.zeros(160, np.float32))
    assert g.generate_samples(HOP + 1) is None
    assert g.generate_samples(-1) is None
    assert len(g.generate_samples(0)) == 0


def test_cng_feature_counts(oracle):
    # comfort_noise_generator_test.cc:58-84 SamplesGeneratedOnlyWithCorrectNumFeatures
    g = oracle.ComfortNoiseGenerator()
    assert not g.add_features(np.zeros(0, np.float32))
    assert g.generate_samples(HOP) is None
    assert not g.add_features(np.ones(159, np.float32)) and g.generate_samples(HOP) is None
    assert not g.add_features(np.ones(161, np.float32)) and g.generate_samples(HOP) is None
    assert g.add_features(np.ones(160, np.float32)) and g.generate_samples(HOP) is not None


def test_cng_silence_in_silence_out(oracle):
    # comfort_noise_generator_test.cc:86-98 BasicUseCaseSucceeds
    g = oracle.ComfortNoiseGenerator()
    assert g.add_features(np.zeros(160, np.float32))
    assert (g.generate_samples(HOP) == 0).all()


@pytest.mark.parametrize("seed", [1, 2, 3, 11, 12345])
def test_cng_generated_noise_has_similar_features(oracle, seed):
    # comfort_noise_generator_test.cc:100-138 GeneratedNoiseHasSimilarFeatures: THE criterion that pins the restated
    # audio_dsp pieces (mel inverse + inverse spectrogram + their gain) - log-spectral distance < 0.7 after 10 hops
    rng = np.random.default_rng(seed)
    x = rng.integers(-10000, 10001, size=HOP).astype(np.int16)     # std::uniform_int_distribution<int16_t>(-10000, 10000)
    ie, oe = oracle.LogMel(16000, HOP, 640, 160), oracle.LogMel(16000, HOP, 640, 160)
    g = oracle.ComfortNoiseGenerator(seed=seed)
    for _ in range(10):
        fi = ie.extract(x)
        assert g.add_features(fi)
        fo = oe.extract(g.generate_samples(HOP))
    assert oracle.log_spectral_distance(fi, fo) < 0.7


def test_cng_is_deterministic_per_seed_and_differs_across_seeds(oracle):
    f = np.full(160, 0.9, np.float32)
    a, b, c = (oracle.ComfortNoiseGenerator(seed=s) for s in (5, 5, 6))
    outs = []
    for g in (a, b, c):
        g.add_features(f)
        g.add_features(f)
        outs.append(np.concatenate([g.generate_samples(HOP), g.generate_samples(HOP)]))
    assert np.array_equal(outs[0], outs[1]) and not np.array_equal(outs[0], outs[2])
    assert not np.array_equal(outs[0][:HOP], outs[0][HOP:])         # the hop counter advances the phases
    ph = [oracle.cng_phase_index(5, 0, i) for i in range(513)]
    assert min(ph) >= 0 and max(ph) < 1024 and len(set(ph)) > 200


# ---------------------------------------------------------------- real components end to end ----

def test_real_decoder_goes_to_comfort_noise_and_back(oracle, sample1):
    """16 kHz speech through the DTX-less encoder, a 12-hop outage, recovery: the decoder conceals for 4 hops, fades to comfort
    noise, plays comfort noise, fades back; the whole-hop recurrence the batched GPU step implements."""
    enc = oracle.Encoder(MODEL_DIR)
    dec = oracle.Decoder(MODEL_DIR, cng_seed=7)
    states = []
    for f in range(40):
        pkt = enc.encode(sample1[f * HOP:(f + 1) * HOP], 64)
        if not 15 <= f < 27:
            assert dec.set_encoded_packet(pkt)
        out = dec.decode_samples(HOP)
        assert len(out) == HOP
        states.append(dec.state + (dec.is_comfort_noise(),))
    assert states[14][:2] == (0, 0)
    assert [s[0] for s in states[15:19]] == [320, 640, 960, 1280]
    assert states[19][1:3] == (320, TO_CNG) and states[20][1:3] == (640, TO_CNG) and states[20][3]
    assert all(s[3] for s in states[20:27])
    assert states[27][:3] == (0, 320, FROM_CNG) and states[28][:3] == (0, 0, FROM_CNG)


def test_dtx_encoder_sends_empty_packets_for_noise(oracle, sample1):
    # lyra_encoder.cc:131-141: with DTX a hop classified as noise becomes an empty packet; speech hops are encoded as usual
    dtx, plain = oracle.Encoder(MODEL_DIR, enable_dtx=True), oracle.Encoder(MODEL_DIR)
    sizes = []
    for f in range(60):
        hop = sample1[f * HOP:(f + 1) * HOP] if f >= 30 else np.zeros(HOP, np.int16)      # digital silence, then speech
        p = dtx.encode(hop, 64)
        sizes.append(len(p))
        if len(p):
            assert len(p) == 8
    assert sizes[0] == 8                     # the first hop lies outside the zero bounds of a fresh estimator: it is encoded
    assert sizes[1:30] == [0] * 29 and 8 in sizes[30:]
    assert dtx.encode(sample1[:100], 64) is None and plain.encode(sample1[:HOP], 64) is not None
