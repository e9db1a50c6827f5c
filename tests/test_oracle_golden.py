"""Pins the CPU oracle against every fixture the reference's own tests hold for the hot path
(SURVEY.md §8c), plus regression vectors of the oracle itself (tests/golden/, drift detection).
CPU only."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR, MODEL_DIR

# ---- lyra/log_mel_spectrogram_extractor_impl_test.cc:37-59 -------------------------------------
K_WAV = np.array([7954, 10085, 8733, 10844, 29949, -549, 20833, 30345, 18086, 11375,
                  -27309, 12323, -22891, -23360, 11958], dtype=np.int16)
K_MEL = np.array([
    [0.62146081, 0.62146081, 0.79771997, 1.00416802, 0.73013308, 0.96676503, 0.87643814, 0.89284485, 0.90586112, 0.8633126],
    [0.62146081, 0.62146081, 0.89000145, 1.09644949, 0.76740002, 1.00403196, 0.8919037, 0.99746922, 1.06052462, 1.08220812],
    [0.62146081, 0.62146081, 0.83526758, 1.04171563, 0.82093681, 1.05756876, 0.96348656, 1.01345318, 1.07686605, 1.12100911]],
    dtype=np.float32)

# ---- lyra/residual_vector_quantizer_test.cc:41-54 ----------------------------------------------
RVQ_FIXTURE = np.array([
    5.18127, 0.156109, -0.875549, 1.90394, 4.27785, 0.184078, 2.03794, 0.895547, 6.61436, 3.61373, 1.84045, 2.34979,
    1.91443, 2.46864, 2.49996, -0.78883, 2.04522, -0.0539977, -0.206427, -0.856873, 1.56033, 1.48176, 1.82138, 0.900604,
    -0.10602, -0.548707, 0.33733, 7.63183, -0.199688, 6.35543, 2.47549, -0.854709, 0.0588712, -0.144105, 7.68603, 2.78211,
    1.89553, 1.46111, 1.60068, -0.310399, 1.4651, 2.05484, 0.460265, 1.88702, -0.186116, 0.134471, -0.304016, 0.924312,
    9.56944, 0.877297, 0.825455, 2.45036, 2.36505, 1.02132, 2.03803, 0.308894, -0.930119, 3.16624, -0.743392, 0.137643,
    2.01814, 3.39578, 4.30634, 0.880378], dtype=np.float32)


def float_eq(a, b, ulps=4):
    """gtest FloatEq: within 4 ULPs."""
    ai = np.asarray(a, dtype=np.float32).view(np.int32).astype(np.int64)
    bi = np.asarray(b, dtype=np.float32).view(np.int32).astype(np.int64)
    return np.all(np.abs(ai - bi) <= ulps)


def test_logmel_known_answer(oracle):
    lm = oracle.LogMel(16000, 5, 10, 10)
    for i in range(3):
        out = lm.extract(K_WAV[5 * i:5 * i + 5])
        assert float_eq(out, K_MEL[i]), (i, out, K_MEL[i])


def test_logmel_rejects_wrong_hop_and_window(oracle):
    lm = oracle.LogMel(16000, 5, 10, 10)
    assert lm.extract(np.zeros(6, np.int16)) is None            # _test.cc:86-92
    assert lm.extract(np.zeros(4, np.int16)) is None
    with pytest.raises(ValueError):
        oracle.LogMel(16000, 10, 5, 10)                          # window < hop -> Create returns nullptr


def test_logmel_silence_value(oracle):
    lm = oracle.LogMel(16000, 320, 640, 160)
    out = lm.extract(np.zeros(320, np.int16))
    assert np.allclose(out, np.log(np.float32(500.0)) / 10.0, atol=1e-7)   # GetSilenceValue


# ---- lyra/packet_test.cc ------------------------------------------------------------------------
def contains_quantized(packet, bits, nh, nq):
    s = "".join(format(b, "08b") for b in packet)
    return s[nh:nh + nq] == bits


def test_packet_unpack_vectors(oracle):
    assert oracle.packet_unpack(bytes([0b00011111, 0b11111111, 0b11100000]), 3, 16) == "1" * 16      # :93-111
    assert oracle.packet_unpack(bytes([0b00111111, 0xFF, 0xFF]), 2, 22) == "1" * 22                   # :113-128
    assert oracle.packet_unpack(bytes([0, 0, 0b00000011, 0xFF, 0xFF, 0b11110000]), 22, 22) == "1" * 22  # :130-148
    nh, nq, size = 8, 104, 14
    for enc in ([0] + [0xFF] * 13, [0] + [0b10101010] * 13, [0 if i % 2 == 1 or i == 0 else 0 for i in range(14)]):
        bits = oracle.packet_unpack(bytes(enc), nh, nq)
        assert contains_quantized(enc, bits, nh, nq)
    alt = [0] * 14
    for i in range(2, 14, 2):
        alt[i] = 0xFF
    assert contains_quantized(alt, oracle.packet_unpack(bytes(alt), nh, nq), nh, nq)
    assert oracle.packet_unpack(bytes([0xFF] * (size - 1)), nh, nq) is None                            # InvalidPacketSize


def test_packet_pack_vectors(oracle):
    nq = 104
    ones = "1" * nq
    alt = "".join("1" if (nq - 1 - i) % 2 == 0 else "0" for i in range(nq))        # bit i set for even i (LSB = index 0)
    bytes_alt = "".join("1" if ((nq - 1 - i) // 8) % 2 == 1 else "0" for i in range(nq))
    for bits in (ones, alt, bytes_alt):
        for nh in (8, 10, 100):
            enc = oracle.packet_pack(bits, nh, nq)
            assert len(enc) == -(-(nh + nq) // 8)
            assert contains_quantized(enc, bits, nh, nq)
            assert oracle.packet_unpack(enc, nh, nq) == bits
    assert oracle.packet_size(7, 52) == 8


def test_packet_sizes_match_lyra_config(oracle):
    # lyra/lyra_config.cc:44-48, lyra_config.h:79-91: 64/120/184 bits <-> 8/15/23 bytes <-> 3200/6000/9200 bps
    for bits, size, rate in ((64, 8, 3200), (120, 15, 6000), (184, 23, 9200)):
        assert oracle.packet_size(0, bits) == size
        assert size * 8 * 50 == rate


# ---- lyra/dsp_utils_test.cc / dsp_utils.h -------------------------------------------------------
def test_dsp_conversions(oracle):
    assert oracle.int16_to_unit(-32768) == -1.0 and oracle.int16_to_unit(16384) == 0.5 and oracle.int16_to_unit(0) == 0.0
    assert oracle.unit_to_int16(1.0) == 32767 and oracle.unit_to_int16(-1.5) == -32768
    assert oracle.unit_to_int16(0.5) == 16384 and oracle.unit_to_int16(-0.99999) == -32767   # truncation toward zero
    a = np.arange(8, dtype=np.float32)
    assert oracle.log_spectral_distance(a, a) == 0.0
    assert abs(oracle.log_spectral_distance(a, a + 1.0) - 10.0) < 1e-6


# ---- lyra/residual_vector_quantizer_test.cc -----------------------------------------------------
@pytest.mark.parametrize("bits,expected", [(64, 1.1073), (120, 0.7393), (184, 0.5108)])
def test_rvq_round_trip_distance(oracle, bits, expected):
    rvq = oracle.Rvq(os.path.join(MODEL_DIR, "quantizer.tflite"))
    q = rvq.quantize(RVQ_FIXTURE, bits)
    assert q is not None and len(q) == bits and set(q) <= {"0", "1"}
    d = rvq.decode_to_lossy_features(q)
    dist = float(np.sqrt(((RVQ_FIXTURE - d) ** 2).sum() / (RVQ_FIXTURE ** 2).sum()))
    assert dist < 1.11                                     # the reference's assertion (:104-111)
    assert abs(dist - expected) < 1e-3                     # SURVEY.md §0.3(d) cross-check value


def test_rvq_argument_errors(oracle):
    rvq = oracle.Rvq(os.path.join(MODEL_DIR, "quantizer.tflite"))
    assert rvq.quantize(RVQ_FIXTURE, 185) is None          # QuantizationFailsWithTooManyBits
    assert rvq.quantize(RVQ_FIXTURE, 62) is None           # QuantizationFailsWithNonDivisibleBits
    assert rvq.decode_to_lossy_features("0" * 185) is None
    assert rvq.decode_to_lossy_features("0" * 62) is None
    assert rvq.num_stages == 46 and rvq.bits_per_stage == 4


def test_rvq_structure(oracle):
    rvq = oracle.Rvq(os.path.join(MODEL_DIR, "quantizer.tflite"))
    idx = rvq.encode(RVQ_FIXTURE, 46)
    assert idx[:16].tolist() == [8, 11, 14, 1, 6, 13, 12, 11, 1, 15, 6, 12, 4, 3, 3, 0]
    # unused stages report -1 and decode to a zero contribution
    idx16 = rvq.encode(RVQ_FIXTURE, 16)
    assert (idx16[:16] == idx[:16]).all() and (idx16[16:] == -1).all()
    cb = rvq.codebooks()
    manual = np.zeros(64, np.float32)
    for s in range(16):
        manual = manual + cb[s, idx16[s]] if s else cb[s, idx16[s]].copy()
    assert np.array_equal(rvq.decode(idx16), manual)


# ---- conv nets ------------------------------------------------------------------------------------
def test_encoder_output_lattice_and_cross_check(oracle):
    c = oracle.Codec(MODEL_DIR)
    _, feat, _ = c.encode(np.zeros(320, np.int16), 64)
    q = feat / np.float32(0.263492) + 20
    assert np.allclose(q, np.round(q), atol=1e-3)          # output lives on the int8 lattice
    assert np.round(q).astype(int)[:16].tolist() == [34, 63, 63, 41, 22, -36, 85, 89, 39, -19, 7, 33, 19, 7, 8, 52]


def test_shapes_and_state_inventory(oracle):
    # soundstream_encoder_test.cc:51-57, lyra_gan_model_test.cc:60-76: 320 -> 64, 64 -> 320
    enc = oracle.Net(os.path.join(MODEL_DIR, "soundstream_encoder.tflite"))
    dec = oracle.Net(os.path.join(MODEL_DIR, "lyragan.tflite"))
    assert enc.invoke(np.zeros(320, np.float32), 64).shape == (64,)
    assert dec.invoke(np.zeros(64, np.float32), 320).shape == (320,)
    ev, dv = enc.variables(), dec.variables()
    assert len(ev) == 14 and sum(v.size for v in ev.values()) == 13808
    assert len(dv) == 18 and sum(v.size for v in dv.values()) == 12912
    with pytest.raises(RuntimeError):
        enc.invoke(np.zeros(319, np.float32), 64)


def test_reset_restores_initial_state(oracle, sample1):
    c = oracle.Codec(MODEL_DIR)
    first = [c.encode(sample1[320 * h:320 * h + 320], 64)[0] for h in range(3)]
    c.reset()
    again = [c.encode(sample1[320 * h:320 * h + 320], 64)[0] for h in range(3)]
    assert first == again


@pytest.mark.parametrize("bits", [64, 120, 184])
def test_integration_log_spectral_distance(oracle, sample1, bits):
    """lyra/lyra_integration_test.cc:49-143 at 16 kHz: every hop LSD < 2.0 over the first 3 s."""
    c = oracle.Codec(MODEL_DIR)
    hops = min(150, len(sample1) // 320)
    a, b = oracle.LogMel(16000, 320, 640, 64), oracle.LogMel(16000, 320, 640, 64)
    worst = 0.0
    for h in range(hops):
        x = sample1[320 * h:320 * h + 320]
        pkt, _, _ = c.encode(x, bits)
        assert len(pkt) == oracle.packet_size(0, bits)
        y, _, _ = c.decode(pkt, bits)
        worst = max(worst, oracle.log_spectral_distance(a.extract(x), b.extract(y)))
    assert worst < 2.0
    assert abs(worst - {64: 1.160, 120: 0.801, 184: 1.123}[bits]) < 5e-3    # SURVEY.md §0.3(c) cross-check


def test_golden_regression_vectors(oracle, sample1):
    """tests/golden/oracle_sample1.json (made by tests/golden/make_golden.py from this oracle): drift detection."""
    with open(os.path.join(GOLDEN_DIR, "oracle_sample1.json")) as f:
        g = json.load(f)
    assert g["packets_64"][0] == "a00809827516b2df" and g["packets_64"][1] == "a6830dde77dcb60b"
    for bits in (64, 120, 184):
        c = oracle.Codec(MODEL_DIR)
        for h, want in enumerate(g["packets_%d" % bits]):
            pkt, _, _ = c.encode(sample1[320 * h:320 * h + 320], bits)
            assert pkt.hex() == want, (bits, h)
            pcm, _, _ = c.decode(pkt, bits)
            assert int(np.int64(pcm.astype(np.int64) * np.arange(1, 321)).sum()) == g["pcm_checksum_%d" % bits][h]


# ---- NoiseEstimator: the reference's own tests are statistical; their properties are
#      re-run here against the restatement (lyra/noise_estimator_test.cc:131-197)
def _silence_value():
    return np.float32(np.log(np.float32(500.0)) / np.float32(10.0))      # GetSilenceValue, log_mel_..._impl.cc:138-140


def test_noise_estimator_noise_identification(oracle):
    ne = oracle.NoiseEstimator()
    ne.set_constants(10, np.float32(0.5) ** np.float32(1.0 / 20), np.float32(0.5) ** np.float32(1.0 / 50))
    sil = _silence_value()
    base = (sil / np.float32(160) * np.arange(160, dtype=np.float32) + sil).astype(np.float32)
    periodic = np.full(160, sil, dtype=np.float32)
    periodic[::20] = 1.0
    rng = np.random.default_rng(7)
    for _ in range(250):
        ne.update(base + rng.uniform(-0.1, 0.1, size=160).astype(np.float32))
    assert ne.compute_is_noise(base)
    assert not ne.compute_is_noise(periodic)


def test_noise_estimator_five_seconds_silence(oracle):
    ne = oracle.NoiseEstimator()
    sil = np.full(160, _silence_value(), dtype=np.float32)
    for i in range(250):
        mel = ne.receive_samples(np.zeros(320, dtype=np.int16))
        assert np.array_equal(mel, sil)
        assert oracle.log_spectral_distance(sil, ne.noise_estimate()) < 0.2, "frame %d" % i
    assert ne.is_noise     # after the first hop silence is classified as noise and only the bounds decay
