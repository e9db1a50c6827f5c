"""Parity cases shared by the CPU tier (emulated kernels) and the GPU tier (real kernels): every case drives
the C ABI and compares with the oracle on the same seeded inputs.  Bar: bit-exact packets, RVQ indices,
features, int16 PCM, log-mel spectra and noise-estimator state (the one libm call on each of those paths, log / exp,
is taken in double and rounded once on both sides).

The decoder's opt-in tensor-core mode (lyra_b200_set_decoder_mode, split-precision TF32) is the one floating-point
path that is compared with a tolerance: decoded int16 PCM within TENSOR_PCM_TOL_LSB of the oracle (the drop-in
target in BASELINE.json allows 1e-3 of full scale = 32.8 LSB); packets stay bit-exact in that mode too."""
import numpy as np

from conftest import MODEL_DIR


TENSOR_PCM_TOL_LSB = 4      # |PCM_tensor - PCM_oracle| <= 4 int16 LSB = 1.2e-4 of full scale


def synth_pcm(rng, n, kind="noise"):
    if kind == "noise":      # 0.25 full-scale uniform noise, the reference benchmark's input (lyra_benchmark_lib.cc:233-239)
        return rng.integers(-8192, 8192, size=(n, 320), dtype=np.int16)
    if kind == "loud":       # full-scale, exercises clipping in UnitToInt16 and int8 saturation
        return rng.integers(-32768, 32768, size=(n, 320), dtype=np.int16)
    if kind == "silence":
        return np.zeros((n, 320), dtype=np.int16)
    raise ValueError(kind)


def run_codec_parity(Context, api, O, *, max_streams, stream_ids, frames, bits, kind="noise", seed=0, loss_every=0,
                     wav=None, check=None, decoder_mode="exact"):
    """Encode+decode `frames` hops of the listed streams; compare `check` (default: all) against per-stream oracles.
    Returns the largest |PCM difference| seen (0 unless decoder_mode == "tensor")."""
    ctx = Context(max_streams, capi=api)
    ctx.set_decoder_mode(decoder_mode)
    pcm_tol = TENSOR_PCM_TOL_LSB if decoder_mode == "tensor" else 0
    worst = 0
    ids = np.asarray(stream_ids, dtype=np.int32)
    n = len(ids)
    check = list(range(n)) if check is None else check
    codecs = {k: O.Codec(MODEL_DIR) for k in check}
    rng = np.random.default_rng(seed)
    for f in range(frames):
        if wav is not None:
            pcm = np.stack([wav[(320 * (f + 7 * k)) % (len(wav) - 320):][:320] for k in range(n)])
        else:
            pcm = synth_pcm(rng, n, kind)
        packets = ctx.encode(pcm, bits, stream_ids=ids)
        received = None
        if loss_every:
            received = np.array([0 if (f + k) % loss_every == 0 else 1 for k in range(n)], dtype=np.uint8)
        out = ctx.decode(packets, bits, stream_ids=ids, received=received)
        for k in check:
            opkt, _, _ = codecs[k].encode(pcm[k], bits)
            lost = received is not None and received[k] == 0
            opcm, _, _ = codecs[k].decode(None if lost else opkt, bits)
            assert bytes(packets[k]) == opkt, "packet mismatch frame %d stream %d" % (f, ids[k])
            diff = int(np.abs(out[k].astype(int) - opcm.astype(int)).max())
            worst = max(worst, diff)
            assert diff <= pcm_tol, "PCM mismatch frame %d stream %d (max |d| %d, allowed %d)" % (f, ids[k], diff, pcm_tol)
    ctx.close()
    return worst


def run_priority_switch(Context, api, O, *, n=3, frames=4, bits=64, split=2):
    """lyra_b200_set_priority between hops (streams re-created, sub-batches in use): the streaming state is untouched, results stay
    the oracle's."""
    ctx = Context(max(16, n), capi=api)
    ctx.set_split(split)
    codecs = [O.Codec(MODEL_DIR) for _ in range(n)]
    rng = np.random.default_rng(21)
    for f in range(frames):
        ctx.set_priority([-1, 0, -2, 0][f % 4])
        pcm = synth_pcm(rng, n)
        packets = ctx.encode(pcm, bits)
        out = ctx.decode(packets, bits)
        for k in range(n):
            opkt, _, _ = codecs[k].encode(pcm[k], bits)
            opcm, _, _ = codecs[k].decode(opkt, bits)
            assert bytes(packets[k]) == opkt and np.array_equal(out[k], opcm), "mismatch after a priority switch (hop %d stream %d)" % (f, k)
    ctx.close()


def run_plugin_surface_parity(Context, api, O, *, n=5, frames=3, seed=1):
    """extract_features / quantize / dequantize / generate one by one against the oracle's pieces."""
    import os
    ctx = Context(n, capi=api)
    rng = np.random.default_rng(seed)
    codecs = [O.Codec(MODEL_DIR) for _ in range(n)]
    rvq = O.Rvq(os.path.join(MODEL_DIR, "quantizer.tflite"))
    for f in range(frames):
        pcm = synth_pcm(rng, n, "noise")
        feats = ctx.extract_features(pcm)
        for bits in (64, 120, 184):
            packets, idx = ctx.quantize(feats, bits, want_indices=True)
            lossy = ctx.dequantize(packets, bits)
            for k in range(n):
                want_idx = rvq.encode(feats[k], bits // 4)
                assert np.array_equal(idx[k], want_idx)
                want_bits = rvq.quantize(feats[k], bits)
                assert bytes(packets[k]) == O.packet_pack(want_bits, 0, bits)
                assert np.array_equal(lossy[k], rvq.decode_to_lossy_features(want_bits))
        lossy = ctx.dequantize(ctx.quantize(feats, 120), 120)
        out = ctx.generate(lossy)
        for k in range(n):
            opkt, ofeat, _ = codecs[k].encode(pcm[k], 120)
            assert np.array_equal(feats[k], ofeat)
            opcm, olossy, _ = codecs[k].decode(opkt, 120)
            assert np.array_equal(lossy[k], olossy)
            assert np.array_equal(out[k], opcm)
    ctx.close()


def run_reset_and_isolation(Context, api, O, *, seed=2):
    """Streams are independent; reset(ids) restores exactly the initial state of those streams only."""
    ctx = Context(40, capi=api)
    rng = np.random.default_rng(seed)
    a = synth_pcm(rng, 3, "noise")
    ids = np.array([3, 19, 33], dtype=np.int32)
    first = [ctx.encode(a, 64, stream_ids=ids) for _ in range(2)]
    # other streams running in between must not disturb 3/19/33
    ctx.encode(synth_pcm(rng, 4, "noise"), 64, stream_ids=np.array([2, 4, 18, 32], dtype=np.int32))
    third = ctx.encode(a, 64, stream_ids=ids)
    ref = O.Codec(MODEL_DIR)
    want = [ref.encode(a[0], 64)[0] for _ in range(3)]
    assert [bytes(first[0][0]), bytes(first[1][0]), bytes(third[0])] == want
    ctx.reset(np.array([19], dtype=np.int32))
    again = ctx.encode(a, 64, stream_ids=ids)
    ref19 = O.Codec(MODEL_DIR)
    assert bytes(again[1]) == ref19.encode(a[1], 64)[0]          # stream 19 restarted from zero state
    assert bytes(again[0]) == ref.encode(a[0], 64)[0]            # stream 3 carried on
    ctx.close()


def run_error_paths(Context, api, LyraB200Error):
    ctx = Context(8, capi=api)
    pcm = np.zeros((2, 320), dtype=np.int16)
    feats = np.zeros((1, 64), dtype=np.float32)

    def fails(fn):
        try:
            fn()
        except LyraB200Error as e:
            assert e.code == -1
            return True
        return False
    assert fails(lambda: ctx.quantize(feats, 185))                                  # too many bits
    assert fails(lambda: ctx.quantize(feats, 62))                                   # not divisible by 4
    assert fails(lambda: ctx.encode(pcm, 64, stream_ids=np.array([1, 1], dtype=np.int32)))   # duplicate id
    assert fails(lambda: ctx.encode(pcm, 64, stream_ids=np.array([1, 8], dtype=np.int32)))   # id out of range
    assert fails(lambda: ctx.encode(np.zeros((9, 320), np.int16), 64))              # more rows than streams
    assert fails(lambda: ctx.logmel(pcm, num_mel_bins=80))
    ctx.close()


def run_logmel_parity(Context, api, O, wav, *, n=4, frames=4, tol=0.0):
    ctx = Context(n, capi=api)
    for nmel, bank in ((160, 0), (64, 1)):
        refs = [O.LogMel(16000, 320, 640, nmel) for _ in range(n)]
        for f in range(frames):
            pcm = np.stack([wav[320 * (f + 11 * k):][:320] for k in range(n)])
            out = ctx.logmel(pcm, num_mel_bins=nmel, bank=bank)
            for k in range(n):
                want = refs[k].extract(pcm[k])
                assert np.abs(out[k] - want).max() <= tol, (nmel, f, k, np.abs(out[k] - want).max())
    ctx.close()


def run_noise_estimator_parity(Context, api, O, wav, *, n=3, frames=30, seed=4):
    """lyra_b200_noise_update against one oracle NoiseEstimator per stream: speech with silent stretches (so both the
    update and the decay branch run), some hops withheld (update_mask 0, as after a lost packet), sparse stream ids.
    Bar: is_noise identical; noise_estimate bit-identical (the kernel evaluates the C++ expressions operation by operation)."""
    ctx = Context(4 * n, capi=api)
    ids = np.arange(n, dtype=np.int32) * 3 + 1
    est = [O.NoiseEstimator() for _ in range(n)]
    rng = np.random.default_rng(seed)
    seen_noise, seen_speech = False, False
    for f in range(frames):
        hop = np.stack([wav[(320 * (f + 11 * k)) % (len(wav) - 320):][:320] for k in range(n)]).copy()
        quiet = rng.random(n) < 0.4                   # silence / faint noise: exercises the is-noise branch
        for k in range(n):
            if quiet[k]:
                hop[k] = rng.integers(-2, 3, size=320, dtype=np.int16) if f % 2 else 0
        mask = (rng.random(n) < 0.85).astype(np.uint8)
        flags, got = ctx.noise_update(hop, stream_ids=ids, update_mask=mask)
        for k in range(n):
            if mask[k]:
                est[k].receive_samples(hop[k])
            assert bool(flags[k]) == est[k].is_noise, "is_noise mismatch frame %d stream %d" % (f, k)
            want = est[k].noise_estimate()
            assert np.array_equal(got[k], want), "noise estimate mismatch frame %d stream %d (max |d| %g)" % (
                f, k, np.abs(got[k] - want).max())
            seen_noise |= bool(flags[k]) and f > 0
            seen_speech |= not bool(flags[k])
    assert seen_noise and seen_speech, "the case must exercise both branches"
    # reset restores the freshly constructed estimator for the listed streams only
    ctx.reset(stream_ids=ids[:1])
    flags, got = ctx.noise_update(np.zeros((n, 320), dtype=np.int16), stream_ids=ids, update_mask=np.zeros(n, dtype=np.uint8))
    assert flags[0] and not got[0].any()
    assert np.array_equal(got[1], est[1].noise_estimate())
    ctx.close()


def run_decode_track_noise_parity(Context, api, O, wav, *, n=3, frames=12, max_streams=None, stream_ids=None, loss_every=4,
                                  check=None):
    """lyra_b200_decode_track_noise = decode + NoiseEstimator::ReceiveSamples for the received streams
    (LyraDecoder::DecodeSamplesInternal, lyra/lyra_decoder.cc:306-311): PCM, is_noise and the estimate bit-exact."""
    ids = np.arange(n, dtype=np.int32) if stream_ids is None else np.asarray(stream_ids, dtype=np.int32)
    n = len(ids)
    ctx = Context(max_streams or int(ids.max()) + 1, capi=api)
    dense = stream_ids is None and (max_streams is None or max_streams == n)
    check = list(range(n)) if check is None else check
    codecs = {k: O.Codec(MODEL_DIR) for k in check}
    est = {k: O.NoiseEstimator() for k in check}
    for f in range(frames):
        pcm = np.stack([wav[(320 * (f + 5 * k)) % (len(wav) - 320):][:320] for k in range(n)]).copy()
        if f % 3 == 2:
            pcm[:] = 0                                  # silent hops: the decoder output becomes noise-like
        packets = ctx.encode(pcm, 64, stream_ids=None if dense else ids)
        received = np.array([0 if (f + k) % loss_every == 0 else 1 for k in range(n)], dtype=np.uint8)
        out, flags = ctx.decode_track_noise(packets, 64, stream_ids=None if dense else ids, received=received)
        for k in check:
            opkt, _, _ = codecs[k].encode(pcm[k], 64)
            opcm, _, _ = codecs[k].decode(opkt if received[k] else None, 64)
            assert np.array_equal(out[k], opcm), "PCM mismatch frame %d stream %d" % (f, k)
            if received[k]:
                est[k].receive_samples(opcm)
            assert bool(flags[k]) == est[k].is_noise, "is_noise mismatch frame %d stream %d" % (f, k)
    _, got = ctx.noise_update(np.zeros((n, 320), dtype=np.int16), stream_ids=None if dense else ids, update_mask=np.zeros(n, dtype=np.uint8))
    for k in check:
        assert np.array_equal(got[k], est[k].noise_estimate()), "noise estimate mismatch stream %d" % k
    ctx.close()


def run_role_contexts(Context, api, O, LyraB200Error, *, frames=4, seed=9):
    """lyra_b200_create_ex: an encoder-only and a decoder-only context together reproduce the oracle; calls of the missing
    role are refused with EINVAL (the mirror of LyraEncoder / LyraDecoder being separate objects)."""
    n = 3
    enc = Context(8, capi=api, roles="encoder")
    dec = Context(8, capi=api, roles="decoder")
    codecs = [O.Codec(MODEL_DIR) for _ in range(n)]
    rng = np.random.default_rng(seed)
    for f in range(frames):
        pcm = synth_pcm(rng, n)
        pk = enc.encode(pcm, 120)
        out = dec.decode(pk, 120)
        for k in range(n):
            opkt, _, _ = codecs[k].encode(pcm[k], 120)
            opcm, _, _ = codecs[k].decode(opkt, 120)
            assert bytes(pk[k]) == opkt and np.array_equal(out[k], opcm), (f, k)
    for bad in (lambda: enc.decode(pk, 120), lambda: dec.encode(pcm, 120), lambda: enc.decode_plc(pk, 120),
                lambda: dec.encode_dtx(pcm, 120), lambda: dec.extract_features(pcm)):
        try:
            bad()
            raise AssertionError("a call of the missing role must fail")
        except LyraB200Error as e:
            assert e.code == -1
    assert enc.quantize(np.zeros((1, 64), dtype=np.float32), 64).shape == (1, 8)      # stateless calls work in any context
    assert enc.noise_update(pcm)[1].shape == (n, 160)                                  # ... and so do the self-contained estimators
    enc.reset()
    dec.reset()
    enc.close()
    dec.close()


# ---- packet-loss concealment / comfort noise / DTX ----

def run_cng_parity(Context, api, O, *, stream_ids=(0, 5), hops=4, seed=9, cng_seed=77):
    """lyra_b200_cng_generate vs the oracle's ComfortNoiseGenerator on the same features and the same seeded phases:
    bit-identical int16 hops (overlap-add state included), several hops in a row."""
    ids = np.asarray(stream_ids, dtype=np.int32)
    ctx = Context(int(ids.max()) + 1, capi=api)
    ctx.set_cng_seed(cng_seed)
    gens = [O.ComfortNoiseGenerator(seed=cng_seed + int(i)) for i in ids]
    rng = np.random.default_rng(seed)
    for h in range(hops):
        feats = rng.uniform(0.62, 1.2, size=(len(ids), 160)).astype(np.float32)       # log-mel values between the floor and loud noise
        out = ctx.cng_generate(feats, stream_ids=ids)
        for k, g in enumerate(gens):
            want = g.condition(feats[k])
            assert np.array_equal(out[k], want), "comfort noise mismatch hop %d stream %d (max |d| %d)" % (
                h, ids[k], int(np.abs(out[k].astype(int) - want.astype(int)).max()))
    ctx.close()


def run_plc_parity(Context, api, O, *, max_streams=16, stream_ids=(1, 6, 9), frames=26, bits=64, wav=None, seed=4, cng_seed=5,
                   outages=((3, 12), (5, 3), (0, 0)), decoder_mode="exact"):
    """lyra_b200_decode_plc tick by tick vs one oracle LyraDecoder per stream: stream k loses `outages[k] = (first, count)` hops.
    PCM (model audio, comfort noise and their cross-fades), the control state and is_comfort_noise must all agree bit for bit
    (PCM within the stated tolerance in the tensor decoder mode)."""
    ids = np.asarray(stream_ids, dtype=np.int32)
    n = len(ids)
    ctx = Context(max_streams, capi=api)
    ctx.set_decoder_mode(decoder_mode)
    ctx.set_cng_seed(cng_seed)
    tol = TENSOR_PCM_TOL_LSB if decoder_mode == "tensor" else 0
    encs = [O.Encoder(MODEL_DIR) for _ in range(n)]
    decs = [O.Decoder(MODEL_DIR, cng_seed=cng_seed + int(i)) for i in ids]
    rng = np.random.default_rng(seed)
    seen_cn = False
    for f in range(frames):
        if wav is not None:
            pcm = np.stack([wav[(320 * (f + 11 * k)) % (len(wav) - 320):][:320] for k in range(n)])
        else:
            pcm = synth_pcm(rng, n, "noise")
        pk = np.stack([np.frombuffer(encs[k].encode(pcm[k], bits), dtype=np.uint8) for k in range(n)])
        rec = np.array([0 if outages[k][0] <= f < outages[k][0] + outages[k][1] else 1 for k in range(n)], dtype=np.uint8)
        out, cn = ctx.decode_plc(pk, bits, stream_ids=ids, received=rec)
        st = ctx.plc_state(stream_ids=ids)
        for k in range(n):
            if rec[k]:
                assert decs[k].set_encoded_packet(bytes(pk[k]))
            want = decs[k].decode_samples(320)
            d = int(np.abs(out[k].astype(int) - want.astype(int)).max())
            assert d <= tol, "PLC PCM mismatch frame %d stream %d: max |d| %d (state %s)" % (f, ids[k], d, decs[k].state)
            assert tuple(int(x) for x in st[k]) == decs[k].state, (f, k, st[k], decs[k].state)
            assert bool(cn[k]) == decs[k].is_comfort_noise()
            seen_cn |= bool(cn[k])
    assert seen_cn, "the case never reached comfort noise"
    ctx.close()


def run_plc_state_peer(Context, api, O):
    """The reference's test peer (lyra_decoder_test.cc:56-90): forced states produce the same next hop as the oracle, and
    misaligned states are refused."""
    ctx = Context(4, capi=api)
    ctx.set_cng_seed(3)
    pk = np.zeros((1, 8), dtype=np.uint8)
    for state, rec in [((1280, 0, 1), 0), ((1280, 640, 1), 0), ((0, 640, -1), 1), ((1280, 320, 1), 1), ((640, 0, -1), 0)]:
        ctx.reset()
        ctx.set_plc_state([state], stream_ids=[2])
        dec = O.Decoder(MODEL_DIR, cng_seed=3 + 2)
        dec.state = state
        if rec:
            assert dec.set_encoded_packet(bytes(pk[0]))
        want = dec.decode_samples(320)
        out, cn = ctx.decode_plc(pk, 64, stream_ids=[2], received=[rec])
        assert np.array_equal(out[0], want), state
        assert tuple(int(x) for x in ctx.plc_state(stream_ids=[2])[0]) == dec.state
    try:
        ctx.set_plc_state([(100, 0, 1)], stream_ids=[0])
    except Exception:
        pass
    else:
        raise AssertionError("a misaligned control state must be refused")
    ctx.close()


def run_dtx_parity(Context, api, O, *, wav, frames=24, bits=64, stream_ids=(0, 3, 4)):
    """lyra_b200_encode_dtx vs the oracle's LyraEncoder(enable_dtx): stream 0 speech, stream 1 digital silence, stream 2
    silence then speech.  Packet sizes (0 = DTX) and bytes agree; encoder state only advances on encoded hops."""
    ids = np.asarray(stream_ids, dtype=np.int32)
    n = len(ids)
    ctx = Context(int(ids.max()) + 1, capi=api)
    encs = [O.Encoder(MODEL_DIR, enable_dtx=True) for _ in range(n)]
    sizes_seen = set()
    for f in range(frames):
        speech = wav[320 * (f + 20):320 * (f + 21)]
        pcm = np.stack([speech, np.zeros(320, np.int16), speech if f >= frames // 2 else np.zeros(320, np.int16)])
        pk, sizes = ctx.encode_dtx(pcm, bits, stream_ids=ids)
        for k in range(n):
            want = encs[k].encode(pcm[k], bits)
            assert sizes[k] == len(want), (f, k, sizes[k], len(want))
            assert bytes(pk[k][:sizes[k]]) == want
            if sizes[k] == 0:
                assert not pk[k].any()
            sizes_seen.add(int(sizes[k]))
    assert sizes_seen == {0, (bits + 7) // 8}
    ctx.close()


def run_resampler_parity(Context, api, O, *, seed=12):
    """lyra_b200_resample vs the oracle's Resampler: every supported pair, both directions, ragged chunk sizes (phase and delay line
    carry over between calls), several streams with different histories, a rate switch; int16 output bit for bit."""
    ctx = Context(8, capi=api)
    rng = np.random.default_rng(seed)
    ids = np.array([0, 3, 5], dtype=np.int32)
    for rate in (8000, 32000, 48000):
        for to_internal in (True, False):
            a, b = (rate, 16000) if to_internal else (16000, rate)
            refs = [O.Resampler(a, b) for _ in ids]
            ctx.reset()
            for chunk in (a // 50, 1, 7, a // 50 - 3, 2, a // 50):
                x = rng.integers(-30000, 30000, size=(len(ids), chunk)).astype(np.int16)
                got = ctx.resample(x, rate, to_internal, stream_ids=ids)
                for k in range(len(ids)):
                    want = refs[k].resample(x[k])
                    assert np.array_equal(got[k], want), (rate, to_internal, chunk, k, len(got[k]), len(want))
    # a stream that switches rate restarts from the fully primed state
    x = rng.integers(-30000, 30000, size=(1, 160)).astype(np.int16)
    ctx.resample(x, 8000, True, stream_ids=[2])
    y = rng.integers(-30000, 30000, size=(1, 960)).astype(np.int16)
    got = ctx.resample(y, 48000, True, stream_ids=[2])
    assert np.array_equal(got[0], O.Resampler(48000, 16000).resample(y[0]))
    ctx.close()


def run_integration_other_rates(Context, api, O, *, rate, wav, bits=64, hops=60):
    """lyra_integration_test.cc:60-149 at an external rate of 8 / 32 / 48 kHz through the batched C ABI: resample to 16 kHz, encode,
    decode, resample back; the log-mel spectra (64 bins at the external rate, the oracle's extractor) of input and output stay
    within LSD 2.0 on every hop once the filters are primed."""
    hop = rate // 50
    ctx = Context(2, capi=api)
    ie, oe = O.LogMel(rate, hop, 2 * hop, 64), O.LogMel(rate, hop, 2 * hop, 64)
    worst = 0.0
    outs = []
    for f in range(hops):
        x = wav[f * hop:(f + 1) * hop]
        internal = ctx.resample(x, rate, True, stream_ids=[1])[0]
        assert len(internal) == 320
        pk = ctx.encode(internal, bits, stream_ids=[1])
        dec = ctx.decode(pk, bits, stream_ids=[1])[0]
        y = ctx.resample(dec, rate, False, stream_ids=[1])[0]
        assert len(y) == hop
        outs.append(y)
    # the codec (one hop) and the two resamplers (17 + 17 * 16000 / rate ... samples) delay the output: compare hop f of the input
    # with hop f of the output like the reference does (its criterion tolerates the misalignment), skipping the priming hops
    for f in range(hops):
        fi = ie.extract(wav[f * hop:(f + 1) * hop])
        fo = oe.extract(outs[f])
        if f >= 3:
            worst = max(worst, O.log_spectral_distance(fi, fo))
    ctx.close()
    return worst
