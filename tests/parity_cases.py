"""Parity cases shared by the CPU tier (emulated kernels) and the GPU tier (real kernels): every case drives
the C ABI and compares with the oracle on the same seeded inputs.  Bar: bit-exact packets, RVQ indices,
features and int16 PCM; log-mel within 2e-6 absolute (float log).

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


def run_logmel_parity(Context, api, O, wav, *, n=4, frames=4, tol=2e-6):
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
