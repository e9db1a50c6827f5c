"""CPU tier: the product kernels' source, compiled for the test-only CUDA block simulator (tests/cuda_emu),
checked against the oracle.  Small sizes (the simulator runs every CUDA thread as a fiber)."""
import numpy as np

import parity_cases as pc
from lyra_b200 import _capi


def test_emu_codec_parity_sparse_ids_and_loss(emu_api, oracle, sample1):
    # 3 streams spread over 2 tiles, speech input, a lost packet every 5th frame, 20 frames (ring wrap at 18)
    pc.run_codec_parity(_capi.Context, emu_api, oracle, max_streams=20, stream_ids=[0, 5, 17], frames=20, bits=64,
                        wav=sample1, loss_every=5)


def test_emu_codec_parity_all_bitrates(emu_api, oracle):
    for bits in (120, 184):
        pc.run_codec_parity(_capi.Context, emu_api, oracle, max_streams=16, stream_ids=[1, 2], frames=3, bits=bits, seed=bits)


def test_emu_loud_and_silent_input(emu_api, oracle):
    pc.run_codec_parity(_capi.Context, emu_api, oracle, max_streams=16, stream_ids=[4], frames=3, bits=64, kind="loud")
    pc.run_codec_parity(_capi.Context, emu_api, oracle, max_streams=16, stream_ids=[4], frames=2, bits=64, kind="silence")


def test_emu_tensor_decoder_mode(emu_api, oracle, sample1):
    # split-precision TF32 decoder: packets bit-exact, PCM within the stated tolerance, over a ring wrap, with loss
    worst = pc.run_codec_parity(_capi.Context, emu_api, oracle, max_streams=16, stream_ids=[0, 9], frames=20, bits=64,
                                wav=sample1, loss_every=7, decoder_mode="tensor")
    assert worst <= pc.TENSOR_PCM_TOL_LSB
    pc.run_codec_parity(_capi.Context, emu_api, oracle, max_streams=8, stream_ids=[3], frames=3, bits=184, kind="loud",
                        decoder_mode="tensor")


def test_emu_plugin_surface(emu_api, oracle):
    pc.run_plugin_surface_parity(_capi.Context, emu_api, oracle, n=3, frames=2)


def test_emu_reset_and_isolation(emu_api, oracle):
    pc.run_reset_and_isolation(_capi.Context, emu_api, oracle)


def test_emu_error_paths(emu_api):
    pc.run_error_paths(_capi.Context, emu_api, _capi.LyraB200Error)


def test_emu_logmel(emu_api, oracle, sample1):
    pc.run_logmel_parity(_capi.Context, emu_api, oracle, sample1, n=2, frames=3)


def test_emu_noise_estimator(emu_api, oracle, sample1):
    pc.run_noise_estimator_parity(_capi.Context, emu_api, oracle, sample1, n=2, frames=14)


def test_emu_decode_track_noise(emu_api, oracle, sample1):
    pc.run_decode_track_noise_parity(_capi.Context, emu_api, oracle, sample1, stream_ids=[1, 10], max_streams=16, frames=5, loss_every=3)
