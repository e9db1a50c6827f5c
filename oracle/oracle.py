"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes binding of oracle/_build/liblyra_oracle.so (plain-C restatement of the reference hot path,
see oracle/lyra_oracle.h).  Imported only by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / ``--impl reference`` legs — never by lyra_b200/.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liblyra_oracle.so")
_lib = None


def build(force=False):
    """Compile the C restatement (gcc, a few seconds). Building the checker is not using it."""
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".c", ".h")) or f == "Makefile"]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        vp, ci, cf = C.c_void_p, C.c_int, C.c_float
        sigs = {
            "lo_net_create": (vp, [C.c_char_p]), "lo_net_free": (None, [vp]), "lo_net_reset": (ci, [vp]),
            "lo_net_invoke": (ci, [vp, vp, ci, vp, ci]), "lo_net_num_tensors": (ci, [vp]),
            "lo_net_tensor_info": (ci, [vp, ci, vp, vp, vp, vp]), "lo_net_read_tensor": (ci, [vp, ci, vp, ci]),
            "lo_net_num_vars": (ci, [vp]), "lo_net_var_name": (C.c_char_p, [vp, ci]),
            "lo_net_var_count": (ci, [vp, ci]), "lo_net_read_var": (ci, [vp, ci, vp, ci]),
            "lo_quantize_multiplier": (None, [C.c_double, vp, vp]), "lo_mbqm": (C.c_int32, [C.c_int32, C.c_int32, ci]),
            "lo_rvq_create": (vp, [C.c_char_p]), "lo_rvq_free": (None, [vp]), "lo_rvq_num_stages": (ci, [vp]),
            "lo_rvq_bits_per_stage": (ci, [vp]), "lo_rvq_codebook": (vp, [vp, ci]),
            "lo_rvq_encode": (ci, [vp, vp, ci, vp]), "lo_rvq_decode": (ci, [vp, vp, vp]),
            "lo_rvq_quantize_bits": (ci, [vp, vp, ci, vp]), "lo_rvq_decode_bits": (ci, [vp, C.c_char_p, ci, vp]),
            "lo_packet_size": (ci, [ci, ci]), "lo_packet_pack": (ci, [C.c_char_p, ci, ci, vp]),
            "lo_packet_unpack": (ci, [vp, ci, ci, ci, vp]),
            "lo_int16_to_unit": (cf, [C.c_int16]), "lo_unit_to_int16": (C.c_int16, [cf]),
            "lo_log_spectral_distance": (cf, [vp, vp, ci]),
            "lo_logmel_create": (vp, [ci, ci, ci, ci]), "lo_logmel_free": (None, [vp]),
            "lo_logmel_extract": (ci, [vp, vp, ci, vp]),
            "lo_noise_create": (vp, [ci, ci, ci, ci]), "lo_noise_free": (None, [vp]),
            "lo_noise_set_constants": (None, [vp, ci, cf, cf]),
            "lo_noise_receive_samples": (ci, [vp, vp, vp]), "lo_noise_update": (None, [vp, vp]),
            "lo_noise_compute_is_noise": (ci, [vp, vp]), "lo_noise_is_noise": (ci, [vp]),
            "lo_noise_estimate": (None, [vp, vp]), "lo_noise_bound": (None, [vp, vp]),
            "lo_noise_receive_partial": (ci, [vp, vp, ci]),
            "lo_gen_add_features": (ci, [vp, vp, ci]), "lo_gen_num_samples_available": (ci, [vp]),
            "lo_gen_generate_samples": (ci, [vp, ci, vp]),
            "lo_cng_create": (vp, [ci, ci, ci, ci, C.c_uint64]), "lo_cng_free": (None, [vp]), "lo_cng_gen": (vp, [vp]),
            "lo_cng_condition": (ci, [vp, vp, vp, vp]), "lo_cng_phase_index": (C.c_uint32, [C.c_uint64, C.c_uint64, ci]),
            "lo_cng_synthesis_gain": (C.c_double, [vp]),
            "lo_decoder_create": (vp, [C.c_char_p, C.c_uint64]), "lo_decoder_create_fake": (vp, [C.c_int16, C.c_int16]),
            "lo_decoder_free": (None, [vp]), "lo_decoder_set_encoded_packet": (ci, [vp, vp, ci]),
            "lo_decoder_decode_samples": (ci, [vp, ci, vp]), "lo_decoder_is_comfort_noise": (ci, [vp]),
            "lo_decoder_get_state": (None, [vp, vp]), "lo_decoder_set_state": (None, [vp, vp]),
            "lo_decoder_counters": (None, [vp, vp]), "lo_decoder_noise": (vp, [vp]), "lo_decoder_cng": (vp, [vp]),
            "lo_encoder_create": (vp, [C.c_char_p, ci]), "lo_encoder_free": (None, [vp]),
            "lo_encoder_encode": (ci, [vp, vp, ci, ci, vp]),
            "lo_resampler_create": (vp, [ci, ci]), "lo_resampler_free": (None, [vp]), "lo_resampler_reset": (None, [vp]),
            "lo_resampler_samples_until_steady_state": (ci, [vp]), "lo_resampler_resample": (ci, [vp, vp, ci, vp, ci]),
            "lo_resampler_design": (ci, [ci, ci, vp, vp, vp, ci]),
            "lo_buffered_resampler_create": (vp, [ci, ci]), "lo_buffered_resampler_free": (None, [vp]),
            "lo_buffered_resampler_leftover": (ci, [vp]), "lo_buffered_resampler_internal_samples": (ci, [vp, ci]),
            "lo_buffered_resampler_filter_and_buffer": (ci, [vp, vp, vp, ci, vp]),
            "lo_codec_create": (vp, [C.c_char_p]), "lo_codec_free": (None, [vp]), "lo_codec_reset": (ci, [vp]),
            "lo_codec_encode": (ci, [vp, vp, ci, vp, vp, vp]), "lo_codec_decode": (ci, [vp, vp, ci, vp, vp, vp]),
            "lo_codec_encoder_net": (vp, [vp]), "lo_codec_decoder_net": (vp, [vp]),
            "lo_cpu_bench": (C.c_double, [C.c_char_p, ci, ci, ci, ci, C.c_uint32, vp, vp]),
        }
        for name, (res, args) in sigs.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


_NP = {0: np.float32, 2: np.int32, 9: np.int8, 4: np.int64, 6: np.bool_, 3: np.uint8}


class Net:
    """One stream of soundstream_encoder.tflite or lyragan.tflite."""

    def __init__(self, path=None, handle=None, owner=None):
        self._own = handle is None
        self._owner = owner
        self.h = lib().lo_net_create(path.encode()) if handle is None else handle
        if not self.h:
            raise RuntimeError("oracle: cannot load %s" % path)

    def __del__(self):
        if getattr(self, "_own", False) and self.h:
            lib().lo_net_free(self.h)
            self.h = None

    def reset(self):
        assert lib().lo_net_reset(self.h) == 0

    def invoke(self, x, n_out):
        x = np.ascontiguousarray(x, dtype=np.float32)
        y = np.empty(n_out, dtype=np.float32)
        rc = lib().lo_net_invoke(self.h, _p(x), x.size, _p(y), n_out)
        if rc != 0:
            raise RuntimeError("oracle invoke failed rc=%d" % rc)
        return y

    def tensor(self, idx):
        t, c, s, z = C.c_int(), C.c_int(), C.c_float(), C.c_int()
        assert lib().lo_net_tensor_info(self.h, idx, C.byref(t), C.byref(c), C.byref(s), C.byref(z)) == 0
        a = np.empty(c.value, dtype=_NP[t.value])
        assert lib().lo_net_read_tensor(self.h, idx, _p(a), a.nbytes) == a.nbytes
        return a

    def tensor_quant(self, idx):
        t, c, s, z = C.c_int(), C.c_int(), C.c_float(), C.c_int()
        assert lib().lo_net_tensor_info(self.h, idx, C.byref(t), C.byref(c), C.byref(s), C.byref(z)) == 0
        return s.value, z.value

    def variables(self):
        out = {}
        for v in range(lib().lo_net_num_vars(self.h)):
            n = lib().lo_net_var_count(self.h, v)
            a = np.empty(n, dtype=np.float32)
            assert lib().lo_net_read_var(self.h, v, _p(a), n) == n
            out[lib().lo_net_var_name(self.h, v).decode()] = a
        return out


class Rvq:
    def __init__(self, path):
        self.h = lib().lo_rvq_create(path.encode())
        if not self.h:
            raise RuntimeError("oracle: cannot load %s" % path)
        self.num_stages = lib().lo_rvq_num_stages(self.h)
        self.bits_per_stage = lib().lo_rvq_bits_per_stage(self.h)

    def __del__(self):
        if getattr(self, "h", None):
            lib().lo_rvq_free(self.h)
            self.h = None

    def codebooks(self):
        out = np.empty((self.num_stages, 16, 64), dtype=np.float32)
        for s in range(self.num_stages):
            ptr = lib().lo_rvq_codebook(self.h, s)
            out[s] = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_float)), shape=(16, 64))
        return out

    def encode(self, features, num_quantizers):
        f = np.ascontiguousarray(features, dtype=np.float32)
        idx = np.empty(self.num_stages, dtype=np.int32)
        assert lib().lo_rvq_encode(self.h, _p(f), num_quantizers, _p(idx)) == 0
        return idx

    def decode(self, indices):
        i = np.ascontiguousarray(indices, dtype=np.int32)
        assert i.size == self.num_stages
        out = np.empty(64, dtype=np.float32)
        assert lib().lo_rvq_decode(self.h, _p(i), _p(out)) == 0
        return out

    def quantize(self, features, num_bits):
        """ResidualVectorQuantizer::Quantize -> '0'/'1' string or None."""
        f = np.ascontiguousarray(features, dtype=np.float32)
        buf = C.create_string_buffer(256)
        rc = lib().lo_rvq_quantize_bits(self.h, _p(f), num_bits, buf)
        return buf.value.decode() if rc == 0 else None

    def decode_to_lossy_features(self, bits):
        out = np.empty(64, dtype=np.float32)
        rc = lib().lo_rvq_decode_bits(self.h, bits.encode(), len(bits), _p(out))
        return out if rc == 0 else None


def packet_size(num_header_bits, num_quantized_bits):
    return lib().lo_packet_size(num_header_bits, num_quantized_bits)


def packet_pack(bits, num_header_bits=0, num_quantized_bits=None):
    nq = len(bits) if num_quantized_bits is None else num_quantized_bits
    buf = np.zeros(32, dtype=np.uint8)
    n = lib().lo_packet_pack(bits.encode(), num_header_bits, nq, _p(buf))
    return None if n < 0 else bytes(buf[:n])


def packet_unpack(data, num_header_bits, num_quantized_bits):
    a = np.frombuffer(bytes(data), dtype=np.uint8).copy()
    out = C.create_string_buffer(256)
    n = lib().lo_packet_unpack(_p(a), a.size, num_header_bits, num_quantized_bits, out)
    return None if n < 0 else out.value.decode()


def int16_to_unit(v):
    return lib().lo_int16_to_unit(int(v))


def unit_to_int16(v):
    return lib().lo_unit_to_int16(float(v))


def log_spectral_distance(a, b):
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = np.ascontiguousarray(b, dtype=np.float32)
    return lib().lo_log_spectral_distance(_p(a), _p(b), a.size)


class LogMel:
    def __init__(self, sample_rate_hz, hop, window, num_mel_bins):
        self.h = lib().lo_logmel_create(sample_rate_hz, hop, window, num_mel_bins)
        self.nmel = num_mel_bins
        if not self.h:
            raise ValueError("oracle logmel: bad parameters")

    def __del__(self):
        if getattr(self, "h", None):
            lib().lo_logmel_free(self.h)
            self.h = None

    def extract(self, audio):
        a = np.ascontiguousarray(audio, dtype=np.int16)
        out = np.empty(self.nmel, dtype=np.float32)
        rc = lib().lo_logmel_extract(self.h, _p(a), a.size, _p(out))
        return out if rc == 0 else None


class NoiseEstimator:
    """NoiseEstimator (lyra/noise_estimator.h): one decoded stream's minimum-statistics noise tracker."""

    def __init__(self, sample_rate_hz=16000, hop=320, window=640, num_features=160):
        self.h = lib().lo_noise_create(sample_rate_hz, hop, window, num_features)
        self.nf = num_features
        if not self.h:
            raise ValueError("oracle noise estimator: bad parameters")

    def __del__(self):
        if getattr(self, "h", None):
            lib().lo_noise_free(self.h)
            self.h = None

    def set_constants(self, hops_per_update, max_smoothing, bound_decay):
        lib().lo_noise_set_constants(self.h, hops_per_update, max_smoothing, bound_decay)

    def receive_samples(self, hop):
        a = np.ascontiguousarray(hop, dtype=np.int16)
        mel = np.empty(self.nf, dtype=np.float32)
        if lib().lo_noise_receive_samples(self.h, _p(a), _p(mel)):
            raise RuntimeError("oracle noise estimator failed")
        return mel

    def update(self, current_power_db):
        lib().lo_noise_update(self.h, _p(np.ascontiguousarray(current_power_db, dtype=np.float32)))

    def compute_is_noise(self, current_power_db):
        return bool(lib().lo_noise_compute_is_noise(self.h, _p(np.ascontiguousarray(current_power_db, dtype=np.float32))))

    @property
    def is_noise(self):
        return bool(lib().lo_noise_is_noise(self.h))

    def noise_estimate(self):
        out = np.empty(self.nf, dtype=np.float32)
        lib().lo_noise_estimate(self.h, _p(out))
        return out

    def noise_bound(self):
        out = np.empty(self.nf, dtype=np.float32)
        lib().lo_noise_bound(self.h, _p(out))
        return out


class ComfortNoiseGenerator:
    """ComfortNoiseGenerator (lyra/comfort_noise_generator.{h,cc}) with the GenerativeModel FIFO; seeded phases."""

    def __init__(self, sample_rate_hz=16000, hop=320, window=640, num_mel_bins=160, seed=0):
        self.h = lib().lo_cng_create(sample_rate_hz, hop, window, num_mel_bins, seed)
        if not self.h:
            raise ValueError("lo_cng_create failed")
        self.hop, self.nmel, self.bins = hop, num_mel_bins, 513
        self.gen = lib().lo_cng_gen(self.h)

    def __del__(self):
        if getattr(self, "h", None):
            lib().lo_cng_free(self.h)
            self.h = None

    def add_features(self, features):
        f = np.ascontiguousarray(features, dtype=np.float32)
        return lib().lo_gen_add_features(self.gen, _p(f), int(f.size)) == 0

    def num_samples_available(self):
        return lib().lo_gen_num_samples_available(self.gen)

    def generate_samples(self, n):
        out = np.zeros(max(n, 0) + 1, dtype=np.int16)
        r = lib().lo_gen_generate_samples(self.gen, n, _p(out))
        return None if r < 0 else out[:r].copy()

    def condition(self, features, phase=None):
        f = np.ascontiguousarray(features, dtype=np.float32)
        out = np.zeros(self.hop, dtype=np.int16)
        ph = None if phase is None else np.ascontiguousarray(phase, dtype=np.uint32)
        assert lib().lo_cng_condition(self.h, _p(f), None if ph is None else _p(ph), _p(out)) == 0
        return out

    @property
    def synthesis_gain(self):
        return lib().lo_cng_synthesis_gain(self.h)


def cng_phase_index(seed, hop, bin_):
    return int(lib().lo_cng_phase_index(seed, hop, bin_))


class Decoder:
    """LyraDecoder at 16 kHz with the packet-loss / comfort-noise / fade state machine (oracle/lyra_decoder.c)."""
    COUNTERS = ["vq_decode", "model_add", "model_generate", "model_last_request", "cng_add", "cng_generate", "cng_last_request",
                "noise_receive", "noise_estimate"]

    def __init__(self, model_dir=None, cng_seed=0, fake=None):
        if fake is not None:
            self.h = lib().lo_decoder_create_fake(int(fake[0]), int(fake[1]))
        else:
            self.h = lib().lo_decoder_create(model_dir.encode(), cng_seed)
        if not self.h:
            raise ValueError("lo_decoder_create failed")

    def __del__(self):
        if getattr(self, "h", None):
            lib().lo_decoder_free(self.h)
            self.h = None

    def set_encoded_packet(self, packet):
        b = np.frombuffer(bytes(packet), dtype=np.uint8).copy() if len(packet) else np.zeros(1, np.uint8)
        return lib().lo_decoder_set_encoded_packet(self.h, _p(b), len(packet)) == 0

    def decode_samples(self, n):
        out = np.zeros(max(n, 0) + 1, dtype=np.int16)
        r = lib().lo_decoder_decode_samples(self.h, n, _p(out))
        return None if r < 0 else out[:r].copy()

    def is_comfort_noise(self):
        return bool(lib().lo_decoder_is_comfort_noise(self.h))

    @property
    def state(self):
        s = np.zeros(3, dtype=np.int32)
        lib().lo_decoder_get_state(self.h, _p(s))
        return tuple(int(x) for x in s)

    @state.setter
    def state(self, v):
        s = np.asarray(v, dtype=np.int32)
        lib().lo_decoder_set_state(self.h, _p(s))

    def counters(self):
        c = np.zeros(9, dtype=np.int32)
        lib().lo_decoder_counters(self.h, _p(c))
        return dict(zip(self.COUNTERS, (int(x) for x in c)))

    def noise_estimate(self):
        out = np.zeros(160, dtype=np.float32)
        lib().lo_noise_estimate(lib().lo_decoder_noise(self.h), _p(out))
        return out


class Encoder:
    """LyraEncoder::Encode at 16 kHz, optionally with DTX (lyra/lyra_encoder.cc:113-156)."""

    def __init__(self, model_dir, enable_dtx=False):
        self.h = lib().lo_encoder_create(model_dir.encode(), 1 if enable_dtx else 0)
        if not self.h:
            raise ValueError("lo_encoder_create failed")

    def __del__(self):
        if getattr(self, "h", None):
            lib().lo_encoder_free(self.h)
            self.h = None

    def encode(self, pcm, num_bits):
        a = np.ascontiguousarray(pcm, dtype=np.int16)
        out = np.zeros(32, dtype=np.uint8)
        r = lib().lo_encoder_encode(self.h, _p(a), int(a.size), num_bits, _p(out))
        return None if r < 0 else bytes(out[:r])


class Resampler:
    """Resampler (lyra/resampler.{h,cc}): polyphase Kaiser-windowed-sinc FIR, started fully primed (17 input samples of delay)."""

    def __init__(self, input_rate_hz, output_rate_hz):
        self.h = lib().lo_resampler_create(input_rate_hz, output_rate_hz)
        if not self.h:
            raise ValueError("unsupported sample rates")
        self.input_rate_hz, self.output_rate_hz = input_rate_hz, output_rate_hz

    def __del__(self):
        if getattr(self, "h", None):
            lib().lo_resampler_free(self.h)
            self.h = None

    def reset(self):
        lib().lo_resampler_reset(self.h)

    def resample(self, audio):
        a = np.ascontiguousarray(audio, dtype=np.int16)
        cap = a.size * 3 + 8
        out = np.zeros(cap, dtype=np.int16)
        r = lib().lo_resampler_resample(self.h, _p(a), int(a.size), _p(out), cap)
        assert r >= 0
        return out[:r].copy()

    def samples_until_steady_state(self):
        return lib().lo_resampler_samples_until_steady_state(self.h)


def resampler_design(input_rate_hz, output_rate_hz):
    """-> (num, den, coeffs[den][taps] float32): the polyphase filter bank (rate ratio input / output = num / den)."""
    num, den = C.c_int(), C.c_int()
    buf = np.zeros(3 * 35, dtype=np.float32)
    taps = lib().lo_resampler_design(input_rate_hz, output_rate_hz, C.byref(num), C.byref(den), _p(buf), buf.size)
    return num.value, den.value, buf[:den.value * taps].reshape(den.value, taps).copy()


_GEN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_void_p)


class BufferedResampler:
    """BufferedResampler (lyra/buffered_resampler.{h,cc}): internal 16 kHz -> external rate with leftover bookkeeping."""

    def __init__(self, internal_rate_hz, external_rate_hz):
        self.h = lib().lo_buffered_resampler_create(internal_rate_hz, external_rate_hz)
        if not self.h:
            raise ValueError("unsupported sample rates")

    def __del__(self):
        if getattr(self, "h", None):
            lib().lo_buffered_resampler_free(self.h)
            self.h = None

    @property
    def leftover(self):
        return lib().lo_buffered_resampler_leftover(self.h)

    def internal_samples(self, num_external):
        return lib().lo_buffered_resampler_internal_samples(self.h, num_external)

    def filter_and_buffer(self, generator, num_external):
        """generator(n) -> int16[n] (or None to fail, like the reference's std::nullopt)."""
        def thunk(_user, n, dst):
            got = generator(n)
            if got is None or len(got) != n:
                return -1
            C.memmove(dst, np.ascontiguousarray(got, dtype=np.int16).ctypes.data, 2 * n)
            return 0
        out = np.zeros(num_external + 1, dtype=np.int16)
        r = lib().lo_buffered_resampler_filter_and_buffer(self.h, _GEN(thunk), None, num_external, _p(out))
        return None if r < 0 else out[:num_external].copy()


class Codec:
    """One 16 kHz stream: LyraEncoder::Encode / LyraDecoder::SetEncodedPacket+DecodeSamples(320)."""

    def __init__(self, model_dir):
        self.h = lib().lo_codec_create(model_dir.encode())
        if not self.h:
            raise RuntimeError("oracle: cannot load models from %s" % model_dir)
        self.encoder_net = Net(handle=lib().lo_codec_encoder_net(self.h), owner=self)
        self.decoder_net = Net(handle=lib().lo_codec_decoder_net(self.h), owner=self)

    def __del__(self):
        if getattr(self, "h", None):
            lib().lo_codec_free(self.h)
            self.h = None

    def reset(self):
        assert lib().lo_codec_reset(self.h) == 0

    def encode(self, pcm, num_bits):
        """-> (packet bytes, features f32[64], indices i32[46])"""
        a = np.ascontiguousarray(pcm, dtype=np.int16)
        assert a.size == 320
        pkt = np.zeros(24, dtype=np.uint8)
        feat = np.empty(64, dtype=np.float32)
        idx = np.empty(46, dtype=np.int32)
        n = lib().lo_codec_encode(self.h, _p(a), num_bits, _p(pkt), _p(feat), _p(idx))
        if n < 0:
            return None
        return bytes(pkt[:n]), feat, idx

    def decode(self, packet, num_bits):
        """packet=None -> concealment with zero features. -> (pcm i16[320], lossy features, unit float[320])"""
        pcm = np.empty(320, dtype=np.int16)
        lossy = np.empty(64, dtype=np.float32)
        unit = np.empty(320, dtype=np.float32)
        if packet is None:
            rc = lib().lo_codec_decode(self.h, None, num_bits, _p(pcm), _p(lossy), _p(unit))
        else:
            a = np.frombuffer(bytes(packet), dtype=np.uint8).copy()
            rc = lib().lo_codec_decode(self.h, _p(a), num_bits, _p(pcm), _p(lossy), _p(unit))
        return None if rc != 0 else (pcm, lossy, unit)


def cpu_bench(model_dir, streams, frames, num_bits, threads, seed=0x4C595241):
    """-> dict(wall_s, frames, frames_per_s, stage_us[4], checksum)"""
    st = (C.c_double * 4)()
    cs = C.c_uint64()
    wall = lib().lo_cpu_bench(model_dir.encode(), streams, frames, num_bits, threads, seed, st, C.byref(cs))
    if wall < 0:
        raise RuntimeError("oracle cpu bench failed")
    return dict(wall_s=wall, frames=streams * frames, frames_per_s=streams * frames / wall,
                stage_us=list(st), checksum=cs.value)
