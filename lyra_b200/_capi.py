"""ctypes binding of the C ABI in include/lyra_b200.h.

``load()`` binds the nvcc-built product library ``lyra_b200/liblyra_b200.so`` and nothing else; it
raises if the library is missing (run ``python -c 'import __graft_entry__ as g; g.build()'``).  There
is no CPU fallback.  (``CApi(path)`` exists so the CPU test tier can bind the test-only emulated
build of the same sources; the package itself never does.)
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PRODUCT_SO = os.path.join(_HERE, "liblyra_b200.so")
MODEL_DIR = os.path.join(_HERE, "model_coeffs")

OK, EINVAL, ENODEV, EMODEL = 0, -1, -2, -3
HOP = 320
NUM_FEATURES = 64
MAX_STAGES = 46


class LyraB200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("lyra_b200 error %d: %s" % (code, msg))
        self.code = code


class CApi:
    def __init__(self, so_path):
        if not os.path.exists(so_path):
            raise FileNotFoundError(
                "%s not found: the CUDA extension is not built (python -c 'import __graft_entry__ as g; g.build()')" % so_path)
        L = C.CDLL(so_path)
        vp, ci = C.c_void_p, C.c_int
        sig = {
            "lyra_b200_create": (ci, [C.c_char_p, ci, ci, C.POINTER(vp)]),
            "lyra_b200_create_ex": (ci, [C.c_char_p, ci, ci, ci, C.POINTER(vp)]),
            "lyra_b200_destroy": (None, [vp]),
            "lyra_b200_last_error": (C.c_char_p, [vp]),
            "lyra_b200_max_streams": (ci, [vp]),
            "lyra_b200_tile_streams": (ci, [vp]),
            "lyra_b200_reset": (ci, [vp, vp, ci]),
            "lyra_b200_encode": (ci, [vp, vp, ci, vp, ci, vp]),
            "lyra_b200_decode": (ci, [vp, vp, ci, vp, vp, ci, vp]),
            "lyra_b200_extract_features": (ci, [vp, vp, ci, vp, vp]),
            "lyra_b200_quantize": (ci, [vp, ci, vp, ci, vp, vp]),
            "lyra_b200_dequantize": (ci, [vp, ci, vp, ci, vp]),
            "lyra_b200_generate": (ci, [vp, vp, ci, vp, vp]),
            "lyra_b200_logmel": (ci, [vp, ci, vp, ci, vp, ci, vp]),
            "lyra_b200_set_stream": (ci, [vp, vp]),
            "lyra_b200_encode_device": (ci, [vp, ci, vp, ci, vp]),
            "lyra_b200_decode_device": (ci, [vp, ci, vp, vp, ci, vp]),
            "lyra_b200_synchronize": (ci, [vp]),
            "lyra_b200_noise_update": (ci, [vp, vp, ci, vp, vp, vp, vp]),
            "lyra_b200_decode_track_noise": (ci, [vp, vp, ci, vp, vp, ci, vp, vp]),
            "lyra_b200_decode_track_noise_device": (ci, [vp, ci, vp, vp, ci, vp, vp]),
            "lyra_b200_noise_update_device": (ci, [vp, ci, vp, vp, vp, vp]),
            "lyra_b200_set_split": (ci, [vp, ci]),
            "lyra_b200_set_blocking_sync": (ci, [vp, ci]),
            "lyra_b200_set_graphs": (ci, [vp, ci]),
            "lyra_b200_set_priority": (ci, [vp, ci]),
            "lyra_b200_graph_replays": (C.c_uint64, [vp]),
            "lyra_b200_set_decoder_mode": (ci, [vp, ci]),
            "lyra_b200_decoder_mode": (ci, [vp]),
            "lyra_b200_launch_count": (C.c_uint64, [vp]),
            "lyra_b200_profile_enable": (ci, [vp, ci]),
            "lyra_b200_profile_read": (ci, [vp, vp, vp]),
            "lyra_b200_noise_estimate": (ci, [vp, vp, ci, vp, vp]),
            "lyra_b200_decode_plc": (ci, [vp, vp, ci, vp, vp, ci, vp, vp]),
            "lyra_b200_decode_plc_device": (ci, [vp, ci, vp, vp, ci, vp, vp]),
            "lyra_b200_plc_get_state": (ci, [vp, vp, ci, vp]),
            "lyra_b200_plc_set_state": (ci, [vp, vp, ci, vp]),
            "lyra_b200_cng_generate": (ci, [vp, vp, ci, vp, vp]),
            "lyra_b200_set_cng_seed": (ci, [vp, C.c_uint64]),
            "lyra_b200_encode_dtx": (ci, [vp, vp, ci, vp, ci, vp, vp]),
            "lyra_b200_encode_dtx_device": (ci, [vp, ci, vp, ci, vp, vp]),
            "lyra_b200_resample": (ci, [vp, ci, vp, ci, ci, vp, ci, vp, ci, vp]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(L, name)   # AttributeError here = the library does not export the declared ABI
            fn.restype = res
            fn.argtypes = args
        self.lib = L
        self.path = so_path

    EXPORTS = ["lyra_b200_create", "lyra_b200_create_ex", "lyra_b200_destroy", "lyra_b200_last_error", "lyra_b200_max_streams",
               "lyra_b200_tile_streams", "lyra_b200_reset", "lyra_b200_encode", "lyra_b200_decode",
               "lyra_b200_extract_features", "lyra_b200_quantize", "lyra_b200_dequantize", "lyra_b200_generate",
               "lyra_b200_logmel", "lyra_b200_set_stream", "lyra_b200_encode_device", "lyra_b200_decode_device",
               "lyra_b200_synchronize", "lyra_b200_noise_update", "lyra_b200_noise_update_device", "lyra_b200_decode_track_noise",
               "lyra_b200_decode_track_noise_device", "lyra_b200_set_split", "lyra_b200_set_blocking_sync", "lyra_b200_set_graphs", "lyra_b200_set_priority", "lyra_b200_graph_replays", "lyra_b200_set_decoder_mode", "lyra_b200_decoder_mode", "lyra_b200_launch_count", "lyra_b200_profile_enable",
               "lyra_b200_profile_read", "lyra_b200_noise_estimate", "lyra_b200_decode_plc", "lyra_b200_decode_plc_device",
               "lyra_b200_plc_get_state", "lyra_b200_plc_set_state", "lyra_b200_cng_generate", "lyra_b200_set_cng_seed",
               "lyra_b200_encode_dtx", "lyra_b200_encode_dtx_device", "lyra_b200_resample"]


_product = None


def load():
    """Bind the product library (nvcc build). Fails loudly if it is missing."""
    global _product
    if _product is None:
        _product = CApi(os.environ.get("LYRA_B200_LIB", PRODUCT_SO))   # the override is for kernel-variant experiments (tools/)
    return _product


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _ids(stream_ids, n):
    if stream_ids is None:
        return None
    a = np.ascontiguousarray(stream_ids, dtype=np.int32)
    if a.size != n:
        raise ValueError("stream_ids must have one id per row")
    return a


def _mask(mask, n, what):
    if mask is None:
        return None
    a = np.ascontiguousarray(mask, dtype=np.uint8)
    if a.size != n:
        raise ValueError("%s must have one entry per row" % what)
    return a


def packet_bytes(num_bits):
    return (num_bits + 7) // 8


class Context:
    """One GPU context: weights + the streaming state of ``max_streams`` independent 16 kHz streams."""

    ROLES = {"both": 3, "encoder": 1, "decoder": 2}

    def __init__(self, max_streams, model_dir=MODEL_DIR, device=0, capi=None, roles="both"):
        self.api = capi or load()
        h = C.c_void_p()
        rc = self.api.lib.lyra_b200_create_ex(str(model_dir).encode(), int(device), int(max_streams), self.ROLES[roles], C.byref(h))
        if rc != OK:
            raise LyraB200Error(rc, (self.api.lib.lyra_b200_last_error(None) or b"").decode())
        self.h = h
        self.max_streams = max_streams

    def close(self):
        if getattr(self, "h", None):
            self.api.lib.lyra_b200_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def _check(self, rc):
        if rc != OK:
            raise LyraB200Error(rc, (self.api.lib.lyra_b200_last_error(self.h) or b"").decode())

    @property
    def launch_count(self):
        return int(self.api.lib.lyra_b200_launch_count(self.h))

    @property
    def tile_streams(self):
        return int(self.api.lib.lyra_b200_tile_streams(self.h))

    def reset(self, stream_ids=None, n=None):
        if stream_ids is None:
            n = self.max_streams if n is None else n
            self._check(self.api.lib.lyra_b200_reset(self.h, None, n))
        else:
            a = np.ascontiguousarray(stream_ids, dtype=np.int32)
            self._check(self.api.lib.lyra_b200_reset(self.h, _ptr(a), a.size))

    def encode(self, pcm, num_bits, stream_ids=None):
        pcm = np.ascontiguousarray(pcm, dtype=np.int16).reshape(-1, HOP)
        n = pcm.shape[0]
        ids = _ids(stream_ids, n)
        out = np.empty((n, packet_bytes(num_bits)), dtype=np.uint8)
        self._check(self.api.lib.lyra_b200_encode(self.h, _ptr(ids), n, _ptr(pcm), num_bits, _ptr(out)))
        return out

    def decode(self, packets, num_bits, stream_ids=None, received=None):
        packets = np.ascontiguousarray(packets, dtype=np.uint8).reshape(-1, packet_bytes(num_bits))
        n = packets.shape[0]
        ids = _ids(stream_ids, n)
        rec = _mask(received, n, "received")
        out = np.empty((n, HOP), dtype=np.int16)
        self._check(self.api.lib.lyra_b200_decode(self.h, _ptr(ids), n, _ptr(packets), _ptr(rec), num_bits, _ptr(out)))
        return out

    def extract_features(self, pcm, stream_ids=None):
        pcm = np.ascontiguousarray(pcm, dtype=np.int16).reshape(-1, HOP)
        n = pcm.shape[0]
        ids = _ids(stream_ids, n)
        out = np.empty((n, NUM_FEATURES), dtype=np.float32)
        self._check(self.api.lib.lyra_b200_extract_features(self.h, _ptr(ids), n, _ptr(pcm), _ptr(out)))
        return out

    def quantize(self, features, num_bits, want_indices=False):
        f = np.ascontiguousarray(features, dtype=np.float32).reshape(-1, NUM_FEATURES)
        n = f.shape[0]
        out = np.empty((n, packet_bytes(num_bits)), dtype=np.uint8)
        idx = np.empty((n, MAX_STAGES), dtype=np.int32) if want_indices else None
        self._check(self.api.lib.lyra_b200_quantize(self.h, n, _ptr(f), num_bits, _ptr(out), _ptr(idx)))
        return (out, idx) if want_indices else out

    def dequantize(self, packets, num_bits):
        p = np.ascontiguousarray(packets, dtype=np.uint8).reshape(-1, packet_bytes(num_bits))
        out = np.empty((p.shape[0], NUM_FEATURES), dtype=np.float32)
        self._check(self.api.lib.lyra_b200_dequantize(self.h, p.shape[0], _ptr(p), num_bits, _ptr(out)))
        return out

    def generate(self, features, stream_ids=None):
        f = np.ascontiguousarray(features, dtype=np.float32).reshape(-1, NUM_FEATURES)
        n = f.shape[0]
        ids = _ids(stream_ids, n)
        out = np.empty((n, HOP), dtype=np.int16)
        self._check(self.api.lib.lyra_b200_generate(self.h, _ptr(ids), n, _ptr(f), _ptr(out)))
        return out

    def logmel(self, pcm, num_mel_bins=160, bank=0, stream_ids=None):
        pcm = np.ascontiguousarray(pcm, dtype=np.int16).reshape(-1, HOP)
        n = pcm.shape[0]
        ids = _ids(stream_ids, n)
        out = np.empty((n, num_mel_bins), dtype=np.float32)
        self._check(self.api.lib.lyra_b200_logmel(self.h, bank, _ptr(ids), n, _ptr(pcm), num_mel_bins, _ptr(out)))
        return out

    # ---- device-resident variants (raw CUDA device pointers as ints, e.g. torch.Tensor.data_ptr()) ----
    def set_stream(self, cuda_stream_ptr):
        self._check(self.api.lib.lyra_b200_set_stream(self.h, C.c_void_p(cuda_stream_ptr or 0)))

    def decode_track_noise(self, packets, num_bits, stream_ids=None, received=None):
        """decode() + noise-estimator update of the received streams on the device -> (pcm[n][320], is_noise[n])."""
        packets = np.ascontiguousarray(packets, dtype=np.uint8).reshape(-1, packet_bytes(num_bits))
        n = packets.shape[0]
        ids = _ids(stream_ids, n)
        rec = _mask(received, n, "received")
        out = np.empty((n, HOP), dtype=np.int16)
        flags = np.empty(n, dtype=np.uint8)
        self._check(self.api.lib.lyra_b200_decode_track_noise(self.h, _ptr(ids), n, _ptr(packets), _ptr(rec), num_bits, _ptr(out), _ptr(flags)))
        return out, flags.astype(bool)

    def decode_track_noise_device(self, n, d_packets, d_received, num_bits, d_pcm, d_is_noise=0):
        self._check(self.api.lib.lyra_b200_decode_track_noise_device(self.h, n, C.c_void_p(d_packets), C.c_void_p(d_received or 0),
                                                                     num_bits, C.c_void_p(d_pcm), C.c_void_p(d_is_noise or 0)))

    def noise_update(self, pcm, stream_ids=None, update_mask=None):
        """NoiseEstimator::ReceiveSamples on one decoded hop per stream -> (is_noise[n] bool, noise_estimate[n][160])."""
        pcm = np.ascontiguousarray(pcm, dtype=np.int16).reshape(-1, HOP)
        n = pcm.shape[0]
        ids = _ids(stream_ids, n)
        mask = _mask(update_mask, n, "update_mask")
        flags = np.empty(n, dtype=np.uint8)
        est = np.empty((n, 160), dtype=np.float32)
        self._check(self.api.lib.lyra_b200_noise_update(self.h, _ptr(ids), n, _ptr(pcm), _ptr(mask), _ptr(flags), _ptr(est)))
        return flags.astype(bool), est

    def noise_estimate(self, n=None, stream_ids=None):
        """Read-only: (noise_estimate[n][160], is_noise[n] bool) of the decoder-side estimators."""
        ids = None if stream_ids is None else np.ascontiguousarray(stream_ids, dtype=np.int32)
        n = ids.size if ids is not None else (self.max_streams if n is None else n)
        est = np.empty((n, 160), dtype=np.float32)
        flags = np.empty(n, dtype=np.uint8)
        self._check(self.api.lib.lyra_b200_noise_estimate(self.h, _ptr(ids), n, _ptr(est), _ptr(flags)))
        return est, flags.astype(bool)

    # ---- packet-loss concealment / comfort noise / DTX ----
    def decode_plc(self, packets, num_bits, stream_ids=None, received=None):
        """One tick of LyraDecoder with the full concealment / comfort-noise / fade behaviour -> (pcm[n][320], is_comfort_noise[n])."""
        packets = np.ascontiguousarray(packets, dtype=np.uint8).reshape(-1, packet_bytes(num_bits))
        n = packets.shape[0]
        ids = _ids(stream_ids, n)
        rec = _mask(received, n, "received")
        out = np.empty((n, HOP), dtype=np.int16)
        cn = np.empty(n, dtype=np.uint8)
        self._check(self.api.lib.lyra_b200_decode_plc(self.h, _ptr(ids), n, _ptr(packets), _ptr(rec), num_bits, _ptr(out), _ptr(cn)))
        return out, cn.astype(bool)

    def decode_plc_device(self, n, d_packets, d_received, num_bits, d_pcm, d_is_cn=0):
        self._check(self.api.lib.lyra_b200_decode_plc_device(self.h, n, C.c_void_p(d_packets), C.c_void_p(d_received or 0), num_bits,
                                                             C.c_void_p(d_pcm), C.c_void_p(d_is_cn or 0)))

    def plc_state(self, n=None, stream_ids=None):
        ids = None if stream_ids is None else np.ascontiguousarray(stream_ids, dtype=np.int32)
        n = ids.size if ids is not None else (self.max_streams if n is None else n)
        st = np.empty((n, 3), dtype=np.int32)
        self._check(self.api.lib.lyra_b200_plc_get_state(self.h, _ptr(ids), n, _ptr(st)))
        return st

    def set_plc_state(self, state, stream_ids=None):
        st = np.ascontiguousarray(state, dtype=np.int32).reshape(-1, 3)
        ids = _ids(stream_ids, st.shape[0])
        self._check(self.api.lib.lyra_b200_plc_set_state(self.h, _ptr(ids), st.shape[0], _ptr(st)))

    def cng_generate(self, features, stream_ids=None):
        f = np.ascontiguousarray(features, dtype=np.float32).reshape(-1, 160)
        n = f.shape[0]
        ids = _ids(stream_ids, n)
        out = np.empty((n, HOP), dtype=np.int16)
        self._check(self.api.lib.lyra_b200_cng_generate(self.h, _ptr(ids), n, _ptr(f), _ptr(out)))
        return out

    def set_cng_seed(self, seed):
        self._check(self.api.lib.lyra_b200_set_cng_seed(self.h, int(seed)))

    def encode_dtx(self, pcm, num_bits, stream_ids=None):
        """LyraEncoder::Encode with DTX -> (packets[n][P], packet_bytes[n]: P, or 0 for an empty (noise) packet)."""
        pcm = np.ascontiguousarray(pcm, dtype=np.int16).reshape(-1, HOP)
        n = pcm.shape[0]
        ids = _ids(stream_ids, n)
        out = np.empty((n, packet_bytes(num_bits)), dtype=np.uint8)
        sizes = np.empty(n, dtype=np.int32)
        self._check(self.api.lib.lyra_b200_encode_dtx(self.h, _ptr(ids), n, _ptr(pcm), num_bits, _ptr(out), _ptr(sizes)))
        return out, sizes

    def encode_dtx_device(self, n, d_pcm, num_bits, d_packets, d_is_noise):
        self._check(self.api.lib.lyra_b200_encode_dtx_device(self.h, n, C.c_void_p(d_pcm), num_bits, C.c_void_p(d_packets), C.c_void_p(d_is_noise)))

    def resample(self, audio, external_rate_hz, to_internal, stream_ids=None):
        """Resampler::Resample for n streams -> list of int16 arrays (one per stream; lengths may differ by one when down-sampling)."""
        a = np.ascontiguousarray(audio, dtype=np.int16)
        a = a.reshape(1, -1) if a.ndim == 1 else a
        n, n_in = a.shape
        ids = _ids(stream_ids, n)
        ratio = (16000 / external_rate_hz) if to_internal else (external_rate_hz / 16000)
        stride = int(np.ceil(n_in * ratio)) + 1
        out = np.zeros((n, stride), dtype=np.int16)
        counts = np.zeros(n, dtype=np.int32)
        self._check(self.api.lib.lyra_b200_resample(self.h, 1 if to_internal else 0, _ptr(ids), n, int(external_rate_hz), _ptr(a), n_in,
                                                    _ptr(out), stride, _ptr(counts)))
        return [out[k, :counts[k]].copy() for k in range(n)]

    def noise_update_device(self, n, d_pcm, d_mask, d_is_noise, d_estimate):
        self._check(self.api.lib.lyra_b200_noise_update_device(self.h, n, C.c_void_p(d_pcm), C.c_void_p(d_mask or 0),
                                                               C.c_void_p(d_is_noise or 0), C.c_void_p(d_estimate or 0)))

    def encode_device(self, n, d_pcm, num_bits, d_packets):
        self._check(self.api.lib.lyra_b200_encode_device(self.h, n, C.c_void_p(d_pcm), num_bits, C.c_void_p(d_packets)))

    def decode_device(self, n, d_packets, d_received, num_bits, d_pcm):
        self._check(self.api.lib.lyra_b200_decode_device(self.h, n, C.c_void_p(d_packets), C.c_void_p(d_received or 0),
                                                         num_bits, C.c_void_p(d_pcm)))

    def synchronize(self):
        self._check(self.api.lib.lyra_b200_synchronize(self.h))

    def set_decoder_mode(self, mode):
        """mode: "exact" (bit-identical PCM, default) or "tensor" (split-precision TF32 tensor-core decoder)."""
        self._check(self.api.lib.lyra_b200_set_decoder_mode(self.h, {"exact": 0, "tensor": 1}[mode]))

    def set_blocking_sync(self, enable):
        self._check(self.api.lib.lyra_b200_set_blocking_sync(self.h, 1 if enable else 0))

    def set_graphs(self, enable):
        self._check(self.api.lib.lyra_b200_set_graphs(self.h, 1 if enable else 0))

    def set_priority(self, priority):
        self._check(self.api.lib.lyra_b200_set_priority(self.h, int(priority)))

    def graph_replays(self):
        return int(self.api.lib.lyra_b200_graph_replays(self.h))

    def set_split(self, parts):
        self._check(self.api.lib.lyra_b200_set_split(self.h, int(parts)))

    KERNEL_NAMES = ["EncoderKernelA", "EncoderKernelB", "RvqEncodeKernel", "RvqDecodeKernel", "DecoderKernelC",
                    "DecoderKernelD", "LogMelKernel", "NoiseEstimatorKernel"]

    def profile_enable(self, enable=True):
        self._check(self.api.lib.lyra_b200_profile_enable(self.h, 1 if enable else 0))

    def profile_read(self):
        """-> {kernel name: (total ms, launches)} since profiling was enabled (CUDA events on the launch stream)."""
        ms = (C.c_double * len(self.KERNEL_NAMES))()
        cnt = (C.c_uint64 * len(self.KERNEL_NAMES))()
        self._check(self.api.lib.lyra_b200_profile_read(self.h, ms, cnt))
        return {k: (ms[i], int(cnt[i])) for i, k in enumerate(self.KERNEL_NAMES)}
