#!/usr/bin/env python3
"""Development probe run under gpurun: parity vs the oracle on a sample of streams + first timings."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from lyra_b200 import _capi  # noqa: E402
from oracle import oracle as O  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    frames = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    bits = 64
    print(torch.cuda.get_device_name(0))
    ctx = _capi.Context(n)
    rng = np.random.default_rng(3)
    check = sorted(set([0, 1, 15, 16, 17, n // 2, n - 1]))
    codecs = {k: O.Codec(_capi.MODEL_DIR) for k in check}
    bad = 0
    for f in range(frames):
        pcm = rng.integers(-8192, 8192, size=(n, 320), dtype=np.int16)
        t0 = time.time()
        pk = ctx.encode(pcm, bits)
        out = ctx.decode(pk, bits)
        dt = time.time() - t0
        for k in check:
            opkt, _, _ = codecs[k].encode(pcm[k], bits)
            opcm, _, _ = codecs[k].decode(opkt, bits)
            if bytes(pk[k]) != opkt or not np.array_equal(out[k], opcm):
                bad += 1
                print("MISMATCH frame", f, "stream", k, bytes(pk[k]).hex(), opkt.hex(), int((out[k] != opcm).sum()))
        print("frame %d host-api enc+dec %.2f ms" % (f, dt * 1e3))
    print("parity mismatches:", bad)
    # device-resident timing
    d_pcm = torch.from_numpy(rng.integers(-8192, 8192, size=(n, 320), dtype=np.int16)).cuda()
    d_pk = torch.zeros((n, 8), dtype=torch.uint8, device="cuda")
    d_out = torch.zeros((n, 320), dtype=torch.int16, device="cuda")
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    for _ in range(3):
        ctx.encode_device(n, d_pcm.data_ptr(), bits, d_pk.data_ptr())
        ctx.decode_device(n, d_pk.data_ptr(), 0, bits, d_out.data_ptr())
    torch.cuda.synchronize()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    reps = 10
    te = td = 0.0
    for _ in range(reps):
        evs[0].record()
        ctx.encode_device(n, d_pcm.data_ptr(), bits, d_pk.data_ptr())
        evs[1].record()
        ctx.decode_device(n, d_pk.data_ptr(), 0, bits, d_out.data_ptr())
        evs[2].record()
        torch.cuda.synchronize()
        te += evs[0].elapsed_time(evs[1])
        td += evs[1].elapsed_time(evs[2])
    print("device-resident: encode %.3f ms decode %.3f ms per step of %d streams -> %.0f frames/s"
          % (te / reps, td / reps, n, n / ((te + td) / reps / 1e3)))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
