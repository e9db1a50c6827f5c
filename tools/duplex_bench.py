#!/usr/bin/env python3
"""Dev tool (GPU): encode and decode as two independent pipelines (two contexts, two CUDA streams): decode(i) waits for
encode(i)'s packets only, so encode(i+1) overlaps decode(i).  Compares with the serial step of bench.py."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa
from lyra_b200 import _capi  # noqa


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    steps = 100
    enc, dec = _capi.Context(n), _capi.Context(n)
    rng = np.random.default_rng(0)
    NB = 4
    d_pcm = [torch.from_numpy(rng.integers(-8192, 8192, size=(n, 320), dtype=np.int16)).cuda() for _ in range(NB)]
    d_pk = [torch.zeros((n, 8), dtype=torch.uint8, device="cuda") for _ in range(NB)]
    d_out = torch.zeros((n, 320), dtype=torch.int16, device="cuda")
    sx, sy = torch.cuda.Stream(), torch.cuda.Stream()
    enc.set_stream(sx.cuda_stream)
    dec.set_stream(sy.cuda_stream)
    ev_pk = [torch.cuda.Event() for _ in range(NB)]      # packets of buffer b written
    ev_free = [torch.cuda.Event() for _ in range(NB)]    # packets of buffer b consumed

    def run(k):
        for i in range(k):
            b = i % NB
            if i >= NB:
                sx.wait_event(ev_free[b])
            enc.encode_device(n, d_pcm[b].data_ptr(), 64, d_pk[b].data_ptr())
            ev_pk[b].record(sx)
            sy.wait_event(ev_pk[b])
            dec.decode_device(n, d_pk[b].data_ptr(), 0, 64, d_out.data_ptr())
            ev_free[b].record(sy)

    run(8)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(sx)
    run(steps)
    sx.wait_stream(sy)
    e1.record(sx)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print("duplex pipeline: %.4f ms per step (%.2fM f/s), checksum %d" % (ms, n / ms / 1e3, int(d_out.to(torch.int64).sum().item())))


if __name__ == "__main__":
    main()
