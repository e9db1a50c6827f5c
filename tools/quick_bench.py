#!/usr/bin/env python3
"""Dev tool (GPU): device-resident step time of an arbitrary build of the library (variants in devtools_build/)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa
from lyra_b200 import _capi  # noqa


def main():
    so = sys.argv[1]
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
    api = _capi.CApi(so)
    ctx = _capi.Context(n, capi=api)
    rng = np.random.default_rng(0)
    d_pcm = torch.from_numpy(rng.integers(-8192, 8192, size=(n, 320), dtype=np.int16)).cuda()
    d_pk = torch.zeros((n, 8), dtype=torch.uint8, device="cuda")
    d_out = torch.zeros((n, 320), dtype=torch.int16, device="cuda")
    st = torch.cuda.current_stream()
    ctx.set_stream(st.cuda_stream)
    res = []
    for split in [int(x) for x in os.environ.get("SPLITS", "1,2").split(",")]:
        ctx.set_split(split)
        for _ in range(5):
            ctx.encode_device(n, d_pcm.data_ptr(), 64, d_pk.data_ptr())
            ctx.decode_device(n, d_pk.data_ptr(), 0, 64, d_out.data_ptr())
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(100):
            ctx.encode_device(n, d_pcm.data_ptr(), 64, d_pk.data_ptr())
            ctx.decode_device(n, d_pk.data_ptr(), 0, 64, d_out.data_ptr())
        e1.record(st)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 100
        res.append("split%d %.4f ms (%.2fM f/s)" % (split, ms, n / ms / 1e3))
    print("%-44s %s  checksum %d" % (os.path.basename(so), "  ".join(res), int(d_out.to(torch.int64).sum().item())))
    ctx.close()


if __name__ == "__main__":
    main()
