#!/usr/bin/env python3
"""Regenerates tests/golden/oracle_sample1.json from the CPU oracle (oracle/), the only runnable
statement of the reference algorithm in this environment (the reference itself cannot be built or
imported offline).  These vectors therefore detect DRIFT of the oracle / kernels; the
reference-pinned fixtures live in tests/test_oracle_golden.py (log-mel KAT, packet layouts, RVQ
fixture, integration LSD bound)."""
import json
import os
import sys
import wave

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

MODEL_DIR = os.path.join(ROOT, "lyra_b200", "model_coeffs")
with wave.open(os.path.join(ROOT, "tests", "data", "sample1_16kHz.wav")) as w:
    pcm = np.frombuffer(w.readframes(w.getnframes()), dtype=np.int16)
out = {"source": "tests/data/sample1_16kHz.wav (lyra/testdata/sample1_16kHz.wav), first 40 hops", "hops": 40}
for bits in (64, 120, 184):
    c = O.Codec(MODEL_DIR)
    pk, cs = [], []
    for h in range(40):
        p, _, _ = c.encode(pcm[320 * h:320 * h + 320], bits)
        d, _, _ = c.decode(p, bits)
        pk.append(p.hex())
        cs.append(int((d.astype(np.int64) * np.arange(1, 321)).sum()))
    out["packets_%d" % bits] = pk
    out["pcm_checksum_%d" % bits] = cs
with open(os.path.join(ROOT, "tests", "golden", "oracle_sample1.json"), "w") as f:
    json.dump(out, f, indent=1)
print("wrote oracle_sample1.json")
