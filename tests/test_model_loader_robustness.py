"""CPU tier: lyra_b200_create on damaged model files must answer LYRA_B200_EMODEL (the reference's Create() returns
nullptr, lyra/tflite_model_wrapper.cc:39-66) or load — never crash.  Runs the product's loader (tflite_model.cc,
model_spec.cc) inside the emulated library, one subprocess per damaged model directory."""
import os
import random
import shutil
import subprocess
import sys

from conftest import MODEL_DIR

CHILD = r'''
import sys, ctypes as C
lib = C.CDLL(sys.argv[1])
lib.lyra_b200_create.argtypes = [C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
h = C.c_void_p()
rc = lib.lyra_b200_create(sys.argv[2].encode(), 0, 8, C.byref(h))
assert rc in (0, -3), rc
assert (rc == 0) == bool(h.value)
print("rc", rc)
'''


def damage(data, mode, rng):
    b = bytearray(data)
    if len(b) < 1024 and mode in ("flip_header", "zero_block"):      # lyra_config.binarypb is two bytes
        mode = "flip"
    if mode == "truncate":
        return b[:rng.randrange(0, len(b))]
    if mode == "empty":
        return bytearray()
    if mode == "flip_header":
        for _ in range(rng.randrange(1, 8)):
            b[rng.randrange(min(4096, len(b)))] ^= 0xFF
        return b
    if mode == "zero_block":
        o = rng.randrange(len(b) - 512)
        b[o:o + 512] = bytes(512)
        return b
    for _ in range(rng.randrange(1, 20)):                      # scattered bit flips
        b[rng.randrange(len(b))] ^= 1 << rng.randrange(8)
    return b


def test_damaged_models_are_rejected_not_crashing(emu_api, tmp_path):
    rng = random.Random(5)
    outcomes = set()
    modes = ["truncate", "empty", "flip_header", "zero_block", "flip"]
    for trial in range(15):
        d = tmp_path / ("m%d" % trial)
        shutil.copytree(MODEL_DIR, d)
        victim = rng.choice(["soundstream_encoder.tflite", "lyragan.tflite", "quantizer.tflite", "lyra_config.binarypb"])
        p = d / victim
        p.write_bytes(bytes(damage(p.read_bytes(), modes[trial % len(modes)], rng)))
        r = subprocess.run([sys.executable, "-c", CHILD, emu_api.path, str(d)], capture_output=True, text=True, timeout=120)
        assert r.returncode == 0, "loader crashed on %s (%s): %s" % (victim, modes[trial % len(modes)], r.stderr[-400:])
        outcomes.add(r.stdout.strip())
    assert "rc -3" in outcomes      # at least the truncated / empty files are refused
    # a missing file is refused too
    d = tmp_path / "missing"
    shutil.copytree(MODEL_DIR, d)
    os.remove(d / "quantizer.tflite")
    r = subprocess.run([sys.executable, "-c", CHILD, emu_api.path, str(d)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == "rc -3", r.stdout + r.stderr
