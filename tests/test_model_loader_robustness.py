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


# Targeted metadata fuzzing: overwrite single 32-bit words that look like flatbuffer metadata (buffer indices,
# vtable offsets, shape entries, vector lengths - anything currently holding a small integer) and require the loader to answer
# OK or EMODEL for every one of them.  Many mutations run inside one child process; a crash kills the child.
CHILD_WORDS = r'''
import sys, os, shutil, struct, ctypes as C
lib = C.CDLL(sys.argv[1])
lib.lyra_b200_create.argtypes = [C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
lib.lyra_b200_destroy.argtypes = [C.c_void_p]
src, work = sys.argv[2], sys.argv[3]
counts = {0: 0, -3: 0}
for line in open(sys.argv[4]):
    name, off, val = line.split()
    off, val = int(off), int(val)
    path = os.path.join(work, name)
    orig = open(os.path.join(src, name), "rb").read()
    with open(path, "wb") as f:
        f.write(orig[:off] + struct.pack("<I", val) + orig[off + 4:])
    h = C.c_void_p()
    rc = lib.lyra_b200_create(work.encode(), 0, 8, C.byref(h))
    assert rc in (0, -3), (name, off, val, rc)
    assert (rc == 0) == bool(h.value)
    if h.value:
        lib.lyra_b200_destroy(h)
    counts[rc] += 1
    with open(path, "wb") as f:
        f.write(orig)
    print("done", name, off, val, rc, flush=True)
print("ok", counts[0], counts[-3])
'''


def test_metadata_word_mutations_never_crash(emu_api, tmp_path):
    import struct
    rng = random.Random(11)
    work = tmp_path / "work"
    shutil.copytree(MODEL_DIR, work)
    lines = []
    for name in ("soundstream_encoder.tflite", "lyragan.tflite", "quantizer.tflite"):
        data = (work / name).read_bytes()
        words = struct.unpack("<%dI" % (len(data) // 4), data[:len(data) // 4 * 4])
        cand = [i for i, w in enumerate(words) if w < 4096 or w >= 0xFFFFF000]       # small ints / small negative offsets
        for i in rng.sample(cand, 70):
            val = rng.choice([0, 1, 2, 3, words[i] + 1, words[i] ^ 1, 0x7FFFFFFF, 0xFFFFFFFF, 0x80000000, len(data), len(data) - 2])
            lines.append("%s %d %d" % (name, 4 * i, val & 0xFFFFFFFF))
    # the reproducer from the advisor's report
    lines.append("lyragan.tflite 1446900 2")
    plan = tmp_path / "plan.txt"
    plan.write_text("\n".join(lines) + "\n")
    r = subprocess.run([sys.executable, "-c", CHILD_WORDS, emu_api.path, MODEL_DIR, str(work), str(plan)],
                       capture_output=True, text=True, timeout=900)
    last = [ln for ln in r.stdout.splitlines() if ln.startswith("done")][-1:] or ["(none)"]
    assert r.returncode == 0, "loader crashed after %s: %s" % (last[0], r.stderr[-400:])
    ok, refused = map(int, r.stdout.strip().splitlines()[-1].split()[1:])
    assert ok + refused == len(lines) and refused > 0
