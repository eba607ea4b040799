"""Prints the clock64 timeline of CTA 0 for a few layers of one chunk (SRWN_TRACE=<chunk index>)."""
import sys, os, ctypes
os.environ.setdefault("SRWN_TRACE", "20")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sr_wavenet_b200 as srwn
from sr_wavenet_b200 import synth, _lib
B, T = 32, 64000
dil = synth.DEFAULT_DILATIONS
m = srwn.WaveNetAutoEncoder(T, 0, 5, dil, skip_channels=128, latent_channels=32, pool_stride=128)
m.set_weights(synth.make_teacher_weights(dil))
x = torch.from_numpy(synth.synthetic_audio(B, T)).cuda(); enc = torch.from_numpy(synth.synthetic_encoding(B, T // 128)).cuda()
for _ in range(2):
    m.nll(x, enc, precision="fp16")
eng = m._eng
ws, wsn = eng.workspace(_lib.OP_TEACHER_NLL, B, T, _lib.FP16)
n = 7 * 40 * 12
out = (ctypes.c_longlong * n)()
lib = _lib.load()
lib.srwn_debug_read_trace.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_int32]
_lib.check(lib.srwn_debug_read_trace(eng.h, B, T, ws, wsn, out, n))
tr = np.array(out[:], dtype=np.int64).reshape(7, 40, 12)
L0, L1 = int(sys.argv[1]) if len(sys.argv) > 1 else 10, int(sys.argv[2]) if len(sys.argv) > 2 else 14
base = tr[3, L0, 0]
names_e = ["waitD1", "D1rdy", "ld1", "math1+st", "fence", "arriveC", "D2rdy", "ld2", "math2", "waits", "stores", "arriveH"]
names_i = ["start", "W/HALO/H ok", "G1 issued", "C ok", "G2 issued"]
names_l = ["start", "Wempty ok", "G1 ok", "halo arrive"]
for l in range(L0, L1):
    print("layer %d (d=%d)" % (l, dil[l]))
    for t in range(3):
        print("  epi%d : " % t + " ".join("%s=%d" % (nm, tr[t, l, i] - base) for i, nm in enumerate(names_e)))
    for t in range(3):
        print("  iss%d : " % t + " ".join("%s=%d" % (nm, tr[3 + t, l, i] - base) for i, nm in enumerate(names_i)))
    print("  load : " + " ".join("%s=%d" % (nm, tr[6, l, i] - base) for i, nm in enumerate(names_l)))
per = (tr[3, 29, 0] - tr[3, 1, 0]) / 28.0
print("avg clocks per layer (issuer 0):", per)
