import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sr_wavenet_b200 import _lib
torch.zeros(1).cuda()
lib = _lib.load()
out = (ctypes.c_longlong * 16)()
_lib.check(lib.srwn_debug_mma_bench(out))
names = ["1 x N32", "4 x N32 (same acc)", "16 x N32 (same acc)", "4 x N128", "4 x N160", "4 x N256", "16 x N32 (8 accs)", "16 x N160"]
for i, n in enumerate(names):
    print("%-22s issue %5d clk   issue+complete %5d clk" % (n, out[2 * i], out[2 * i + 1]))
