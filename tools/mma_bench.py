import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sr_wavenet_b200 import _lib
torch.zeros(1).cuda()
lib = _lib.load()
out = (ctypes.c_longlong * 24)()
_lib.check(lib.srwn_debug_mma_bench(out))
names = ["1 warp : 4 x N32", "1 warp : 16 x N32", "1 warp : 2 x N160", "1 warp : 8 x N128",
         "warps 0,1,2 : 4 x N32 each", "warps 0,1,2 : 16 x N32 each", "warps 0,4,8 : 4 x N32 each",
         "warps 0,4,8 : 16 x N32 each", "warps 0,1,2 : 2 x N160 each", "warps 0,4,8 : 2 x N160 each",
         "warps 0,4,8,12 : 16 x N32 each", "warps 0,1,2,3 : 16 x N32 each"]
for i, n in enumerate(names):
    print("%-32s issue %5d clk   issue+complete %5d clk" % (n, out[2 * i], out[2 * i + 1]))
