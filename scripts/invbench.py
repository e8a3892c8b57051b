import ctypes, sys
sys.path.insert(0,'/root/repo')
from montgomery_b200 import _native
lib=_native.lib()
for mode,bps,thr in ((8,1,32),(9,1,32),(9,4,128),(8,4,128),(9,1,128)):
    ops=ctypes.c_double(); ms=ctypes.c_float()
    iters=50
    rc=lib.mgb_microbench(0,mode,bps,thr,iters,ctypes.byref(ops),ctypes.byref(ms))
    print(mode,bps,thr,rc,"ms per inversion (latency): %.4f"%(ms.value/iters), "inv/s %.3g"%ops.value)
