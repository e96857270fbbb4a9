import sys, ctypes as C, numpy as np
sys.path.insert(0,'depth-map-fusion-utils_b200')
import dmf_b200 as D
from dmf_b200._lib import check
ctx = D.Context(0)
m = np.zeros(5, np.uint64)
check(ctx.lib.dmf_selftest_div1000(ctx.h, m.ctypes.data_as(C.POINTER(C.c_uint64))))
print("div1000 mismatches long/short:", m[:2], "failing |a| bit range (long form, finite):", hex(int(m[3])), hex(int(m[2])), "as floats", np.array([m[3], m[2]], np.uint32).view(np.float32), "short-form failures above 2^-101:", m[4])
