import sys, ctypes
sys.path.insert(0, ".")
from tools.quick_gpu import msm_time, L, lib
for rep in range(2):
    for on in (1, 0):
        L.check(lib.h2_set_glv(on))
        print("glv", on, end=" : ")
        msm_time(20, reps=10)
        print("glv", on, end=" : ")
        msm_time(14, curve=1, reps=10, extra=1)
