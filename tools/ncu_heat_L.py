"""One heat_3d preset-L call for ncu (development aid): python tools/ncu_heat_L.py [mode]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import npbench_b200 as nb
nb.init(0)
L = nb.lib()
n, ts = 70, 100
A = nb.DeviceArray((n, n, n)); B = nb.DeviceArray((n, n, n))
L.init_heat3d_f64(n, 0, n, A.ptr, B.ptr)
L.heat3d_set_mode(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
for _ in range(3):
    nb.heat_3d(ts, A, B)
L.sync()
print("path", L.heat3d_last_path())
