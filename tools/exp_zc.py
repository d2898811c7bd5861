import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, '/root/repo')
import physis_b200
from physis_b200 import api
lib = physis_b200.load_programs()
lib.initialize_physis.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int]
lib.copyin_physis.argtypes = [C.c_void_p]
lib.run_sweeps_only_physis.argtypes = [C.c_int] * 4 + [C.c_float] * 7
co = [0.1] * 6 + [0.4]
for (nx, ny, nz) in [(512,512,512),(1024,1024,128),(1024,1024,512),(1024,1024,1024),(256,256,256),(512,512,128),(384,384,384)]:
    lib.initialize_physis(0, None, nx, ny, nz)
    api.set_option("star7_fuse=0")
    lib.initialize_benchmark_physis(nx, ny, nz)
    f0 = np.random.default_rng(0).random(nx * ny * nz, dtype=np.float32)
    lib.copyin_physis(f0.ctypes.data)
    r = api.rt()
    lib.run_sweeps_only_physis(4, nx, ny, nz, *co)
    r.__PSB200TimerStart()
    lib.run_sweeps_only_physis(40, nx, ny, nz, *co)
    ms = r.__PSB200TimerStopMs() / 40
    print(f"{nx}x{ny}x{nz}: {ms:.4f} ms/sweep {8.0*nx*ny*nz/ms/1e6:.0f} GB/s", flush=True)
    lib.finalize_benchmark_physis()
