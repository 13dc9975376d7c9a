#!/usr/bin/env python
"""GPU probe: UMMA SWIZZLE_128B operand with a row-shifted start address."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "object-detection-tensorflow_b200"))
import numpy as np, torch
import subprocess
# the probe kernel is test-only code: built on demand next to the other native test harnesses, never linked into
# the product library
SRC = os.path.join(ROOT, "tests", "native", "tc_probe.cu")
OUT = os.path.join(ROOT, "tests", "native", "_build", "libodt_probe.so")
os.makedirs(os.path.dirname(OUT), exist_ok=True)
if not os.path.exists(OUT) or os.path.getmtime(SRC) > os.path.getmtime(OUT):
    subprocess.check_call(["nvcc", "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC",
                           "-I", os.path.join(ROOT, "object-detection-tensorflow_b200", "csrc"), "-shared", "-o", OUT, SRC,
                           os.path.join(ROOT, "object-detection-tensorflow_b200", "csrc", "api.cu")])
lib = C.CDLL(OUT)
rng = np.random.default_rng(0)
X = rng.standard_normal((136, 64)).astype(np.float16)
W = rng.standard_normal((64, 64)).astype(np.float16)
Xd, Wd = torch.from_numpy(X).cuda(), torch.from_numpy(W).cuda()
for mode in (0, 1):
    out = torch.zeros((3, 128, 64), dtype=torch.float32, device="cuda")
    rc = lib.odt_test_umma_rowoffset(C.c_void_p(Xd.data_ptr()), C.c_void_p(Wd.data_ptr()), C.c_void_p(out.data_ptr()), mode, None)
    torch.cuda.synchronize()
    o = out.cpu().numpy()
    for s in range(3):
        ref = X[s:s + 128].astype(np.float32) @ W.astype(np.float32).T
        err = np.abs(o[s] - ref).max()
        print("mode %d (base_offset=%s) shift %d rows: max|err| %.4g %s" % (mode, "addr>>7&7" if mode else "0", s, err, "OK" if err < 1e-2 else "WRONG"))
