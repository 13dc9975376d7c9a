#!/usr/bin/env python
"""Per-role timeline of CTA 0 of one tensor-core conv (debug build, -DODT_TC_TIMELINE): where a tile's
latency goes (producer start -> MMA issue -> commit -> accumulator complete -> hand-back).
usage (on a GPU box): tc_timeline.py B H W Cin Cout k stride [pool] [in_halo] [out_halo]"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
PKG = os.path.join(ROOT, "object-detection-tensorflow_b200")
sys.path.insert(0, ROOT)
sys.path.insert(0, PKG)
import __graft_entry__ as ge  # noqa: E402


def build_debug_lib():
    ge.build()
    csrc = os.path.join(PKG, "csrc")
    obj = os.path.join(csrc, "conv_tc_tl.o")
    out = os.path.join(PKG, "odt_b200", "libodt_b200_tl.so")
    subprocess.check_call(["nvcc"] + ge.NVCC_FLAGS + ["-DODT_TC_TIMELINE", "-c", "conv_tc.cu", "-o", obj], cwd=csrc)
    objs = [os.path.join(csrc, s[:-3] + ".o") for s in ge.SRCS if s != "conv_tc.cu"] + [obj]
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", out] + objs, cwd=csrc)
    return out


def main():
    from odt_b200 import lib as L
    L.LIB_PATH = build_debug_lib()
    lib = L.load()
    lib.odt_debug_tc_timeline.restype = C.c_int
    lib.odt_debug_tc_timeline.argtypes = [C.c_void_p, C.c_int]
    cap = 1 << 16
    buf = torch.zeros(2 * cap, dtype=torch.int64, device="cuda")
    L.check(lib.odt_debug_tc_timeline(buf.data_ptr(), cap), "timeline")
    args = sys.argv[1:11]
    args += ["0", "1", "1"][len(args) - 7:] if len(args) < 10 else []
    sys.argv = [sys.argv[0]] + args[:10] + ["1"]  # reps = 1 after conv_micro's 3 warm-ups
    import conv_micro
    conv_micro.main()
    torch.cuda.synchronize()
    raw = buf.cpu().numpy().view(np.uint64).reshape(-1, 2)
    raw = raw[raw[:, 0] != 0]
    role = (raw[:, 1] >> np.uint64(56)).astype(int)
    event = ((raw[:, 1] >> np.uint64(48)) & np.uint64(0xFF)).astype(int)
    tile = (raw[:, 1] & np.uint64(0xFFFFFFFF)).astype(np.int64)
    t0 = raw[:, 0].min()
    names = {(0, 0): "producer: tile start", (1, 0): "mma: accumulator free", (1, 1): "mma: first stage issued",
             (1, 2): "mma: last stage committed", (2, 0): "epi0: accumulator complete", (2, 1): "epi0: handed back",
             (3, 0): "epi1: accumulator complete", (3, 1): "epi1: handed back"}
    order = np.argsort(raw[:, 0], kind="stable")
    # the last launch only (the buffer accumulates over the warm-up launches): keep the final 1/4 of entries
    order = order[-len(order) // 4:]
    print("%12s  %-30s %s" % ("clk", "event", "tile"))
    for i in order[:400]:
        print("%12d  %-30s %d" % (int(raw[i, 0] - t0), names.get((role[i], event[i]), "%d/%d" % (role[i], event[i])), tile[i]))


if __name__ == "__main__":
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    main()
