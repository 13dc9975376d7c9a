#!/usr/bin/env python
"""decode / NMS kernels in isolation on one model's real head rows (warm, CUDA-event timed, 50 reps each).
usage: tail_micro.py model batch"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _odt_path  # noqa: F401,E402
from helpers import model_cfg  # noqa: E402
from odt_b200 import lib as L  # noqa: E402


def build_model(kind, B):
    import FCOS
    import RetinaNet
    import SSD300
    import YOLOv3
    if kind == "ssd300":
        m, hw = SSD300.SSD300(model_cfg("ssd"), None), (300, 300)
    elif kind == "retinanet":
        m, hw = RetinaNet.RetinaNet(model_cfg("retinanet", data_shape=[800, 800, 3]), None), (800, 800)
    elif kind == "yolov3":
        m, hw = YOLOv3.YOLOv3(model_cfg("yolov3", data_shape=[416, 416, 3]), None), (416, 416)
    else:
        m, hw = FCOS.FCOS(model_cfg("fcos", data_shape=[1024, 1024, 3]), None), (1024, 1024)
    net = m.engine(B, graph=False)
    img = np.random.default_rng(0).integers(0, 256, (B,) + hw + (3,)).astype(np.float32)
    net.image_buf.copy_(torch.from_numpy(img))
    return m, net


def main():
    name, B = sys.argv[1], int(sys.argv[2])
    m, net = build_model(name, B)
    net.forward()
    torch.cuda.synchronize()
    t = net.tail
    lib = net.lib

    def decode():
        st = torch.cuda.current_stream().cuda_stream  # the capturing stream inside torch.cuda.graph
        L.check(lib.odt_decode_candidates(net.head_buf.data_ptr(), C.byref(t.p), net.batch, t.cand_keys.data_ptr(),
                                          t.cand_count.data_ptr(), st))

    def nms():
        st = torch.cuda.current_stream().cuda_stream
        L.check(lib.odt_nms_per_class(net.head_buf.data_ptr(), C.byref(t.p), net.batch, t.cand_keys.data_ptr(),
                                      t.cand_count.data_ptr(), t.dets.data_ptr(), t.det_anchor.data_ptr(),
                                      t.det_count.data_ptr(), t.scratch.data_ptr(), t.work.data_ptr(),
                                      t.status.data_ptr(), t.box_pool.data_ptr() if t.box_pool is not None else None,
                                      t.pool_entries, t.rec.shape[1], st))

    def timeit(fn, graph=True, reps=50):
        fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(reps):
                fn()
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    mb = net.head_buf.numel() * 4 / 1e6
    d = timeit(decode)
    decode()
    n = timeit(lambda: (decode(), nms()))
    print("%s B=%d: rows %.1f MB ; decode (memset + kernel) %.2f us -> %.0f GB/s ; decode+nms %.2f us" % (
        name, B, mb, d * 1e3, mb / d, n * 1e3))


if __name__ == "__main__":
    main()
