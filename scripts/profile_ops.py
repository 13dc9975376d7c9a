#!/usr/bin/env python
"""Per-op CUDA-event timing of one forward (not a bench value): which launches
own the step.  usage: python scripts/profile_ops.py [ssd300|ssd512|retinanet|yolov3|fcos] [batch]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "object-detection-tensorflow_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import torch

from helpers import model_cfg


def main():
    kind = sys.argv[1] if len(sys.argv) > 1 else "ssd300"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    import FCOS, RetinaNet, SSD300, SSD512, YOLOv3
    if kind == "ssd300":
        m, hw = SSD300.SSD300(model_cfg("ssd"), None), (300, 300)
    elif kind == "ssd512":
        m, hw = SSD512.SSD512(model_cfg("ssd"), None), (512, 512)
    elif kind == "retinanet":
        m, hw = RetinaNet.RetinaNet(model_cfg("retinanet", data_shape=[800, 800, 3]), None), (800, 800)
    elif kind == "yolov3":
        m, hw = YOLOv3.YOLOv3(model_cfg("yolov3", data_shape=[416, 416, 3]), None), (416, 416)
    else:
        m, hw = FCOS.FCOS(model_cfg("fcos", data_shape=[1024, 1024, 3]), None), (1024, 1024)
    net = m.engine(B, graph=False)
    img = np.random.default_rng(0).integers(0, 256, (B,) + hw + (3,)).astype(np.float32)
    net.image_buf.copy_(torch.from_numpy(img))
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(2):
        net.forward()
    torch.cuda.synchronize()
    reps, rec = 3, {}
    for _ in range(reps):
        if getattr(net, "gn_arena", None) is not None:
            net.gn_arena.zero_()  # per-op launches bypass Net.forward, which zeroes the GroupNorm accumulators
        evs = []
        for i, op in enumerate(net.ops):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); op.launch(net, st); b.record()
            evs.append((i, a, b))
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); net.tail.launch(net, st); b.record()
        evs.append((-1, a, b))
        torch.cuda.synchronize()
        for i, a, b in evs:
            rec.setdefault(i, []).append(a.elapsed_time(b))
    total = 0.0
    print("%-3s %-10s %-52s %9s %9s %8s" % ("#", "kind", "name", "ms", "GFLOP", "TFLOP/s"))
    for i, op in enumerate(net.ops):
        ms = float(np.median(rec[i]))
        total += ms
        fl = getattr(op, "flops", 0)
        kindn = type(op).__name__ + ("/tc" if getattr(op, "use_tc", False) else "")
        name = getattr(op, "kernel", None) or getattr(getattr(op, "y", None), "name", "")
        geo = ""
        if hasattr(op, "p"):
            p = op.p
            geo = " [%dx%d %d->%d k%d s%d]" % (p.H, p.W, p.Cin, p.Cout, p.R, p.stride)
        print("%-3d %-10s %-52s %9.4f %9.1f %8.1f" % (i, kindn[:10], (name + geo)[-52:], ms, fl / 1e9,
                                                     fl / ms / 1e9 if ms > 0 else 0))
    ms = float(np.median(rec[-1]))
    total += ms
    print("tail (decode+nms) %.4f ms ; sum of ops %.3f ms ; conv GFLOP %.1f" % (ms, total, net.conv_flops / 1e9))
    # split the tail
    import ctypes as C
    from odt_b200 import lib as L
    t = net.tail
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    ev[0].record()
    L.check(net.lib.odt_decode_candidates(net.head_buf.data_ptr(), C.byref(t.p), net.batch, t.cand_keys.data_ptr(),
                                          t.cand_count.data_ptr(), st))
    ev[1].record()
    L.check(net.lib.odt_nms_per_class(net.head_buf.data_ptr(), C.byref(t.p), net.batch, t.cand_keys.data_ptr(),
                                      t.cand_count.data_ptr(), t.dets.data_ptr(), t.det_anchor.data_ptr(),
                                      t.det_count.data_ptr(), t.scratch.data_ptr(), t.work.data_ptr(),
                                      t.status.data_ptr(), t.box_pool.data_ptr() if t.box_pool is not None else None,
                                      t.pool_entries, t.rec.shape[1], st))
    ev[2].record()
    torch.cuda.synchronize()
    # whole forward as ONE CUDA graph replay (what the bench / API actually runs)
    net.capture()
    for _ in range(3):
        net.run()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    g0.record()
    for _ in range(10):
        net.run()
    g1.record()
    torch.cuda.synchronize()
    gms = g0.elapsed_time(g1) / 10
    print("CUDA-graph replay: %.3f ms/step -> %.0f images/s (batch %d), %d launches, conv %.1f TFLOP/s whole-step"
          % (gms, B / gms * 1e3, B, net.num_launches(), net.conv_flops / gms / 1e9))
    cc = t.cand_count.cpu().numpy()
    print("decode %.4f ms (%.1f MB rows -> %.0f GB/s) ; nms %.4f ms ; candidates/class max %d mean %.1f ; dets/img mean %.1f"
          % (ev[0].elapsed_time(ev[1]), net.head_buf.numel() * 4 / 1e6,
             net.head_buf.numel() * 4 / 1e6 / ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), cc.max(), cc.mean(),
             t.det_count.float().mean().item()))


if __name__ == "__main__":
    main()
