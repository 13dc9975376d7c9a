#!/usr/bin/env python
"""torchrun --nproc-per-node N scripts/check_sharded.py: the sharded multi-GPU
result must equal the single-GPU result image for image (weak-scaling shards,
one NCCL all-gather of detection records)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "object-detection-tensorflow_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import torch

from helpers import model_cfg
from odt_b200 import dist as od


def main():
    rank, world, local = od.init_from_env()
    import SSD300
    m = SSD300.SSD300(model_cfg("ssd", nms_score_threshold=0.3), None)
    per = 4
    img = np.random.default_rng(0).integers(0, 256, (per * world, 300, 300, 3)).astype(np.float32)
    lo, hi = od.shard_range(per * world, rank, world)
    out = m.detect_batch_sharded(img[lo:hi])
    assert len(out) == per * world
    # every rank recomputes a few images alone and compares
    for g in range(per * world):
        if g % world != rank:
            continue
        ref = m.detect_batch(img[g:g + 1])[0]
        for a, b in zip(out[g], ref):
            np.testing.assert_array_equal(a, b)
    torch.distributed.barrier()
    if rank == 0:
        print("sharded == single-GPU for %d images on %d ranks: OK" % (per * world, world))
    torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
