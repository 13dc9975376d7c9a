#!/usr/bin/env python
"""torchrun --nproc-per-node N scripts/check_sharded.py: the sharded multi-GPU
results must equal the single-GPU results image for image (weak-scaling shards,
one NCCL all-gather of the packed detection records) -- the per-batch call, the
stream with every rank reading back, and the stream with a consumer rank."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "object-detection-tensorflow_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import torch

from helpers import model_cfg
from odt_b200 import dist as od


def same(a, b):
    assert len(a) == len(b) == 3
    for x, y in zip(a, b):
        np.testing.assert_array_equal(x, y)


def main():
    rank, world, local = od.init_from_env()
    import SSD300
    m = SSD300.SSD300(model_cfg("ssd", nms_score_threshold=0.3), None)
    per, steps = 4, 3
    rng = np.random.default_rng(0)
    imgs = [rng.integers(0, 256, (per * world, 300, 300, 3)).astype(np.float32) for _ in range(steps)]
    lo, hi = od.shard_range(per * world, rank, world)
    # single-GPU truth for every image of every step (each rank computes all of it: small)
    truth = [[m.detect_batch(im[g:g + 1])[0] for g in range(per * world)] for im in imgs]
    # (1) per-batch call, every rank gets everything
    out = m.detect_batch_sharded(imgs[0][lo:hi])
    assert len(out) == per * world
    for g in range(per * world):
        same(out[g], truth[0][g])
    # (2) stream, every rank reads back the gathered records
    shards = [torch.from_numpy(im[lo:hi].copy()).pin_memory() for im in imgs]
    got = list(m.detect_stream_sharded(shards))
    assert len(got) == steps
    for s in range(steps):
        assert len(got[s]) == per * world
        for g in range(per * world):
            same(got[s][g], truth[s][g])
    # (3) stream with consumer rank 0: rank 0 yields all images, the others their own shard
    got = list(m.detect_stream_sharded(shards, consumer=0))
    for s in range(steps):
        if rank == 0:
            assert len(got[s]) == per * world
            for g in range(per * world):
                same(got[s][g], truth[s][g])
        else:
            assert len(got[s]) == per
            for i, g in enumerate(range(lo, hi)):
                same(got[s][i], truth[s][g])
    torch.distributed.barrier()
    if rank == 0:
        print("sharded == single-GPU for %d images x %d steps on %d ranks (batch call, stream, consumer stream): OK"
              % (per * world, steps, world))
    torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
