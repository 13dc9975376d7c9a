"""Batch-parallel multi-GPU inference: images shard across ranks (one process per
GPU, torch.distributed over NCCL/NVLink); weights are replicated (seeded init or
broadcast); the only data-path collective is ONE all-gather of the fixed-size
detection records per batch.  The reference itself is
single-device (SSD300.py:458-462)."""
import os

import numpy as np
import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Reads RANK / WORLD_SIZE / LOCAL_RANK / MASTER_* (torchrun)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if torch.cuda.is_available():
        torch.cuda.set_device(local)
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        dist.init_process_group(backend or ("nccl" if torch.cuda.is_available() else "gloo"),
                                rank=rank, world_size=world)
    return rank, world, local


def shard_range(total, rank, world):
    """Contiguous shard: rank r takes images [r*B/G, (r+1)*B/G)."""
    assert total % world == 0, "global batch must divide evenly across ranks"
    per = total // world
    return rank * per, (rank + 1) * per


def gather_records(rec, group=None, out=None):
    """ONE all-gather of the packed per-image records [B, D*6+2] (written by the NMS kernel itself, see
    engine.Tail); returns [world*B, D*6+2] on every rank (`out`: preallocated destination)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return rec
    if out is None:
        out = torch.empty((world * rec.shape[0], rec.shape[1]), dtype=rec.dtype, device=rec.device)
    dist.all_gather_into_tensor(out, rec, group=group)
    return out


def broadcast_weights(weights, src=0):
    """Replicate rank `src`'s variables (name -> ndarray) to every rank."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return weights
    dev = torch.device("cuda") if torch.cuda.is_available() else torch.device("cpu")
    out = {}
    for k in sorted(weights):
        t = torch.from_numpy(np.ascontiguousarray(weights[k], dtype=np.float32)).to(dev)
        dist.broadcast(t, src)
        out[k] = t.cpu().numpy()
    return out


def detect_sharded(model, images_local, consumer=None):
    """Run this rank's image shard and all-gather everybody's detection records.  consumer=None: every
    rank reads back and returns all ranks' detections; consumer=r: only rank r does (the others return
    their own shard's detections)."""
    from .api import _as_host_tensor
    from .engine import unpack_records
    images_local = _as_host_tensor(images_local)
    net = model.engine(images_local.shape[0])
    net.image_buf.copy_(images_local, non_blocking=True)
    net.run()
    rec = gather_records(net.tail.rec)
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world > 1 and consumer is not None and dist.get_rank() != consumer:
        rec = net.tail.rec
    return unpack_records(rec.cpu().numpy(), net.tail.p.cap)
