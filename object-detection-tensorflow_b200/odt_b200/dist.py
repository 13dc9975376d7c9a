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


def pack_records(dets, det_count):
    """[B,D,6] f32 + [B] i32 -> one [B, D*6+1] f32 record (count rides as a float)."""
    B = dets.shape[0]
    return torch.cat([dets.reshape(B, -1), det_count.reshape(B, 1).to(torch.float32)], dim=1).contiguous()


def gather_records(rec, group=None):
    """ONE all-gather of the packed records; returns [world*B, D*6+1] on every rank."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return rec
    out = torch.empty((world * rec.shape[0], rec.shape[1]), dtype=rec.dtype, device=rec.device)
    dist.all_gather_into_tensor(out, rec, group=group)
    return out


def unpack_records(rec):
    rec = rec.cpu().numpy()
    D = (rec.shape[1] - 1) // 6
    out = []
    for b in range(rec.shape[0]):
        k = int(rec[b, -1])
        d = rec[b, :D * 6].reshape(D, 6)[:k]
        out.append([d[:, 0].copy(), d[:, 1:5].copy(), d[:, 5].astype(np.int32)])
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


def finish_sharded(net):
    """After net.run(): pack this rank's detections, all-gather, read back, unpack."""
    rec = pack_records(net.tail.dets, net.tail.det_count)
    return unpack_records(gather_records(rec))


def detect_sharded(model, images_local):
    """Run this rank's image shard and all-gather everybody's detections."""
    from .api import _as_host_tensor
    images_local = _as_host_tensor(images_local)
    net = model.engine(images_local.shape[0])
    net.image_buf.copy_(images_local, non_blocking=True)
    net.run()
    return finish_sharded(net)
