"""Multi-GPU plumbing for the two workloads that shard: one process per GPU, torch.distributed.

* batch offline mapping: scans are block-partitioned over ranks, every rank holds a replica of the map, and there
  is NO collective on the data path (only the gather of the S x 7 poses at the end);
* global relocalisation: pose hypotheses are dealt out round-robin, each rank finds its local best, and ONE
  min-all-reduce of a packed (float32 score bits << 32 | global hypothesis index) key picks the winner
  (lowest index wins ties), followed by a broadcast of the winning pose from its owner.
torch is used for process-group plumbing only; all registration work goes through liblocreg.so.

The product path keeps the exchange step INSIDE liblocreg.so (locreg_comm_init / locreg_relocalise_sharded /
locreg_align_batch_sharded: NCCL on the handle's stream, usable from a C++ host without Python): `comm_init` below only
carries the 128-byte NCCL id from rank 0 to the others through the process group torchrun set up, and
`relocalise_sharded` / `align_batch_sharded` call the C ABI when the handle has a communicator.  The torch-collective
forms remain for backends NCCL cannot serve (gloo on CPU: the host-logic tests of tests/test_dist.py).
"""
import struct

import numpy as np


def shard_range(n_items, rank, world):
    """Block partition [lo, hi) of n_items over world ranks (remainder spread over the first ranks)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pack_score(score, index):
    """(float32 bits of score) << 32 | index, as a Python int; NaN / negative scores map to +inf."""
    f = np.float32(score)
    if not (f == f) or f < 0:
        f = np.float32(np.inf)
    bits = struct.unpack("<I", struct.pack("<f", float(f)))[0]
    return (bits << 32) | (int(index) & 0xFFFFFFFF)


def unpack_score(key):
    bits = (key >> 32) & 0xFFFFFFFF
    return struct.unpack("<f", struct.pack("<I", bits))[0], key & 0xFFFFFFFF


def allreduce_argmin(local_score, local_global_index, device=None):
    """Global (score, index) minimum over all ranks with one MIN all-reduce.

    float32 score bits of a non-negative float are < 2^31, so the packed key fits a signed int64 and integer order
    equals (score, index) order."""
    import torch
    import torch.distributed as dist
    key = pack_score(local_score, local_global_index)
    t = torch.tensor([key], dtype=torch.int64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return unpack_score(int(t.item()))


def broadcast_pose(pose7, owner_rank, device=None):
    import torch
    import torch.distributed as dist
    t = torch.as_tensor(np.asarray(pose7, np.float64), device=device).clone()
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(t, src=owner_rank)
    return t.cpu().numpy()


def comm_init(reg, device=None):
    """Gives `reg` (an IcpRegistration / NdtRegistration of this rank) an NCCL communicator over all ranks of the
    initialised torch process group: rank 0 draws the id (locreg_comm_unique_id), the process group carries it."""
    import torch
    import torch.distributed as dist
    if not (dist.is_initialized() and dist.get_world_size() > 1):
        return False
    rank, world = dist.get_rank(), dist.get_world_size()
    t = torch.zeros(128, dtype=torch.uint8, device=device)
    if rank == 0:
        t = torch.frombuffer(bytearray(reg.CommUniqueId()), dtype=torch.uint8).to(device) if device is not None else \
            torch.frombuffer(bytearray(reg.CommUniqueId()), dtype=torch.uint8).clone()
    dist.broadcast(t, src=0)
    reg.CommInit(bytes(t.cpu().numpy().tobytes()), rank, world)
    return True


def align_batch_sharded(reg, clouds, offsets, predict_poses, S_global):
    """Batch mapping over the ranks of reg's communicator: this rank's block in, ALL poses out (C ABI, NCCL inside)."""
    return reg.ScanMatchBatchSharded(clouds, offsets, predict_poses, S_global)


def relocalise_sharded(reg, scan, hypotheses, rank=0, world=1, device=None):
    """Global relocalisation over `world` ranks; `reg` is this rank's IcpRegistration with the map set.

    Hypotheses are dealt out round-robin (rank r registers hypotheses r, r + world, ...): neighbouring hypotheses of
    the search grid cost about the same (the far-off ones several times more than those near the truth), so a
    strided split balances the ranks where a block split would not.  Returns (best_pose, best_global_index,
    best_score), identical on every rank."""
    hyp = np.ascontiguousarray(hypotheses, np.float64).reshape(-1, 7)
    if hasattr(reg, "CommInfo") and reg.CommInfo()[1] == world:
        return reg.RelocaliseSharded(scan, hyp)  # the product path: key, all-reduce and broadcast on the handle's stream
    mine = hyp[rank::world]
    if len(mine):
        pose, idx, score, _, _ = reg.Relocalise(scan, mine)
        gidx = rank + idx * world
    else:
        pose, gidx, score = np.zeros(7), 0xFFFFFFFF, np.inf
    best_score, best_idx = allreduce_argmin(score, gidx, device)
    owner = best_idx % world if best_idx != 0xFFFFFFFF else 0
    best_pose = broadcast_pose(pose if owner == rank else np.zeros(7), owner, device)
    return best_pose, int(best_idx), float(best_score)


def gather_poses(local_poses, device=None):
    """All-gather of per-rank pose blocks -> (sum of block sizes, 7), in rank order.  Blocks may differ in size
    (shard_range gives the first S % world ranks one scan more): the sizes are exchanged first, the blocks are padded
    to the largest and trimmed after the gather."""
    import torch
    import torch.distributed as dist
    t = torch.as_tensor(np.ascontiguousarray(local_poses, np.float64).reshape(-1, 7), device=device)
    if not (dist.is_initialized() and dist.get_world_size() > 1):
        return t.cpu().numpy()
    world = dist.get_world_size()
    n = torch.tensor([t.shape[0]], dtype=torch.int64, device=device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n)
    sizes = [int(x.item()) for x in sizes]
    pad = torch.zeros((max(sizes), 7), dtype=torch.float64, device=device)
    pad[:t.shape[0]] = t
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad)
    return torch.cat([o[:k] for o, k in zip(out, sizes)]).cpu().numpy()
