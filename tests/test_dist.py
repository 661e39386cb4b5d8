"""CPU tests of the multi-GPU host logic (loc_lib_b200/dist.py) on the gloo backend, world_size 2 and 3:
block sharding of scans / hypotheses, the single MIN all-reduce argmin with lowest-index tie-break, the pose
broadcast from the winner's owner, and the end-of-batch pose gather.  The registration itself is faked (a
deterministic score per hypothesis): only the plumbing around liblocreg is under test here."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class FakeReg:
    """Stands in for IcpRegistration.Relocalise: score = squared distance of the hypothesis translation to a target."""

    def __init__(self, target):
        self.target = np.asarray(target, float)

    def Relocalise(self, scan, hyp, want_all=False):
        sc = np.float32(((hyp[:, 4:] - self.target) ** 2).sum(1)).astype(np.float64)
        i = int(np.lexsort((np.arange(len(sc)), sc))[0])
        pose = hyp[i].copy()
        pose[4:] += 0.125  # "refined" pose: the broadcast must carry the owner's result, not the input hypothesis
        return pose, i, float(sc[i]), sc, None


def _worker(rank, world, port, out_dir):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
    import torch.distributed as dist
    from loc_lib_b200 import dist as D
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    rng = np.random.default_rng(42)  # same hypotheses on every rank (replicated inputs)
    hyp = np.zeros((101, 7))
    hyp[:, 3] = 1
    hyp[:, 4:] = rng.uniform(-5, 5, (101, 3))
    hyp[77, 4:] = hyp[13, 4:]  # exact tie across ranks: the lower global index must win
    target = hyp[13, 4:].copy()
    pose, idx, score = D.relocalise_sharded(FakeReg(target), None, hyp, rank, world)
    # batch mapping: every rank registers its block, poses are gathered in rank order
    S = 4 * world + 1  # not divisible: the first rank's block is one scan longer (ragged gather)
    lo, hi = D.shard_range(S, rank, world)
    local = np.arange(lo, hi, dtype=np.float64)[:, None] * np.ones((1, 7))
    allp = D.gather_poses(local)
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), pose=pose, idx=idx, score=score, allp=allp, hyp13=hyp[13])
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_relocalise_and_gather_over_gloo(tmp_path, world):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    outs = [np.load(tmp_path / f"r{r}.npz") for r in range(world)]
    for o in outs:
        assert int(o["idx"]) == 13 and float(o["score"]) == 0.0
        exp = o["hyp13"].copy()
        exp[4:] += 0.125
        assert np.array_equal(o["pose"], exp)  # identical winner pose on every rank
        assert np.array_equal(o["allp"][:, 0], np.arange(4 * world + 1))


def test_shard_range_partitions_exactly():
    from loc_lib_b200 import dist as D
    for n in (0, 1, 7, 4096, 65536, 65537):
        for world in (1, 2, 3, 8):
            spans = [D.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    assert [D.shard_range(4096, r, 8) for r in (0, 7)] == [(0, 512), (3584, 4096)]


def test_c_abi_shard_range_equals_python():
    """locreg_shard_range (pure host code of liblocreg.so: no device needed) is the partition dist.shard_range states."""
    import ctypes as C
    from loc_lib_b200 import _lib, dist as D
    L = _lib.lib()
    for n in (0, 1, 7, 4097, 65536):
        for world in (1, 2, 3, 8):
            for r in range(world):
                lo, hi = C.c_size_t(), C.c_size_t()
                assert L.locreg_shard_range(n, r, world, C.byref(lo), C.byref(hi)) == 0
                assert (lo.value, hi.value) == D.shard_range(n, r, world)
    lo, hi = C.c_size_t(), C.c_size_t()
    assert L.locreg_shard_range(10, 3, 3, C.byref(lo), C.byref(hi)) == -1


def test_single_process_paths_need_no_process_group():
    from loc_lib_b200 import dist as D
    assert D.allreduce_argmin(1.5, 9) == (1.5, 9)
    p = np.arange(7.0)
    assert np.array_equal(D.broadcast_pose(p, 0), p)
    assert np.array_equal(D.gather_poses(np.ones((3, 7))), np.ones((3, 7)))
    rng = np.random.default_rng(1)
    hyp = np.zeros((20, 7))
    hyp[:, 4:] = rng.uniform(-1, 1, (20, 3))
    pose, idx, score = D.relocalise_sharded(FakeReg(hyp[5, 4:]), None, hyp)
    assert idx == 5 and score == 0.0
