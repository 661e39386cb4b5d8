"""The parallel formulation of the incremental-NDT voxel cache (loc_lib_b200/csrc/inc_ndt.cu) against the literal
sequential LRU of the reference (NdtRegistration::SetIncNdtTargetCloud, ndt_registration.cpp:150-183, restated in
oracle/oracle.cpp IncNdt::AddCloud).

The device cannot walk a std::list point by point.  It uses the stack property of LRU instead: with C = capacity_ - 1
entries retained after every insertion,
  * an access is a HIT iff the key was accessed before and fewer than C distinct other keys were accessed since;
  * the cache content is the C most recently accessed distinct keys;
  * the points a voxel is re-estimated from are those it received since its last MISS inside the cloud (a re-inserted
    voxel starts empty), or all of its points of the cloud when every access was a hit.
`lru_model_add_cloud` below is the numpy statement of exactly the steps the kernels take (runs, groups, previous
occurrence, reuse-distance count, survivors by last access); this test pins it to the oracle on adversarial sequences
(tiny capacities, voxels evicted and re-inserted within one cloud)."""
import numpy as np
import pytest

import oracle_py as O

NEG = -(1 << 40)


def lru_model_add_cloud(order, keys, capacity):
    """order: list of keys, oldest first (the cache before the cloud).  keys: voxel key per valid point, in cloud order.
    Returns (new order, {key: points used for the re-estimate}) for the keys present afterwards that the cloud touched."""
    C = capacity - 1
    m = len(order)
    rank = {k: r for r, k in enumerate(order)}
    # runs of consecutive points with the same key
    heads = [i for i in range(len(keys)) if i == 0 or keys[i] != keys[i - 1]]
    run_key = [keys[i] for i in heads]
    run_len = [(heads[j + 1] if j + 1 < len(heads) else len(keys)) - heads[j] for j in range(len(heads))]
    nr = len(heads)
    # previous occurrence on the unified time line: old entry of rank r sits at time r - m, run j at time j
    prev = [NEG] * nr
    last_seen = {}
    for j, k in enumerate(run_key):
        if k in last_seen:
            prev[j] = last_seen[k]
        elif k in rank:
            prev[j] = rank[k] - m
        last_seen[k] = j
    prev_arr = np.array(prev, dtype=np.int64)
    miss = np.zeros(nr, bool)
    for j in range(nr):
        i = prev[j]
        if i == NEG:
            miss[j] = True
            continue
        if j - i - 1 < C:
            continue  # fewer than C accesses in between: a hit whatever they were
        distinct = (-1 - i) if i < 0 else 0  # the old entries newer than this one
        lo = max(0, i + 1)
        distinct += int(np.count_nonzero(prev_arr[lo:j] < i))  # first occurrences inside (i, j)
        miss[j] = distinct >= C
    groups = {}
    for j, k in enumerate(run_key):
        groups.setdefault(k, []).append(j)
    touched = set(groups)
    total = (m - sum(1 for k in order if k in touched)) + len(groups)
    E = max(0, total - C)
    untouched = [k for k in order if k not in touched]
    e_old = min(E, len(untouched))
    e_grp = E - e_old
    by_last = sorted(groups, key=lambda k: groups[k][-1])
    new_order = untouched[e_old:] + by_last[e_grp:]
    used = {}
    for k in by_last[e_grp:]:
        js = groups[k]
        last_miss = max([j for j in js if miss[j]], default=-1)
        used[k] = sum(run_len[j] for j in js if j >= last_miss)
    return new_order, used


def keys_to_cloud(keys):
    """One point in the middle of voxel (kx, 0, 0) per key (voxel size 1)."""
    pts = np.zeros((len(keys), 4), np.float32)
    pts[:, 0] = np.asarray(keys, np.float32) + 0.5
    pts[:, 1] = 0.5
    pts[:, 2] = 0.5
    return pts


@pytest.mark.parametrize("capacity,n_keys,n_pts,seed", [(4, 6, 60, 0), (9, 12, 400, 1), (33, 40, 3000, 2), (33, 200, 3000, 3),
                                                         (120, 150, 5000, 4), (2, 5, 50, 5), (50, 30, 2000, 6)])
def test_lru_model_equals_sequential_lru(capacity, n_keys, n_pts, seed):
    rng = np.random.default_rng(seed)
    ref = O.OracleIncNdt(capacity=capacity, skip_nonfinite=1, voxel_size=1.0)
    order, counts = [], {}
    for cloud in range(6):
        # bursts (runs) of the same key, a drifting window of active keys: hits, evictions and re-insertions all occur
        base = cloud * (n_keys // 3)
        ks = []
        while len(ks) < n_pts:
            k = int(base + rng.integers(0, n_keys))
            ks += [k] * int(rng.integers(1, 4))
        ks = ks[:n_pts]
        ref.set_target(keys_to_cloud(ks))
        order, used = lru_model_add_cloud(order, ks, capacity)
        counts.update(used)
        counts = {k: v for k, v in counts.items() if k in set(order)}
        rk, _, _, rn = ref.voxels()
        assert sorted(order) == [int(k) for k in rk[:, 0]], cloud
        assert [counts[int(k)] for k in rk[:, 0]] == [int(c) for c in rn], cloud
        assert len(order) <= capacity - 1
