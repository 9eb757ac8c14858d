"""Work list of the split-K tail experiment (reazonspeech_b200/csrc/kernels.h: splitk_plan / splitk_item), enumerated on the
host through rs_debug_splitk_schedule.  The device kernel (gemm_bf16_tn_2cta_sk_kernel) walks exactly this list in its three
roles; what must hold for it to be correct and deadlock-free:

* every (tile, k-step) is computed exactly once;
* a split tile has exactly one owner (part 0) and S - 1 contributors with adjacent k ranges;
* contributors never wait, and every contributor of a tile sits at a position no later than its owner's position in
  the per-cluster lists (an owner only ever waits for work that is already running or done);
* no cluster gets more than one item beyond the even share."""
import ctypes as C

import numpy as np
import pytest

from reazonspeech_b200 import engine as E


def schedule(num_tiles, ncl, num_k):
    lib = E.load_library()
    split = C.c_int(0)
    n = lib.rs_debug_splitk_schedule(num_tiles, ncl, num_k, None, 0, C.byref(split))
    rows = np.zeros((n, 7), np.int32)
    assert lib.rs_debug_splitk_schedule(num_tiles, ncl, num_k, rows.ctypes.data_as(C.c_void_p), n, C.byref(split)) == n
    return rows, split.value


@pytest.mark.parametrize("num_tiles,ncl,num_k", [
    (196, 74, 64),     # FFN W2 at 32 clips: 49 x 4 tiles, K = 4096 -> 2.65 waves
    (784, 74, 16),     # FFN W1: 49 x 16 tiles, K = 1024
    (776, 74, 64),     # 128 clips
    (148, 74, 64),     # exact waves: nothing to split
    (75, 74, 40), (73, 74, 64), (1, 74, 64), (150, 74, 9), (223, 74, 33), (37, 74, 128),
])
def test_splitk_schedule_covers_every_k_step_once(num_tiles, ncl, num_k):
    rows, S = schedule(num_tiles, ncl, num_k)
    ncl_eff = ncl
    cover = np.zeros((num_tiles, num_k), np.int32)
    for cid, it, tile, k0, k1, kind, part in rows:
        assert 0 <= tile < num_tiles and 0 <= k0 < k1 <= num_k
        cover[tile, k0:k1] += 1
    assert (cover == 1).all()
    tail = num_tiles % ncl_eff
    if tail == 0:
        assert S == 1
    assert (rows[:, 5] == 0).sum() == (num_tiles - tail if S > 1 else num_tiles)
    # positions within a cluster are consecutive from 0 (a cluster may have nothing to do when tiles * S < clusters)
    for cid in range(ncl_eff):
        its = rows[rows[:, 0] == cid][:, 1]
        assert list(its) == list(range(len(its)))
    if S > 1:
        first_tail = num_tiles - tail
        for tile in range(first_tail, num_tiles):
            parts = rows[rows[:, 2] == tile]
            assert sorted(parts[:, 6]) == list(range(S))
            owner = parts[parts[:, 5] == 2]
            assert len(owner) == 1 and owner[0, 6] == 0 and owner[0, 3] == 0
            contrib = parts[parts[:, 5] == 1]
            assert len(contrib) == S - 1
            assert (contrib[:, 1] <= owner[0, 1]).all(), "a contributor is scheduled after its owner"
            assert (contrib[:, 0] != owner[0, 0]).all() or (contrib[contrib[:, 0] == owner[0, 0]][:, 1] < owner[0, 1]).all()
            ks = sorted((int(k0), int(k1)) for k0, k1 in parts[:, 3:5])
            assert ks[0][0] == 0 and ks[-1][1] == num_k and all(a[1] == b[0] for a, b in zip(ks, ks[1:]))
        # the tail costs ceil(tail * S / ncl) / S tile times instead of 1
        per_cluster = np.bincount(rows[rows[:, 5] != 0][:, 0], minlength=ncl_eff)
        assert per_cluster.max() == -(-tail * S // ncl_eff)
        assert -(-tail * S // ncl_eff) / S < 1.0


def test_splitk_choice_for_the_ffn_shape():
    _, S = schedule(196, 74, 64)
    assert S == 3          # 48 tail tiles x 3 parts = 144 items on 74 clusters: 2 rounds of a third -> 2.67 waves instead of 3


def test_splitk_schedule_random_shapes():
    """Coverage and ordering invariants over random (tiles, clusters, k-steps): every k-step once, contributors scheduled
    no later than their owner, never on a position behind it in the same cluster."""
    hyp = pytest.importorskip("hypothesis")
    st = pytest.importorskip("hypothesis.strategies")

    @hyp.settings(max_examples=150, deadline=None)
    @hyp.given(st.integers(1, 900), st.integers(1, 80), st.integers(1, 130))
    def check(num_tiles, ncl, num_k):
        rows, S = schedule(num_tiles, ncl, num_k)
        cover = np.zeros((num_tiles, num_k), np.int32)
        for cid, it, tile, k0, k1, kind, part in rows:
            assert 0 <= cid < ncl and 0 <= k0 < k1 <= num_k
            cover[tile, k0:k1] += 1
        assert (cover == 1).all()
        assert 1 <= S <= 4 and (S == 1 or S * 8 <= num_k)
        owners = {int(r[2]): r for r in rows if r[5] == 2}
        for r in rows[rows[:, 5] == 1]:
            o = owners[int(r[2])]
            assert r[1] <= o[1] and (r[0] != o[0] or r[1] < o[1])
        if S == 1:
            assert (rows[:, 5] == 0).all()

    check()
