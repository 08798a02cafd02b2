"""Sliding-window inference host logic on CPU (lib/funcs.py:4-51, BASELINE config 5): the window list, the contiguous
z-sharding of the windows over ranks (no collective inside the loop, one sum of the accumulators at
the end) and the whole predict_tumor_inwindow call driven through the plain-PyTorch backend -- single process and
2 gloo ranks -- against the oracle's restatement of the reference loop."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_window_starts_match_the_reference_loop():
    from h_denseunet_b200.inference import window_starts
    from oracle import hdense_oracle as orc
    for z, lo, hi, cols in [(20, 4, 15, 8), (512, 0, 511, 8), (64, 30, 40, 8), (9, 0, 8, 8), (100, 90, 99, 8), (48, 3, 20, 12)]:
        assert window_starts(z, lo, hi, cols) == orc.window_starts(z, lo, hi, cols)
    # BASELINE config 5: 512 slices, full liver mask -> windows range(0, 506, 2) = 253
    assert window_starts(512, 0, 511, 8) == list(range(0, 506, 2))


def test_shard_windows_tile_the_list_contiguously():
    from h_denseunet_b200.inference import shard_windows, window_starts
    starts = window_starts(512, 0, 511, 8)
    for world in (1, 2, 3, 4, 8, 300):
        parts = [shard_windows(starts, world, r) for r in range(world)]
        assert sum(parts, []) == starts                                  # every window exactly once, z order kept
        sizes = [len(p) for p in parts if p]
        assert max(sizes) - min(sizes) <= (len(starts) + world - 1) // world
        for a, b in zip(parts, parts[1:]):
            if a and b:
                assert a[-1] < b[0]
    # neighbouring ranges overlap in at most cols - 2 - step output slices (8-slice windows, stride 2: 4 slices)
    a, b = shard_windows(starts, 2, 0), shard_windows(starts, 2, 1)
    out_a = set(range(a[0] + 1, a[-1] + 7))
    out_b = set(range(b[0] + 1, b[-1] + 7))
    assert len(out_a & out_b) == 4


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _setup_paths():
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)


def _model_and_volume():
    import h_denseunet_b200 as hdn
    from torch_backend import TorchBackend
    from util import Args, perturb_params
    a = Args(b=1, input_size=32, input_cols=8)
    m = hdn.dense_rnn_net(a, backend=TorchBackend(), device="cpu", precision="fp32", seed=0)
    perturb_params(m)
    rng = np.random.default_rng(11)
    vol = rng.normal(0, 60, (32, 32, 20)).astype(np.float32)
    return m, a, vol, np.array([0, 0, 4]), np.array([31, 31, 15])


def _oracle_result(m, vol, mini, maxi):
    from oracle import hdense_oracle as orc
    w = m.get_weights_dict()

    def pred(box):
        return orc.forward_hybrid(w, box, training=False)[1].numpy()

    return orc.predict_tumor_inwindow(pred, vol, 3, mini, maxi, 32, 8)


def test_predict_tumor_inwindow_single_process_matches_oracle():
    _setup_paths()
    import h_denseunet_b200 as hdn
    torch.set_num_threads(4)
    m, a, vol, mini, maxi = _model_and_volume()
    s1, s2 = hdn.predict_tumor_inwindow(m, vol, 3, mini, maxi, a)
    o1, o2 = _oracle_result(m, vol, mini, maxi)
    assert np.abs(s1 - o1).max() < 1e-4 and np.abs(s2 - o2).max() < 1e-4


def _worker(rank, world, port, out):
    _setup_paths()
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import h_denseunet_b200 as hdn
    from h_denseunet_b200.inference import shard_windows, window_starts
    m, a, vol, mini, maxi = _model_and_volume()
    mine = shard_windows(window_starts(20, 4, 15, 8), world, rank)
    s1, s2 = hdn.predict_tumor_inwindow(m, vol, 3, mini, maxi, a)
    r1, r2 = hdn.predict_tumor_inwindow(m, vol, 3, mini, maxi, a, reuse_2d=True)     # sharding + 2-D slice reuse together
    res = {"n_windows": len(mine), "s1": s1, "s2": s2, "reuse_equal": bool(np.array_equal(s1, r1) and np.array_equal(s2, r2))}
    if rank == 0:
        o1, o2 = _oracle_result(m, vol, mini, maxi)
        res["e1"], res["e2"] = float(np.abs(s1 - o1).max()), float(np.abs(s2 - o2).max())
    out[rank] = res
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_predict_tumor_inwindow_two_ranks_gloo():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert len(out) == world
    from h_denseunet_b200.inference import window_starts
    assert out[0]["n_windows"] + out[1]["n_windows"] == len(window_starts(20, 4, 15, 8))
    assert out[0]["n_windows"] > 0 and out[1]["n_windows"] > 0
    assert out[0]["e1"] < 1e-4 and out[0]["e2"] < 1e-4                      # sharded result == the reference loop
    assert np.array_equal(out[0]["s1"], out[1]["s1"]) and np.array_equal(out[0]["s2"], out[1]["s2"])
    assert out[0]["reuse_equal"] and out[1]["reuse_equal"]


def test_new_slices_of_a_moved_window():
    from h_denseunet_b200.inference import new_slices
    assert new_slices(8, 2) == [0, 5, 6, 7]          # stride 2: two new interior slices + the two edges
    assert new_slices(8, 1) == [0, 6, 7]             # clamped tail window one slice further
    assert new_slices(12, 3) == [0, 8, 9, 10, 11]


def test_slice_reuse_is_bit_identical_and_halves_the_2d_work():
    """with the 2-D per-slice results reused across the 75 %-overlapping windows the output is
    unchanged (inference-mode BN: slices are independent) and a window after the first costs step + 2 = 4 slice
    evaluations of the 2-D network instead of 8."""
    _setup_paths()
    import h_denseunet_b200 as hdn
    from h_denseunet_b200.inference import window_starts
    torch.set_num_threads(4)
    m, a, vol, mini, maxi = _model_and_volume()
    vol = np.concatenate([vol, vol[:, :, :3]], axis=2)            # z = 23: the tail window is clamped (odd step of 1)
    st0, st1 = {}, {}
    s1, s2 = hdn.predict_tumor_inwindow(m, vol, 3, mini, maxi, a, reuse_2d=False, stats=st0)
    r1, r2 = hdn.predict_tumor_inwindow(m, vol, 3, mini, maxi, a, reuse_2d=True, stats=st1)
    assert np.array_equal(s1, r1) and np.array_equal(s2, r2)
    starts = window_starts(23, 4, 15, 8)
    steps = [b - a_ for a_, b in zip(starts, starts[1:])]
    assert st0["windows"] == st1["windows"] == len(starts)
    assert st0["slices_2d"] == 8 * len(starts)
    assert st1["slices_2d"] == 8 + sum(d + 2 for d in steps)
    assert st1["slices_2d"] <= 0.6 * st0["slices_2d"]
    o1, o2 = _oracle_result(m, vol, mini, maxi)
    assert np.abs(r1 - o1).max() < 1e-4 and np.abs(r2 - o2).max() < 1e-4
