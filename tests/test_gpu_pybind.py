"""The drop-in boundary from C++: simulator._bgs_pybind (bindings/bgs_pybind.cpp, pybind11 over
include/bgs_b200.h -- the counterpart of the reference's nanobind modules, connect.cpp:19-62) against the oracle."""
import numpy as np
import pytest
import torch

from conftest import DEFAULT_BOUNCE_GRID

pytestmark = pytest.mark.gpu


def test_cpp_extension_host_rollouts_equal_oracle(oracle):
    from simulator import _bgs_pybind as m

    assert m.version() == 100 and m.device_count() >= 1
    for cfg in ((6, 7, 4), (8, 9, 5)):
        H, W, K = cfg
        n = 3000
        out = dict(actions=np.zeros((n, H * W), np.uint8), length=np.zeros(n, np.uint8), winner=np.zeros(n, np.int8),
                   final_grid=np.zeros((n, H, W), np.int8), reward=np.zeros((n, 2), np.float32), stats=np.zeros(256, np.int64))
        m.connect_rollout(H, W, K, n, 10, 3, **out)
        ref = oracle.connect_rollout(H, W, K, n, gid0=10, seed=3)
        for k, v in out.items():
            np.testing.assert_array_equal(v, ref[k], err_msg=k)
    grid0 = np.array(DEFAULT_BOUNCE_GRID, dtype=np.int8)
    n, T = 1000, 96
    out = dict(moves=np.zeros((n, T, 2), np.uint8), length=np.zeros(n, np.uint16), winner=np.zeros(n, np.int8),
               final_grid=np.zeros((n, 9, 6), np.int8), reward=np.zeros((n, 2), np.float32), stats=np.zeros(256, np.int64))
    m.bounce_rollout(grid0, 0, T, n, 5, 7, **out)
    ref = oracle.bounce_rollout(grid0, n, max_plies=T, gid0=5, seed=7)
    for k, v in out.items():
        np.testing.assert_array_equal(v, ref[k], err_msg=k)
    # errors come back as RuntimeError with the library's message; wrong sizes as TypeError
    with pytest.raises(RuntimeError, match="unsupported"):
        m.connect_rollout(16, 16, 4, 1, 0, 0, length=np.zeros(1, np.uint8))
    with pytest.raises(TypeError):
        m.connect_rollout(6, 7, 4, 5, 0, 0, length=np.zeros(4, np.uint8))


def test_cpp_extension_device_pointers(oracle):
    """Raw device addresses (what a DLPack consumer holds) through the C++ extension."""
    from simulator import _bgs_pybind as m

    H, W, K, n = 10, 12, 6, 4099
    dev = "cuda"
    actions = torch.empty((n, H * W), dtype=torch.uint8, device=dev)
    length = torch.empty(n, dtype=torch.uint8, device=dev)
    winner = torch.empty(n, dtype=torch.int8, device=dev)
    grid = torch.empty((n, H, W), dtype=torch.int8, device=dev)
    reward = torch.empty((n, 2), dtype=torch.float32, device=dev)
    stats = torch.zeros(256, dtype=torch.int64, device=dev)
    m.connect_rollout_device(H, W, K, n, 77, 5, actions.data_ptr(), length.data_ptr(), winner.data_ptr(), grid.data_ptr(),
                             reward.data_ptr(), stats.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    ref = oracle.connect_rollout(H, W, K, n, gid0=77, seed=5)
    for got, k in ((actions, "actions"), (length, "length"), (winner, "winner"), (grid, "final_grid"), (reward, "reward"), (stats, "stats")):
        np.testing.assert_array_equal(got.cpu().numpy(), ref[k], err_msg=k)
    # one batched transition
    g0 = torch.full((4, 6, 7), -1, dtype=torch.int8, device=dev)
    pl = torch.zeros(4, dtype=torch.int8, device=dev)
    wi = torch.full((4,), -1, dtype=torch.int8, device=dev)
    act = torch.tensor([0, 3, 6, 9], dtype=torch.int32, device=dev)
    go, po, wo = torch.empty_like(g0), torch.empty_like(pl), torch.empty_like(wi)
    st = torch.empty(4, dtype=torch.int32, device=dev)
    m.connect_step_device(6, 7, 4, 4, g0.data_ptr(), pl.data_ptr(), wi.data_ptr(), act.data_ptr(), go.data_ptr(), po.data_ptr(),
                          wo.data_ptr(), 0, 0, 0, st.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert st.tolist() == [0, 0, 0, 1] and go[1, 0, 3] == 0 and po.tolist() == [1, 1, 1, 0]
