"""CPU: the host-side logic of the multi-GPU path, exercised with a world_size-2 gloo group.

The kernels themselves need a GPU (tests/multigpu_check.py runs them under torchrun on >= 2 B200s); here we
check what the host decides: shards are disjoint, complete, balanced and never split a key, and the
combination rule used on the device (sum of rank-local tables / Gram partials == the global quantity) holds.
"""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from leanfe_b200.sharding import key_range_cuts, shard_rows_by_key

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_key_range_cuts_balanced():
    cnt = np.array([5, 1, 1, 1, 8, 2, 2, 4])
    cuts = key_range_cuts(cnt, 2)
    assert cuts[0] == 0 and cuts[-1] == cnt.size and np.all(np.diff(cuts) >= 0)
    left = cnt[: cuts[1]].sum()
    assert abs(left - cnt.sum() / 2) <= cnt.max()
    assert list(key_range_cuts(np.array([3]), 4)) == [0, 1, 1, 1, 1]          # more ranks than keys: empty shards
    assert list(key_range_cuts(np.array([], dtype=np.int64), 2)) == [0, 0, 0]


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_shards_are_a_partition_and_keep_keys_together(world):
    rng = np.random.default_rng(world)
    keys = rng.integers(0, 500, 20_000) * 13 - 99
    shards = shard_rows_by_key(keys, world)
    allrows = np.concatenate(shards)
    assert np.array_equal(np.sort(allrows), np.arange(keys.size))               # disjoint + complete
    owners = {}
    for r, idx in enumerate(shards):
        assert np.all(np.diff(idx) > 0)                                          # row order preserved
        for kv in np.unique(keys[idx]):
            assert owners.setdefault(int(kv), r) == r                            # a key lives on one rank
    sizes = np.array([s.size for s in shards])
    assert sizes.max() - sizes.min() <= 2 * np.bincount((keys + 99) // 13).max() + 1


_WORKER = r"""
import os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from leanfe_b200.sharding import shard_rows_by_key
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
rng = np.random.default_rng(7)
n = 30_000
unit, time_ = rng.integers(0, 700, n), rng.integers(0, 9, n)
x, y = rng.normal(size=n), rng.normal(size=n)
mine = shard_rows_by_key(unit, world)[rank]
# (1) the rank-local table of the shard key needs no exchange: groups are complete on their owner
cnt_local = np.bincount(unit[mine], minlength=700)
cnt_full = np.bincount(unit, minlength=700)
assert np.all((cnt_local == 0) | (cnt_local == cnt_full))
# (2) non-local FE tables and Gram partials are combined by a sum all-reduce
t = torch.from_numpy(np.bincount(time_[mine], weights=y[mine], minlength=9))
dist.all_reduce(t)
assert np.allclose(t.numpy(), np.bincount(time_, weights=y, minlength=9), rtol=1e-12)
Z = np.column_stack([np.ones(mine.size), x[mine], y[mine]])
g = torch.from_numpy(Z.T @ Z)
dist.all_reduce(g)
Zf = np.column_stack([np.ones(n), x, y])
assert np.allclose(g.numpy(), Zf.T @ Zf, rtol=1e-12)
# (3) kept-row counts: local sums to the global n_obs
k = torch.tensor([mine.size], dtype=torch.int64)
dist.all_reduce(k)
assert int(k) == n
dist.destroy_process_group()
print("ok", rank)
"""


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_gloo_world2_combination_rules(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(_free_port()), str(script), ROOT]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-3000:]
    assert out.stdout.count("ok") == 2


@pytest.mark.gpu
def test_multigpu_parity_under_torchrun():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "multigpu_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "multigpu_check: OK" in out.stdout
