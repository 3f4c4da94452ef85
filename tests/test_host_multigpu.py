"""CPU tests of the multi-GPU host logic: the partition rule of afmg_partition and the
rank plumbing over torch.distributed with the gloo backend (world size 2).  No GPU calls."""
import os
import subprocess
import sys

import numpy as np
import pytest

from afivo_streamer_b200 import mg as M
from afivo_streamer_b200 import tree as T

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("n_ranks", [1, 2, 3, 4, 8])
def test_partition_properties(n_ranks):
    for tree in (T.uniform_tree(3, 8, 8, 4), T.corner_refined_tree(3, 8, 8, 5), T.shell_tree(8, 8, 4, 0.35),
                 T.build_tree(3, 8, [16, 8, 24], 3, lambda l, ix, c: np.linalg.norm(c - 0.4, axis=1) < 0.45)):
        counts = [len(a) for a in tree.lvl_ids]
        cuts = M.partition(n_ranks, counts)
        assert cuts.shape == (tree.highest_lvl, n_ranks + 1)
        for l, n in enumerate(counts):
            c = cuts[l]
            assert c[0] == 0 and c[-1] == n
            assert np.all(np.diff(c) >= 0)
            if l == 0:
                assert np.all(c[1:] == n)  # the coarse grid is rank 0's
            else:
                assert np.all(c % 8 == 0)  # sibling groups are never split
                sizes = np.diff(c)
                assert sizes.max() - sizes.min() <= 8  # balanced to one sibling group


def test_partition_s3_matches_design():
    counts = [1, 8, 64, 512, 4096, 32768, 216000]
    cuts = M.partition(8, counts)
    assert list(np.diff(cuts[-1])) == [27000] * 8
    assert list(cuts[1]) == [0] + [8] * 8  # 8 boxes = one sibling group -> rank 0


def test_partition_keeps_small_levels_on_rank_0():
    counts = [1, 8, 64, 512, 4096, 32768, 216000]
    cuts = M.partition(8, counts, min_split_boxes=1024)  # the default of a handle for 16^3 boxes
    for l in range(4):  # 1, 8, 64, 512 boxes: not split
        assert list(cuts[l]) == [0] + [counts[l]] * 8
    assert list(np.diff(cuts[4])) == [512] * 8
    assert list(np.diff(cuts[6])) == [27000] * 8
    assert np.array_equal(M.partition(8, counts, 0), M.partition(8, counts))


def test_partition_rejects_bad_arguments():
    with pytest.raises(M.AfmgError):
        M.partition(9, [1, 8])
    with pytest.raises(M.AfmgError):
        M.partition(0, [1, 8])


WORKER = r"""
import os, sys
sys.path.insert(0, sys.argv[1])
import numpy as np
import torch.distributed as dist
from afivo_streamer_b200 import mg as M, tree as T, _lib
dist.init_process_group("gloo")
rank, world, allgather = M.comm_from_torch()
assert world == 2 and rank == dist.get_rank()
tree = T.shell_tree(8, 8, 4, 0.35)
cuts = M.partition(world, [len(a) for a in tree.lvl_ids])
# every rank derives the same partition; the union of the ranges covers every level exactly once
blob = cuts.astype(np.int32).tobytes() + bytes([rank])
blobs = allgather(blob.ljust(_lib.AFMG_COMM_BLOB_BYTES * 4, b"\0"))
assert len(blobs) == world
assert blobs[0][:cuts.nbytes] == blobs[1][:cuts.nbytes]
assert blobs[0][cuts.nbytes] == 0 and blobs[1][cuts.nbytes] == 1
own = sum(int(cuts[l, rank + 1] - cuts[l, rank]) for l in range(tree.highest_lvl))
import torch
t = torch.tensor([own])
dist.all_reduce(t)
assert int(t.item()) == tree.n_boxes, (int(t.item()), tree.n_boxes)
dist.destroy_process_group()
print("ok", rank)
"""


def test_world_size_2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, OMP_NUM_THREADS="1")
    import socket
    with socket.socket() as sock:  # a free port: a fixed one may be held by another job on the machine
        sock.bind(("127.0.0.1", 0))
        port = sock.getsockname()[1]
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), str(script), ROOT],
                       capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert r.stdout.count("ok") == 2
