"""World-size-2 gloo worker for tests/test_host_logic.py: every rank fabricates the edge tables of its shard
(deterministic in the seed), rank 0 gathers them with octa_autosegmentation_b200.distributed.gather_edge_tables and
checks that the union equals the single-process result."""
import os
import sys

import numpy as np
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from octa_autosegmentation_b200 import distributed as od  # noqa: E402
from octa_autosegmentation_b200.pipeline import shard_seeds  # noqa: E402


def fake_graph(seed):
    rng = np.random.RandomState(seed)
    return rng.uniform(0, 1, (10 + seed % 7, 7))


dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
mine = shard_seeds(50, 9, rank, world)
tables = {s: fake_graph(s) for s in mine}
gathered = od.gather_edge_tables(tables, dst=0)
if rank == 0:
    assert sorted(gathered) == list(range(50, 59))
    for s, t in gathered.items():
        assert np.array_equal(t, fake_graph(s))
    print("GATHER_OK")
dist.barrier()
dist.destroy_process_group()
