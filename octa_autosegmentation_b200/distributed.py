"""Multi-GPU plumbing.  The path shards per sample (every graph is independent: own config, own two RNG
streams, no shared state -- the reference already runs one process per sample,
generate_vessel_graph.py:125-126), so ranks own disjoint samples and there is NO data-path collective.
The only exchange is the optional gather of finished edge tables (and label buffers) to rank 0 at the end,
one padded `gather` over the process group (NCCL on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

import numpy as np


def gather_edge_tables(tables: dict, dst: int = 0, device=None) -> dict:
    """tables: {sample_id: float64 [E_i, 7]} owned by this rank.  Returns the union on rank `dst`, {} elsewhere."""
    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(), dist.get_rank()
    dev = device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu")
    ids = sorted(tables)
    meta = torch.tensor([len(ids), sum(len(tables[i]) for i in ids)], dtype=torch.int64, device=dev)
    metas = [torch.zeros_like(meta) for _ in range(world)]
    dist.all_gather(metas, meta)
    max_ids = int(max(m[0] for m in metas))
    max_rows = int(max(m[1] for m in metas))
    # one pinned staging block per rank: (sample id, rows) heads + all rows back to back, uploaded with ONE copy; rank `dst`
    # brings every rank's block back with one copy each and slices it on the host
    head_h = np.zeros((max(max_ids, 1), 2), dtype=np.int64)
    body_h = np.zeros((max(max_rows, 1), 7), dtype=np.float64)
    r = 0
    for k, i in enumerate(ids):
        t = np.ascontiguousarray(tables[i], dtype=np.float64).reshape(-1, 7)
        head_h[k] = (int(i), len(t))
        body_h[r:r + len(t)] = t
        r += len(t)
    head = torch.from_numpy(head_h).to(dev)
    body = torch.from_numpy(body_h).to(dev)
    heads = [torch.zeros_like(head) for _ in range(world)] if rank == dst else None
    bodies = [torch.zeros_like(body) for _ in range(world)] if rank == dst else None
    dist.gather(head, heads, dst=dst)
    dist.gather(body, bodies, dst=dst)          # the one data exchange of the job
    out = {}
    if rank == dst:
        for w in range(world):
            hw, bw = heads[w].cpu().numpy(), bodies[w].cpu().numpy()
            r = 0
            for k in range(int(metas[w][0])):
                sid, rows = int(hw[k, 0]), int(hw[k, 1])
                out[sid] = bw[r:r + rows]
                r += rows
    return out
