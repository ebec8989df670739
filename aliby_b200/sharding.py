"""Multi-GPU driver for the extraction hot path: shard independent units, gather tables.

Every (position / field) x (time point) of an experiment is independent of every other: the
reference itself fans out per position (``examples/01_cell_painting_tiff.py:141-144``) and
writes one parquet per position (``src/aliby/pipe_core.py:403``).  So the units are partitioned
across ranks — one process per GPU — with NO collective on the data path; only the final
per-object tables travel, and only to the host of rank 0 (``torch.distributed.gather_object``;
over gloo on CPU in the tests, over NCCL-initialised groups the object gather uses the
host path as well).
"""

from __future__ import annotations

import numpy as np


def shard_units(n_units: int, rank: int, world: int, mode: str = "contiguous") -> np.ndarray:
    """Indices of the units owned by ``rank``.

    ``contiguous`` blocks keep the time points of one position together (the tile table is fixed
    after time point 0); ``round_robin`` balances fields of unequal cost."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside [0, {world})")
    idx = np.arange(n_units)
    if mode == "round_robin":
        return idx[rank::world]
    if mode == "contiguous":
        bounds = np.linspace(0, n_units, world + 1).round().astype(int)
        return idx[bounds[rank] : bounds[rank + 1]]
    raise ValueError(f"unknown sharding mode {mode!r}")


def extract_sharded(tree: dict, units, load_unit, compute=None, rank: int | None = None, world: int | None = None,
                    mode: str = "contiguous", gather: bool = True):
    """Run the extraction of ``units`` sharded over the ranks of the default process group.

    ``load_unit(u) -> (masks, pixels)`` produces the inputs of one unit on the calling rank;
    ``compute(tree, masks, pixels) -> (objects (n, 2), names, values (n, k))`` defaults to the CUDA
    path (:func:`aliby_b200.extract.extract_table`).  Returns, on rank 0 (or on every rank when
    ``gather=False`` for the local part), a list of ``(unit, objects, names, values)`` ordered by unit.
    """
    import torch.distributed as dist

    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    if compute is None:
        from .extract import extract_table

        def compute(tree_, masks, pixels):
            t = extract_table(tree_, masks, pixels)
            return t.objects, t.names, t.values

    mine = shard_units(len(units), rank, world, mode)
    local = []
    for i in mine:
        masks, pixels = load_unit(units[i])
        objects, names, values = compute(tree, masks, pixels)
        local.append((int(i), units[i], np.asarray(objects), list(names), np.asarray(values)))
    if not gather or world == 1:
        return [(u, o, n, v) for _, u, o, n, v in sorted(local, key=lambda t: t[0])]
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(local, gathered, dst=0)
    if rank != 0:
        return None
    merged = sorted((item for part in gathered for item in part), key=lambda t: t[0])
    return [(u, o, n, v) for _, u, o, n, v in merged]
