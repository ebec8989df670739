"""Pipeline seam: serve ``extract_*`` (and pre-located ``tile``) steps of an ALIBY pipeline
from the CUDA hot path.

The reference has no plugin registry; a step is chosen by the prefix of its name inside
``init_step`` (``src/aliby/pipe.py:47-72``) and the sanctioned seam is the ``init_step_fn``
argument of ``_run_pipeline_and_post_impl`` (``src/aliby/pipe_core.py:381-389``).  This module
provides exactly that:

* :func:`init_step` — returns our extractor for ``extract_*`` steps (same ``partial`` shape as
  ``pipe_core._init_extract``, ``pipe_core.py:68-81``) and defers every other step to the
  reference's own ``aliby.pipe.init_step`` when ALIBY is importable;
* :func:`run_pipeline_and_post` — ``partial(_run_pipeline_and_post_impl, init_step_fn=init_step)``
  (what ``aliby/pipe.py:75-77`` does with its own ``init_step``), available when ALIBY is importable;
* :func:`get_profiles_from_state` — table assembly of ``pipe_core.py:453-512`` (rename
  ``tile/label`` -> ``metadata_*``, add ``metadata_object`` / ``metadata_tp`` (uint16), concatenate
  per step prefix, join ``extract`` with ``extractmulti``) on top of our ``format_extraction``.
"""

from __future__ import annotations

from functools import partial

import numpy as np

from .extract import extract_tree, format_extraction, process_tree_masks, process_tree_masks_overlap


def _init_extract(step_name: str, parameters: dict, *, overlap: bool = False):
    """pipe_core.py:68-81: ``partial(process, measure_fn=..., tree=..., **kwargs)``."""
    if "tree" not in parameters:
        raise ValueError(f"Step '{step_name}' is missing required 'tree'.")
    process = process_tree_masks
    measure_fn = extract_tree
    if overlap:
        process = process_tree_masks_overlap
        measure_fn = partial(extract_tree, overlap=True)
    return partial(process, measure_fn=measure_fn, tree=parameters["tree"], **parameters.get("kwargs", {}))


def init_step(step_name: str, parameters: dict, other_steps: dict | None = None, *, overlap: bool = False):
    """Drop-in ``init_step_fn``: ours for ``extract_*``, the reference's for everything else."""
    if step_name.startswith("extract_"):
        return _init_extract(step_name, parameters, overlap=overlap)
    if step_name.startswith("extractmulti_"):
        raise NotImplementedError(
            "extractmulti_* (cp_measure colocalisation) has no CUDA kernel in aliby_b200; "
            "route this step through aliby.pipe.init_step"
        )
    try:
        from aliby.pipe import init_step as reference_init_step
    except ImportError as e:  # ALIBY itself is not installed next to us
        raise ImportError(
            f"step '{step_name}' is not an extract step and the reference (aliby.pipe.init_step) is not importable"
        ) from e
    return reference_init_step(step_name, parameters, other_steps)


def run_pipeline_and_post(*args, **kwargs):
    """``aliby.pipe.run_pipeline_and_post`` with the extract steps served by the GPU."""
    from aliby.pipe_core import _run_pipeline_and_post_impl

    return _run_pipeline_and_post_impl(*args, init_step_fn=init_step, **kwargs)


def get_profiles_from_state(state: dict, pipeline: dict):
    """Wide profile table of all feature steps (pipe_core.py:453-512)."""
    import pyarrow as pa

    profiles = pa.Table.from_pylist(
        [],
        schema=pa.schema(
            [
                pa.field("metadata_tile", pa.int64()),
                pa.field("metadata_label", pa.int64()),
                pa.field("metadata_object", pa.string()),
                pa.field("metadata_tp", pa.int64()),
            ]
        ),
    )
    feature_steps = [s for s in pipeline["steps"] if s.startswith("extract") or s.startswith("nahual_embed")]
    data = {k.split("_")[0]: [] for k in feature_steps}
    for ext_step in feature_steps:
        prefix = ext_step.split("_")[0]
        for tp, ext_output in enumerate(state["data"][ext_step]):
            if isinstance(ext_output, np.ndarray):  # arbitrary embedders: one (instructions, metrics) pair
                ext_output = ((("__", "__"),), (ext_output,))
            table = format_extraction(ext_output)
            rename = {"tile": "metadata_tile", "label": "metadata_label"}
            table = table.rename_columns([rename.get(c, c) for c in table.column_names])
            if len(table):
                table = table.append_column("metadata_object", pa.array([ext_step.split("_")[-1]] * len(table), pa.string()))
                table = table.append_column("metadata_tp", pa.array([tp] * len(table), pa.uint16()))
                data[prefix].append(table)
    wide = [pa.concat_tables(t) for t in data.values() if len(t)]
    if wide:
        profiles = wide[0]
        for table in wide[1:]:
            profiles = profiles.join(table, keys=[f"metadata_{k}" for k in ("tp", "tile", "object", "label")])
    return profiles


def profiles_from_tables(tables, object_name: str, prefix: str = "extract"):
    """Profile table of one extract step straight from dense per-time-point results.

    ``tables[tp]`` is an :class:`aliby_b200.extract.ExtractionTable` (or ``None`` for a time point without
    objects).  Produces what :func:`get_profiles_from_state` builds for a step named ``{prefix}_{object_name}``
    (pipe_core.py:482-497): ``metadata_tile, metadata_label, <sorted features>, metadata_object, metadata_tp``
    (uint16), rows concatenated over time points — without the per-item Python lists."""
    import pyarrow as pa

    parts = []
    for tp, tab in enumerate(tables):
        if tab is None or not len(tab.objects):
            continue
        t = tab.to_arrow()
        t = t.rename_columns([{"tile": "metadata_tile", "label": "metadata_label"}.get(c, c) for c in t.column_names])
        t = t.append_column("metadata_object", pa.array([object_name] * len(t), pa.string()))
        t = t.append_column("metadata_tp", pa.array(np.full(len(t), tp, dtype=np.uint16), pa.uint16()))
        parts.append(t)
    if not parts:
        return pa.Table.from_pylist([], schema=pa.schema([
            pa.field("metadata_tile", pa.int64()), pa.field("metadata_label", pa.int64()),
            pa.field("metadata_object", pa.string()), pa.field("metadata_tp", pa.int64())]))
    return pa.concat_tables(parts)


def write_profiles(profiles, output_path, pipeline_name: str):
    """``<output_path>/profiles/<pipeline_name>.parquet`` with zstd compression (pipe_core.py:403,411-413)."""
    from pathlib import Path

    import pyarrow.parquet as pq

    path = Path(output_path) / "profiles" / f"{pipeline_name}.parquet"
    path.parent.mkdir(parents=True, exist_ok=True)
    pq.write_table(profiles, path, compression="zstd")
    return path
