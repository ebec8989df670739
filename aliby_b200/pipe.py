"""Pipeline seam: serve the ``extract_*`` and ``tile*`` steps of an ALIBY pipeline from the CUDA hot path.

The reference has no plugin registry; a step is chosen by the prefix of its name inside
``init_step`` (``src/aliby/pipe.py:47-72``) and the sanctioned seam is the ``init_step_fn``
argument of ``_run_pipeline_and_post_impl`` (``src/aliby/pipe_core.py:381-389``).  This module
provides exactly that:

* :func:`init_step` — returns our extractor for ``extract_*`` steps (same ``partial`` shape as
  ``pipe_core._init_extract``, ``pipe_core.py:68-81``), our fused tiler for ``tile*`` steps (the crop of
  ``tiler.py:309-366`` becomes a view that the extraction kernels read in place), our two-image extractor for
  ``extractmulti_*`` steps (``pipe_core._init_extract_multi``, ``pipe_core.py:84-92``; features without a kernel are
  measured by the reference's step next to ours) and defers every other step to the reference's own
  ``aliby.pipe.init_step`` when ALIBY is importable;
* :func:`run_pipeline_and_post` — ``partial(_run_pipeline_and_post_impl, init_step_fn=init_step)``
  (what ``aliby/pipe.py:75-77`` does with its own ``init_step``), available when ALIBY is importable;
* :func:`get_profiles_from_state` — table assembly of ``pipe_core.py:453-512`` (rename
  ``tile/label`` -> ``metadata_*``, add ``metadata_object`` / ``metadata_tp`` (uint16), concatenate
  per step prefix, join ``extract`` with ``extractmulti``) on top of our ``format_extraction``.
"""

from __future__ import annotations

from functools import partial

import numpy as np

from .extract import (extract_tree, extract_tree_multi, format_extraction, process_tree_masks,
                      process_tree_masks_overlap)


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


def _split_tree(tree: dict, cp_measure_kwargs=None):
    """``(ours, theirs)``: the branches of an extract / extractmulti tree with and without a CUDA kernel (same nesting)."""
    from . import engine

    ours: dict = {}
    theirs: dict = {}
    for inst in engine.kv(engine.flatten(tree)):
        ok = engine.compile_instructions([inst], cp_measure_kwargs).error is None
        node = ours if ok else theirs
        for key in inst[:-2]:
            node = node.setdefault(key, {})
        node.setdefault(inst[-2], []).append(inst[-1])
    return ours, theirs


_split_multi_tree = _split_tree


def _split_step(step_name: str, parameters: dict, other_steps: dict | None, make_gpu_step, why: str):
    """A step whose tree names features without a kernel next to features with one: our kernels measure their part, the
    reference's own step — when ALIBY is importable — measures the rest with the same arguments, and the two
    ``(instructions, results)`` lists are concatenated.  ``format_extraction`` pivots by (tile, label) and sorts the
    columns (extract.py:574-598), so the table is the one the reference builds.  Without the reference the whole tree
    stays with us and raises at its first call with objects, naming the feature."""
    kwargs = dict(parameters.get("kwargs", {}))
    ours, theirs = _split_tree(parameters["tree"], kwargs.get("cp_measure_kwargs"))
    if not theirs:
        return make_gpu_step(parameters["tree"], kwargs)
    try:
        reference_step = _reference_init_step(step_name, {**parameters, "tree": theirs}, other_steps, why)
    except ImportError:
        return make_gpu_step(parameters["tree"], kwargs)
    if not ours:
        return reference_step
    gpu_step = make_gpu_step(ours, kwargs)

    def split_step(masks, pixels, **kw):
        from .tile import TileView

        items_a, res_a = gpu_step(masks=masks, pixels=pixels, **kw)
        # (the reference indexes pixels[tile, channel]: a fused tile view is materialised for it)
        items_b, res_b = reference_step(masks=masks, pixels=np.asarray(pixels) if isinstance(pixels, TileView) else pixels, **kw)
        return tuple(items_a) + tuple(items_b), list(res_a) + list(res_b)

    return split_step


def _init_extract_multi(step_name: str, parameters: dict, other_steps: dict | None = None):
    """``extractmulti_*`` steps (pipe.py:65-66, pipe_core.py:84-92): ``partial(process_tree_masks,
    measure_fn=extract_tree_multi, tree=..., **kwargs)`` on the GPU; two-image features without a kernel (``costes``,
    which the stock builder requests, pipe_builder.py:19-43) go to the reference's step (:func:`_split_step`)."""
    if "tree" not in parameters:
        raise ValueError(f"Step '{step_name}' is missing required 'tree'.")
    return _split_step(step_name, parameters, other_steps,
                       lambda tree, kwargs: partial(process_tree_masks, measure_fn=extract_tree_multi, tree=tree, **kwargs),
                       "names two-image features without a CUDA kernel")


def _reference_init_step(step_name: str, parameters: dict, other_steps: dict | None, why: str):
    """Hand a step to the reference's own ``aliby.pipe.init_step`` (pipe.py:47-72)."""
    try:
        from aliby.pipe import init_step as reference_init_step
    except ImportError as e:  # ALIBY itself is not installed next to us
        raise ImportError(f"step '{step_name}' {why} and the reference (aliby.pipe.init_step) is not importable") from e
    return reference_init_step(step_name, parameters, other_steps)


def _init_tile(step_name: str, parameters: dict, other_steps: dict | None):
    """``tile*`` steps (pipe.py:56-57, pipe_core.py:54-65) with the crop fused into the extraction.

    * Pre-located tiles — ``{"pixels": <(T, C, Z, Y, X) array>, "tile_centres": [(row, col), ...], "tile_size": n}``
      (what a position looks like after the reference's trap detection at time point 0, tiler.py:407-417) — are served
      by :class:`aliby_b200.tile.FusedTiler` alone.
    * Anything else is initialised by the reference (image readers, trap detection, drift: out of scope here) and its
      crop method is swapped for the fused view (:func:`aliby_b200.tile.fuse_reference_tiler`): ``run_tp`` then returns
      ``{"drift": ..., "pixels": TileView}`` and the extract step reads the tile windows in place.
    """
    from .tile import FusedTiler, fuse_reference_tiler

    if "tile_centres" in parameters:
        if "pixels" not in parameters or "tile_size" not in parameters:
            raise ValueError(f"Step '{step_name}' with 'tile_centres' also needs 'pixels' and 'tile_size'.")
        return FusedTiler(parameters["pixels"], parameters["tile_centres"], parameters["tile_size"])
    return fuse_reference_tiler(_reference_init_step(step_name, parameters, other_steps, "needs the reference's image readers"))


def init_step(step_name: str, parameters: dict, other_steps: dict | None = None, *, overlap: bool = False):
    """Drop-in ``init_step_fn``: ours for ``extract_*`` and ``tile*``, the reference's for everything else.

    The branches of an ``extract_*`` / ``extractmulti_*`` tree that name a metric without a CUDA kernel (cp_measure
    features beyond ``intensity`` / ``sizeshape`` / the four two-image features, decided here with the plan compiler) go
    to the reference's own step when ALIBY is importable, the rest of the tree runs on the GPU (:func:`_split_step`);
    otherwise the error names the metric — there is no CPU fallback inside this package."""
    if step_name.startswith("extract_"):
        if "tree" not in parameters or overlap:
            return _init_extract(step_name, parameters, overlap=overlap)
        return _split_step(step_name, parameters, other_steps,
                           lambda tree, kwargs: _init_extract(step_name, {"tree": tree, "kwargs": kwargs}),
                           "names features without a CUDA kernel")
    if step_name.startswith("extractmulti_"):
        return _init_extract_multi(step_name, parameters, other_steps)
    if step_name.startswith("tile"):
        return _init_tile(step_name, parameters, other_steps)
    return _reference_init_step(step_name, parameters, other_steps, "is not an extract or tile step")


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
