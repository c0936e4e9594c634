"""Binding point to the reference package.

When ``recpack`` is importable (the normal case for a drop-in: the user already has it) the classes of
this package SUBCLASS the reference's own ``TopKItemSimilarityMatrixAlgorithm`` / ``ItemKNN`` /
``ListwiseMetricK`` ... classes, so ``isinstance`` checks, the ``fit`` / ``predict`` wrappers, ``name``,
``identifier``, the constructor validation and the registries are the reference's code, not a restatement
(SURVEY.md 8b).  Only ``_fit`` / ``_predict`` / ``calculate`` and the two O(nnz) Python-set checks are
replaced.  Without ``recpack`` (or with RPK_NO_RECPACK=1) a small stand-alone mirror of those wrappers is
used instead (``_mirror.py``).  Nothing on the compute path comes from ``recpack`` either way."""
from __future__ import annotations

import os

HAVE_RECPACK = False
ref_base = ref_nn = ref_metrics = ref_metrics_base = ref_matrix_util = ref_splitters = None

if os.environ.get("RPK_NO_RECPACK", "0") != "1":
    try:
        import recpack.algorithms.base as ref_base
        import recpack.algorithms.nearest_neighbour as ref_nn
        import recpack.matrix.util as ref_matrix_util
        import recpack.metrics as ref_metrics
        import recpack.metrics.base as ref_metrics_base

        try:  # needs pandas / tqdm like the reference itself; optional for the algorithm and metric classes
            import recpack.scenarios.splitters as ref_splitters
        except Exception:
            ref_splitters = None

        HAVE_RECPACK = True
    except Exception:  # not installed, or an incompatible environment: use the mirror
        HAVE_RECPACK = False
        ref_base = ref_nn = ref_metrics = ref_metrics_base = ref_matrix_util = ref_splitters = None
