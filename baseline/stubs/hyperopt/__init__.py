"""Stand-in for the `hyperopt` package, which the reference imports at module level
(recpack/pipelines/pipeline.py:13) but which is not installed in this image (no network).

Only the names the import needs exist; calling the optimiser raises.  Grid search and plain
pipelines (the paths the parity tests and bench.py use) never touch hyperopt."""

STATUS_OK = "ok"


class Trials:  # pragma: no cover - placeholder
    def __init__(self, *a, **k):
        raise NotImplementedError("hyperopt is not installed in this image")


def fmin(*a, **k):  # pragma: no cover - placeholder
    raise NotImplementedError("hyperopt is not installed in this image")


def space_eval(*a, **k):  # pragma: no cover - placeholder
    raise NotImplementedError("hyperopt is not installed in this image")


class _NotAvailable:  # pragma: no cover - placeholder
    def __getattr__(self, name):
        raise NotImplementedError("hyperopt is not installed in this image")


tpe = _NotAvailable()
hp = _NotAvailable()
