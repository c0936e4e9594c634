"""Post-filters of the recommendation lists -- drop-ins for recpack.postprocessing.ExcludeItems / SelectItems
(recpack/postprocessing/filters.py:58-101).

With ``recpack`` importable the classes subclass the reference's, so ``apply(X_pred)`` on a full prediction matrix
is the reference's own code and a ``PipelineBuilder.add_post_filter`` accepts them.  What they add is ``item_mask``:
the GPU scorer applies the filters INSIDE predict (``algo.set_postfilters([...])``), before the lists are truncated
to N -- a filter applied after the truncation would leave holes that the next-best items should have filled."""
from __future__ import annotations

import numpy as np

from . import _ref

if _ref.HAVE_RECPACK:
    import importlib

    _f = importlib.import_module(_ref.ref_base.__name__.split(".")[0] + ".postprocessing.filters")
    _Exclude, _Select = _f.ExcludeItems, _f.SelectItems
else:

    class _PostFilter:
        def apply_all(self, *matrices):
            return [self.apply(m) for m in matrices]

        def __str__(self):
            return f"{type(self).__name__}({', '.join(f'{k}={v}' for k, v in self.__dict__.items())})"

    class _Exclude(_PostFilter):
        def __init__(self, items):
            self.items = items

        def apply(self, X_pred):
            return X_pred.multiply(self.item_mask(X_pred.shape[1]).astype(bool)).tocsr()

    class _Select(_Exclude):
        pass


def _check(items, n_items):
    items = np.asarray(items)
    if len(items) == 0 or np.amax(items) > n_items:  # filters.py:71-72 / 93-94
        raise ValueError(f"{n_items} items and {items.shape}")
    return items


class ExcludeItems(_Exclude):
    """Remove the recommendations of the given items."""

    def item_mask(self, n_items: int) -> np.ndarray:
        mask = np.ones(n_items, dtype=np.uint8)
        mask[_check(self.items, n_items)] = 0
        return mask


class SelectItems(_Select):
    """Keep only the recommendations of the given items."""

    def item_mask(self, n_items: int) -> np.ndarray:
        mask = np.zeros(n_items, dtype=np.uint8)
        mask[_check(self.items, n_items)] = 1
        return mask


def combined_mask(filters, n_items: int):
    """uint8[n_items], 1 = every filter lets the item through (None without filters)."""
    mask = None
    for f in filters or ():
        m = f.item_mask(n_items)
        mask = m if mask is None else (mask & m)
    return None if mask is None else np.ascontiguousarray(mask)
