"""Oracle shim: `knn_points` is imported by lib/model/aggregation.py:15 but never called on the live path."""


def knn_points(*args, **kwargs):  # pragma: no cover
    raise NotImplementedError("pytorch3d.ops.knn_points is not on the VPHO eval hot path")
