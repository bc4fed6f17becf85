"""Shim so the read-only reference imports in this container (lsq-ellipse is
not installed; it is only used by training-target code, never on the hot path)."""


class LsqEllipse:  # pragma: no cover
    def fit(self, *a, **k):
        raise RuntimeError("lsq-ellipse shim: not available")
