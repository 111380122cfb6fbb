"""Stub for the absent `mlconfig` package so /root/reference/model/centernet.py imports.
Test infrastructure only (used by oracle/gen_golden.py in the build container)."""


def register(obj=None, *a, **k):
    if callable(obj):
        return obj
    return lambda f: f
