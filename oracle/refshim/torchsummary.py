"""Stub for the absent `torchsummary` package (the reference only calls it under __main__)."""


def summary(*a, **k):
    raise RuntimeError("torchsummary stub")
