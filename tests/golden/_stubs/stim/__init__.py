"""Import-only stand-in for `stim` (absent from this image).  See galois stub for why."""


class DetectorErrorModel:  # type annotation target only
    pass


class Circuit:
    def __init__(self, *a, **k):
        raise NotImplementedError("stim stub: no circuit simulation in this container")
