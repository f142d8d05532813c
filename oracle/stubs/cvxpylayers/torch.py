class CvxpyLayer:
    def __init__(self, *a, **k): raise RuntimeError("cvxpylayers stub: not available in this container")
