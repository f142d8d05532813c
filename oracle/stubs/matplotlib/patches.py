class Polygon:
    def __init__(self, *a, **k): pass
class Circle:
    def __init__(self, *a, **k): pass
class Arc:
    def __init__(self, *a, **k): pass
