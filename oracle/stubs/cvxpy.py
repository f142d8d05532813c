"""Stub: names only. Any attempt to build a problem raises."""
OPTIMAL = "optimal"
class Expression: pass
class Constraint: pass
class Parameter(Expression):
    def __init__(self, *a, **k): raise RuntimeError("cvxpy stub: not available in this container")
class Variable(Expression):
    def __init__(self, *a, **k): raise RuntimeError("cvxpy stub: not available in this container")
