class Trial: pass
class TrialPruned(Exception): pass
