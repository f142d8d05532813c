class VectorEnv:
    def close(self): pass
class VectorWrapper(VectorEnv):
    def __init__(self, env): self.env = env
class VectorActionWrapper(VectorWrapper):
    def step(self, actions): return self.env.step(self.actions(actions))
    def actions(self, actions): raise NotImplementedError
