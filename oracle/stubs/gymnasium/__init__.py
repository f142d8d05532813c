from . import vector
