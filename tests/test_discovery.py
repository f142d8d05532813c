"""The reference finds its plug-ins by NAME: `gather_custom_modules` (utils.py:200-235) walks a directory, parses every file and
collects the classes whose base-class name contains a given string (utils.py:132-147); main.py then imports them by module path.
The same walk over this package must find every environment, safeguard and learning algorithm of the hot path under the
reference's class names, and the modules it names must import and expose the class."""
import ast
import importlib
from pathlib import Path

import safegbpo_b200

PKG = Path(safegbpo_b200.__file__).resolve().parent


def gather(directory: Path, subclass):
    """utils.py:200-235 restated: {class name: module name}."""
    found = {}
    for path in sorted(directory.rglob("*.py")):
        if path.name == "__init__.py":
            continue
        tree = ast.parse(path.read_text(), filename=str(path))
        module = ".".join(path.resolve().relative_to(PKG.parent).with_suffix("").parts)
        for node in ast.walk(tree):
            if isinstance(node, ast.ClassDef):
                bases = [b.id if isinstance(b, ast.Name) else b.attr if isinstance(b, ast.Attribute) else "" for b in node.bases]
                if subclass is None or any(subclass in b for b in bases):
                    found[node.name] = module
    return found


def test_environments_safeguards_algorithms_are_discoverable():
    envs = gather(PKG / "envs", "Env")
    for name in ("BalancePendulumEnv", "BalanceCartPoleEnv", "SwingUpCartPoleEnv", "BalanceQuadrotorEnv", "ManageHouseholdEnv",
                 "NavigateSeekerEnv", "NavigateQuadrotorEnv"):
        assert name in envs, name
    guards = gather(PKG / "safeguards", "Safeguard")
    assert {"BoundaryProjectionSafeguard", "RayMaskSafeguard"} <= set(guards)
    algos = gather(PKG / "learning_algorithms", "LearningAlgorithm")
    assert "SHAC" in algos
    for table in (envs, guards, algos):
        for name, module in table.items():
            assert hasattr(importlib.import_module(module), name), (module, name)


def test_wrapped_environment_is_a_simulator():
    """safeguard.py:229 `Simulator.register(Safeguard)`."""
    from safegbpo_b200.envs.simulators.interfaces.simulator import Simulator
    from safegbpo_b200.safeguards.interfaces.safeguard import Safeguard
    from safegbpo_b200.safeguards.boundary_projection import BoundaryProjectionSafeguard
    assert issubclass(Safeguard, Simulator) and issubclass(BoundaryProjectionSafeguard, Simulator)
