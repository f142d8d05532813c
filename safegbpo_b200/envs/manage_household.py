"""ManageHousehold (reference: src/envs/manage_household.py)."""
from .. import sets
from .interfaces.safe_state_env import SafeStateEnv
from .simulators.household import HouseholdEnv


class ManageHouseholdEnv(HouseholdEnv, SafeStateEnv):
    ENV_NAME = "ManageHousehold"

    def __init__(self, num_envs: int, num_steps: int):
        SafeStateEnv.__init__(self, num_state_gens=3)
        HouseholdEnv.__init__(self, num_envs, num_steps)

    def safe_state_set(self) -> sets.Zonotope:
        return self.state_set
