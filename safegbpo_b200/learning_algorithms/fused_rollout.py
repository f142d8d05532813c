"""Whole-horizon fused rollout engine (placeholder until csrc/sgbpo_rollout.cu lands)."""


class FusedRollout:
    @staticmethod
    def try_create(algo):
        return None
