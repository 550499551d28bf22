"""Drop-in import path: `from optimizer_BLS import BacktrackingLineSearchOptimizer`."""
from irm_motion_planning_b200.optimizer_BLS import BacktrackingLineSearchOptimizer  # noqa: F401
