"""Drop-in import path: `from optimizer_GD import GradientDescentOptimizer`."""
from irm_motion_planning_b200.optimizer_GD import GradientDescentOptimizer  # noqa: F401
