"""B200-native batched functional-gradient-descent trajectory optimisation.

Drop-in for the FGD hot path of simongroeger/irm_motion_planning: the Python
entry points keep the reference's names (``main``, ``optimizer_BLS``,
``optimizer_GD``, ``trajectory``, ``robot``, ``environment``); the iteration
itself runs in hand-written sm_100a CUDA kernels behind ``include/fgd_b200.h``.
"""
__version__ = "0.1.0"
