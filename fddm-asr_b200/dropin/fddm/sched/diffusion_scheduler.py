"""Drop-in for the reference module fddm/sched/diffusion_scheduler.py: same import path and class,
B200-native implementation (see fddm_b200/scheduler.py).  Put `fddm-asr_b200/dropin` ahead of the
reference checkout on PYTHONPATH."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import _path  # noqa: E402,F401
from fddm_b200.scheduler import DiscreteDiffusionScheduler  # noqa: E402,F401
