"""Drop-in for the reference module sampler/jumpy_sampler.py (`DiffusionJumpySampler`, `ModelAdapter`)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _path  # noqa: E402,F401
from fddm_b200.sampler import DiffusionJumpySampler, ModelAdapter  # noqa: E402,F401
