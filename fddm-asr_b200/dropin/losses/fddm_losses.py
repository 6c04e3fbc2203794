"""Drop-in for the reference module losses/fddm_losses.py (`lfd_loss`)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _path  # noqa: E402,F401
from fddm_b200.losses import lfd_loss  # noqa: E402,F401
