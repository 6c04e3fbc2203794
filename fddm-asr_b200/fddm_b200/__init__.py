"""fddm_b200 -- B200-native (sm_100a) categorical discrete-diffusion token path of FDDM-ASR.

Host-side mirror of the reference's interfaces over libfddm_b200.so (include/fddm_b200.h):
  DiscreteDiffusionScheduler  <- fddm/sched/diffusion_scheduler.py
  SchedulerAdapter            <- train.py:176-273
  lfd_loss                    <- losses/fddm_losses.py
  DiffusionJumpySampler, ModelAdapter <- sampler/jumpy_sampler.py
  calculate_cer, calculate_wer (+ batch_*) <- models/evaluate.py:94-134
There is no CPU path: importing needs the built shared library, ops need CUDA tensors.
"""
from . import _lib
from ._lib import set_sm_reserve
from .scheduler import DiscreteDiffusionScheduler
from .adapter import SchedulerAdapter
from .sampler import DiffusionJumpySampler, ModelAdapter
from . import losses
from .losses import LfdPipeline, lfd_loss, symmetric_exchange_available
from .metrics import batch_cer, batch_wer, calculate_cer, calculate_wer

__all__ = ["DiscreteDiffusionScheduler", "SchedulerAdapter", "DiffusionJumpySampler", "ModelAdapter", "lfd_loss", "LfdPipeline",
           "_lib", "set_sm_reserve", "symmetric_exchange_available", "calculate_cer", "calculate_wer", "batch_cer", "batch_wer"]
