"""Runs the reference's own `train.py` / `inference.py` UNCHANGED on top of the B200 token path.

    python fddm-asr_b200/dropin/launch.py /path/to/FDDM-asr train --config configs/fddm_zhTW_base.yaml
    python fddm-asr_b200/dropin/launch.py /path/to/FDDM-asr inference --wav a.wav --ckpt ... 

What it does:
  1. puts this directory ahead of the reference checkout on sys.path, so that the reference's
     `from fddm.sched.diffusion_scheduler import DiscreteDiffusionScheduler`,
     `from losses.fddm_losses import lfd_loss` and `from sampler.jumpy_sampler import ...` resolve to
     the shims here (models/, scripts/, configs/ still come from the reference);
  2. `SchedulerAdapter` is defined inside train.py itself (train.py:176-273), so after importing the
     reference's `train` module its `SchedulerAdapter` name is rebound to ours before `main()` runs.
"""
import os
import runpy
import sys


def main():
    if len(sys.argv) < 3 or sys.argv[2] not in ("train", "inference"):
        raise SystemExit(__doc__)
    ref, which = os.path.abspath(sys.argv[1]), sys.argv[2]
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path[:0] = [here, os.path.dirname(here), ref]
    os.chdir(ref)                                       # the reference reads configs/ relative to its root
    sys.argv = [os.path.join(ref, f"{which}.py")] + sys.argv[3:]
    if which == "train":
        import train                                    # the reference's module, unchanged
        from fddm_b200.adapter import SchedulerAdapter
        train.SchedulerAdapter = SchedulerAdapter
        train.main()
    else:
        runpy.run_path(sys.argv[0], run_name="__main__")


if __name__ == "__main__":
    main()
