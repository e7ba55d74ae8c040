"""NoBRS predictor plumbing around the per-click forward (host side, Python like the reference's).

Mirrors the reference's isegm/inference package for this path -- same class and function names, argument meaning
and results -- so `evaluate_vpumodel.py`-style drivers work unchanged on top of the B200 model:
  clicker.py      Click, Clicker                                   (reference inference/clicker.py)
  transforms.py   ZoomIn, AddHorizontalFlip, SigmoidForPred        (reference inference/transforms/*.py)
  predictor.py    BasePredictor (NoBRS), get_predictor             (reference inference/predictors/base.py, __init__.py)
  evaluation.py   evaluate_sample / evaluate_dataset, get_iou, compute_noc_metric, and the additions of this
                  repo: lock-step batched evaluation and rank sharding with an all_gather of IoU tallies
  datasets.py     synthetic ellipse dataset (the reference's isegm/data was never published)
"""
from .clicker import Click, Clicker  # noqa: F401
from .evaluation import (compute_noc_metric, evaluate_dataset, evaluate_lockstep, evaluate_sample, gather_iou_tables,  # noqa: F401
                         get_iou, shard_range)
from .predictor import BasePredictor, get_predictor  # noqa: F401
from .transforms import AddHorizontalFlip, SigmoidForPred, ZoomIn  # noqa: F401
