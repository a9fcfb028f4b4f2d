"""Mirror of the reference's engine.criterion (engine.py:10-15) -- what train_one_epoch calls every step.

    loss = criterion(logits, lbl, loss_weight, num_classes=args.nb_classes, dice=args.dice, ignore_index=args.ignore_index)

`inputs` may be the head's low-resolution logits (the upsample of models/build_models.py:65 is fused), or full-resolution
ones as in the unmodified reference.
"""
from __future__ import annotations

from .functional import fused_criterion


def criterion(inputs, target, loss_weight=None, num_classes: int = 2, dice: bool = True, ignore_index: int = -100):
    if inputs.shape[1] != num_classes:
        raise ValueError(f"inputs have {inputs.shape[1]} channels, num_classes = {num_classes}")   # one_hot would raise
    return fused_criterion(inputs, target, loss_weight, dice, ignore_index)
