"""Mirror of the reference's ``MaskedLoss`` (loss.py:6-25): sum-reduced loss over ``pred*mask``
divided by the number of real utterances (or the summed class weights)."""
import torch
import torch.nn as nn


class MaskedLoss(nn.Module):
    def __init__(self, losser, weight=None):
        super().__init__()
        self.weight = weight
        self.loss = losser(weight=weight, reduction="sum")

    def forward(self, pred, target, mask):
        flat = mask.reshape(-1, 1)
        total = self.loss(pred * flat, target)
        if self.weight is None:
            return total / mask.sum()
        return total / (self.weight[target] * flat.squeeze(1)).sum()
