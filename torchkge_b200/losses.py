"""``MarginLoss``, ``LogisticLoss`` and ``BinaryCrossEntropyLoss`` with the reference's
signatures (torchkge/utils/losses.py:12-112)."""
from torch import nn

from . import _lib
from .training import margin_loss, pair_loss


class MarginLoss(nn.Module):
    """sum_i max(0, margin - f(pos_i) + f(neg_i)): ``MarginRankingLoss(margin, reduction='sum')``
    with target +1, computed by a CUDA reduction kernel with its own backward."""

    def __init__(self, margin):
        super().__init__()
        self.margin = float(margin)

    def forward(self, positive_triplets, negative_triplets):
        return margin_loss(positive_triplets, negative_triplets, self.margin)


class LogisticLoss(nn.Module):
    """sum_i log(1 + exp(-f(pos_i))) + log(1 + exp(f(neg_i))): ``SoftMarginLoss(reduction='sum')``
    on (pos, +1) and (neg, -1) (utils/losses.py:47-78), one CUDA reduction with its own backward."""

    def forward(self, positive_triplets, negative_triplets):
        return pair_loss(positive_triplets, negative_triplets, _lib.LOSS_LOGISTIC)


class BinaryCrossEntropyLoss(nn.Module):
    """``BCELoss(reduction='sum')`` of sigmoid(pos) against 1 plus sigmoid(neg) against 0
    (utils/losses.py:81-112)."""

    def forward(self, positive_triplets, negative_triplets):
        return pair_loss(positive_triplets, negative_triplets, _lib.LOSS_BCE)
