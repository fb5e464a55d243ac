"""``MarginLoss`` with the reference's signature (torchkge/utils/losses.py:12-44)."""
from torch import nn

from .training import margin_loss


class MarginLoss(nn.Module):
    """sum_i max(0, margin - f(pos_i) + f(neg_i)): ``MarginRankingLoss(margin, reduction='sum')``
    with target +1, computed by a CUDA reduction kernel with its own backward."""

    def __init__(self, margin):
        super().__init__()
        self.margin = float(margin)

    def forward(self, positive_triplets, negative_triplets):
        return margin_loss(positive_triplets, negative_triplets, self.margin)
