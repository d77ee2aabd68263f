"""SpecAugmentation stand-in: the reference only applies it when `self.training` (htsat.py:903-904); identity here."""
import torch.nn as nn


class SpecAugmentation(nn.Module):
    def __init__(self, **kwargs):
        super().__init__()

    def forward(self, x):
        return x
