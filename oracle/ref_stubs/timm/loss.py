import torch.nn as nn


class LabelSmoothingCrossEntropy(nn.CrossEntropyLoss):
    def __init__(self, smoothing=0.1):
        super().__init__(label_smoothing=smoothing)
