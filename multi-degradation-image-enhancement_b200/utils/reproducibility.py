"""Seeds + cuDNN determinism flags (reference utils/reproducibility.py:6-24).  The native kernels are
deterministic by construction (fixed-order reductions, no float atomics)."""
import random

import numpy as np
import torch


def set_seed_and_cudnn(seed_value=42):
    random.seed(seed_value)
    np.random.seed(seed_value)
    torch.manual_seed(seed_value)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed_value)
    torch.backends.cudnn.deterministic = True
    torch.backends.cudnn.benchmark = False
    torch.backends.cudnn.enabled = True
