import pytest
import torch

requires_gpu = pytest.mark.gpu


def dev():
    assert torch.cuda.is_available(), "gpu test selected without a GPU"
    return "cuda:0"
