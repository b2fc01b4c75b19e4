"""Tensor construction helpers with the reference's names (ccsmeth/utils/constants_torch.py:5-16)."""
import torch

use_cuda = torch.cuda.is_available()


def FloatTensor(tensor, device=0):
    if use_cuda:
        return torch.tensor(tensor, dtype=torch.float, device='cuda:{}'.format(device))
    return torch.tensor(tensor, dtype=torch.float)


def FloatTensor_cpu(tensor):
    return torch.tensor(tensor, dtype=torch.float)
