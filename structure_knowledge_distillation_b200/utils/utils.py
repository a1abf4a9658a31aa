"""Affinity helpers with the reference's names (utils/utils.py:170-183), backed by the pair-wise CUDA kernels."""
import torch

from .. import functions as Fn


def sim_dis_compute(f_S, f_T):
    """sum((A_T - A_S)^2) / nodes^2 / N on already-pooled features (utils/utils.py:180-183).  The affinity matrices are
    never returned to Python: pooling with a 1x1 window is the identity, the rest is the fused gram/L2 kernel."""
    return Fn.PairWiseLoss.apply(f_S, f_T.detach(), 1, 1)


def load_state_dict_compat(model, state_dict, strict=True):
    """Loads reference checkpoints: parameter names/shapes are the reference's."""
    return model.load_state_dict(state_dict, strict=strict)
