"""Oracle: losses, per-batch metrics and the TTA table.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  PARITY UNPINNED.

Follows deepcalcium/utils/neurons.py:13-137 with keras.backend ops restated in
torch (so autograd gives the reference gradients).  Keras reduces every loss as
``mean(loss_fn(y_true, y_pred))``; ``K.epsilon()`` is 1e-7 in Keras 2.0.6.
"""
import numpy as np
import torch

K_EPS = 1e-7


def dice_loss(yt, yp):
    """utils/neurons.py:78-83: one scalar over the whole batch."""
    inter = torch.sum(yt * yp)
    dsmooth = (2. * inter) / (torch.sum(yt) + torch.sum(yp) + 1e-7)
    return 1 - dsmooth


def dicesq_loss(yt, yp):
    """utils/neurons.py:86-94."""
    nmr = 2 * torch.sum(yt * yp)
    dnm = torch.sum(yt ** 2) + torch.sum(yp ** 2) + K_EPS
    return -1 * (nmr / dnm)


def binary_crossentropy(yt, yp):
    """keras.losses.binary_crossentropy (unet_2d_summary.py:9,374): probabilities
    clipped to [eps, 1-eps], per-pixel BCE, mean over the last axis, then Keras
    takes the mean of that -> overall mean."""
    p = torch.clamp(yp, K_EPS, 1 - K_EPS)
    return torch.mean(-(yt * torch.log(p) + (1 - yt) * torch.log(1 - p)))


def weighted_binary_crossentropy(yt, yp, weightpos=2., weightneg=1.):
    """utils/neurons.py:13-29; Keras then takes the mean of the returned map."""
    losspos = yt * torch.log(yp + 1e-7)
    lossneg = (1 - yt) * torch.log(1 - yp + 1e-7)
    return torch.mean(-1 * ((weightpos * losspos) + (weightneg * lossneg)))


LOSSES = {
    'binary_crossentropy': binary_crossentropy,
    'weighted_binary_crossentropy': weighted_binary_crossentropy,
    'dice_loss': dice_loss,
    'dicesq_loss': dicesq_loss,
}


def batch_metrics(yt, yp):
    """The seven per-batch metrics compiled at unet_2d_summary.py:398-399
    (F1, prec, reca, dice, dicesq, posyt, posyp; utils/neurons.py:32-106).
    ``K.round`` is round-half-to-even (tf.round)."""
    yt = torch.as_tensor(yt, dtype=torch.float64)
    yp = torch.as_tensor(yp, dtype=torch.float64)
    ypr = torch.round(yp)
    tp = torch.sum(ypr * yt)
    prec = tp / (torch.sum(ypr) + K_EPS)
    fn = torch.sum(torch.clamp(yt - ypr, 0, 1))
    reca = tp / (tp + fn + K_EPS)
    f1 = (2 * prec * reca) / (prec + reca + K_EPS)
    dice = (2. * tp) / (torch.sum(yt) + torch.sum(ypr) + 1e-7)
    dicesq = 2 * torch.sum(yt * yp) / (torch.sum(yt ** 2) + torch.sum(yp ** 2) + K_EPS)
    size = float(yt.numel())
    posyt = torch.sum(yt) / (size + K_EPS)
    posyp = torch.sum(ypr) / (size + K_EPS)
    return {k: float(v) for k, v in dict(F1=f1, prec=prec, reca=reca, dice=dice, dicesq=dicesq,
                                         posyt=posyt, posyp=posyp).items()}


# utils/neurons.py:112-137, restated on numpy arrays of shape [N,H,W].
INVERTIBLE_2D_AUGMENTATIONS = [
    ('identity', lambda x: x, lambda x: x),
    ('vflip', lambda x: x[:, ::-1, ...], lambda x: x[:, ::-1, ...]),
    ('hflip', lambda x: x[:, :, ::-1], lambda x: x[:, :, ::-1]),
    ('rot90', lambda x: np.rot90(x, 1, axes=(1, 2)), lambda x: np.rot90(x, -1, axes=(1, 2))),
    ('rot180', lambda x: np.rot90(x, 2, axes=(1, 2)), lambda x: np.rot90(x, -2, axes=(1, 2))),
    ('rot270', lambda x: np.rot90(x, 3, axes=(1, 2)), lambda x: np.rot90(x, -3, axes=(1, 2))),
    ('rot90vflip', lambda x: np.rot90(x, 1, axes=(1, 2))[:, ::-1, ...],
     lambda x: np.rot90(x, 1, axes=(1, 2))[:, ::-1, ...]),
    ('rot90hflip', lambda x: np.rot90(x, 1, axes=(1, 2))[:, :, ::-1],
     lambda x: np.rot90(x, 1, axes=(1, 2))[:, :, ::-1]),
]
