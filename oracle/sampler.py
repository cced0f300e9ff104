"""Oracle: the reference's training crop sampler on the host.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Checker for the product's device sampler
(``UNet2DSummary._crop_descriptors`` + ``dcb_crop_batch``), never used by the product.

Follows deepcalcium/models/neurons/unet_2d_summary.py:434-530 (``_batch_gen``): per crop it draws, from the
GLOBAL numpy RNG and in this order, (1) a dataset (``choice`` with probabilities), (2) a neuron pixel inside the
allowed row range, (3) a row jitter and (4) a column jitter in [-5, 5), (5) the number of augmentations in
[0, nb_max_augment], (6) that many picks from the six flips / rot90s; the window is clamped to the row range and
the image, copied into the top-left corner of a zero window, and the picked augmentations are applied in order.
"""
import os
import pickle

import numpy as np

# the six augmentations of the reference table (:457-464) as (kind, argument)
AUGMENTATIONS = (('id', 0), ('flip', 1), ('flip', 0), ('rot', 1), ('rot', 2), ('rot', 3))


def apply_augmentation(k, a):
    kind, arg = AUGMENTATIONS[k]
    if kind == 'flip':
        return np.flip(a, axis=arg)
    if kind == 'rot':
        return np.rot90(a, arg)
    return a


def neuron_pixels(mask, row_range):
    """(row, col) of every mask pixel inside the row range, rows relative to the range start (:467-470)"""
    lo, hi = row_range
    rr, cc = np.nonzero(mask[lo:hi, :] == 1)
    return list(zip(rr, cc))


def window_bounds(center, jitter, row_range, width, window):
    """clamped window [y0, y1) x [x0, x1) around a jittered centre (:505-510)"""
    (cy, cx), (jy, jx), (lo, hi), (hw, ww) = center, jitter, row_range, window
    cy = min(max(lo, cy + jy), hi)
    cx = min(max(0, cx + jx), width)
    y0 = max(lo, int(cy - (hw / 2)))
    x0 = max(0, int(cx - (ww / 2)))
    return y0, min(y0 + hw, hi), x0, min(x0 + ww, width)


def host_batches(S_summ, M_summ, names, y_coords, batch_size, nb_steps, window_shape, nb_max_augment=0,
                 scores_path=None):
    """generator of (float32 [B,h,w], uint8 [B,h,w]) batches, same stream as the reference's _batch_gen"""
    rng = np.random                          # the reference samples from the global numpy RNG (:455)
    hw, ww = window_shape
    locs = [neuron_pixels(m, yc) for m, yc in zip(M_summ, y_coords)]
    n_ds = len(S_summ)
    probs = np.ones(n_ds) / n_ds
    yields = 0
    while True:
        if scores_path and os.path.exists(scores_path) and (yields - 1) % nb_steps == 0:      # :482-489
            with open(scores_path, 'rb') as fp:
                table = pickle.load(fp)
            probs = np.array([1 - np.mean(table[n]) for n in names])
            probs /= probs.sum()
        xs = np.zeros((batch_size, hw, ww), dtype=np.float32)
        ys = np.zeros((batch_size, hw, ww), dtype=np.uint8)
        for b in range(batch_size):
            d = rng.choice(np.arange(n_ds), p=probs)
            center = locs[d][rng.randint(0, len(locs[d]))]
            jitter = (rng.randint(-5, 5), rng.randint(-5, 5))
            y0, y1, x0, x1 = window_bounds(center, jitter, y_coords[d], S_summ[d].shape[1], (hw, ww))
            xs[b, :y1 - y0, :x1 - x0] = S_summ[d][y0:y1, x0:x1]
            ys[b, :y1 - y0, :x1 - x0] = M_summ[d][y0:y1, x0:x1]
            for k in rng.choice(len(AUGMENTATIONS), rng.randint(0, nb_max_augment + 1)):
                xs[b], ys[b] = apply_augmentation(int(k), xs[b]), apply_augmentation(int(k), ys[b])
        yields += 1
        yield xs, ys
