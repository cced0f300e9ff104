"""Generate the committed known-answer fixtures under tests/golden/.

TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED: the reference ships no golden vectors and cannot be
run here (Keras 2.0.6 / TF 1.2.1 absent), so these vectors come from the oracle restatement
itself (float64).  They pin the oracle against accidental change and give the GPU tests a
fixture that does not depend on recomputing the oracle.

    python -m oracle.make_golden
"""
import os

import numpy as np
import torch

from . import (UNetSpec, init_weights, unet_forward, tta_predict, train_step, project_mean_max,
               summarize_series)

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')


def main():
    os.makedirs(OUT, exist_ok=True)
    # --- projection: small int16-like movie
    rng = np.random.default_rng(7535)
    movie = (rng.random((37, 24, 40), dtype=np.float32) * 4096).astype(np.float32)
    mean, mx = project_mean_max(movie)
    np.savez_compressed(os.path.join(OUT, 'projection_small.npz'), movie=movie, mean=mean, max=mx,
                        summary=summarize_series(mean.astype(np.float16)))
    # --- tiny U-Net (nb_filters_base=4): forward, TTA, one training step per loss
    spec = UNetSpec(nb_filters_base=4)
    w = init_weights(spec, seed=7535)
    x = np.random.default_rng(865).standard_normal((2, 32, 32)).astype(np.float32)
    y = (np.random.default_rng(866).random((2, 32, 32)) < 0.126).astype(np.uint8)
    fwd = unet_forward(w, x, spec, dtype=torch.float64)
    s = np.random.default_rng(3).standard_normal((27, 30)).astype(np.float32)
    mask, act = tta_predict(w, s, spec, window=32, dtype=torch.float64)
    d = {'x': x, 'y': y, 'logit': fwd['logit'].numpy(), 'prob': fwd['prob'].numpy(), 's': s, 'tta_mask': mask,
         'tta_act': act}
    for k, v in w.items():
        d['w:' + k] = v
    for loss in ('dice_loss', 'binary_crossentropy', 'dicesq_loss', 'weighted_binary_crossentropy'):
        L, nw, st, g, out = train_step(w, x, y, spec=spec, loss=loss)
        d['%s:loss' % loss] = np.float64(L)
        for k in ('enc0a/kernel', 'botb/kernel', 'up2/kernel', 'dec0b/gamma', 'head/kernel', 'head/bias'):
            d['%s:grad:%s' % (loss, k)] = g[k]
            d['%s:new:%s' % (loss, k)] = nw[k]
        d['%s:new:enc0a/moving_mean' % loss] = nw['enc0a/moving_mean']
        d['%s:new:up1/moving_var' % loss] = nw['up1/moving_var']
    np.savez_compressed(os.path.join(OUT, 'unet_nfb4_32.npz'), **d)
    print('wrote', os.listdir(OUT))


if __name__ == '__main__':
    main()
