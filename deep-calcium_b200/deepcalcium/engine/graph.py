"""Static description of the UNet2DS graph (deepcalcium/models/neurons/unet_2d_summary.py:123-224
of the reference) and the Keras-layout weight container.

Layer names are ours; the ORDER and per-layer array order reproduce Keras'
``model.get_weights()`` for the reference graph: per layer trainable then
non-trainable arrays, i.e. conv [kernel, bias], BN [gamma, beta, moving_mean,
moving_variance] -> 134 arrays for the default (transpose) graph.
"""
from collections import OrderedDict

import numpy as np

BN_EPS = 1e-3            # keras.layers.BatchNormalization default epsilon
BN_MOMENTUM_CONV = 0.99  # unet_2d_summary.py:165 (Keras default)
BN_MOMENTUM_UP = 0.5     # unet_2d_summary.py:157

PARAMS = {'conv': ('kernel', 'bias', 'gamma', 'beta', 'moving_mean', 'moving_var'),
          'up': ('kernel', 'bias', 'gamma', 'beta', 'moving_mean', 'moving_var'),
          'head': ('kernel', 'bias')}
TRAINABLE = ('kernel', 'bias', 'gamma', 'beta')


class Block(object):
    __slots__ = ('name', 'kind', 'cin', 'cout', 'level')

    def __init__(self, name, kind, cin, cout, level):
        self.name, self.kind, self.cin, self.cout, self.level = name, kind, cin, cout, level

    def kernel_shape(self):
        if self.kind == 'conv':
            return (3, 3, self.cin, self.cout)       # Keras Conv2D: HWIO
        if self.kind == 'up':
            return (2, 2, self.cout, self.cin)       # Keras Conv2DTranspose: (kh, kw, Cout, Cin)
        return (1, 1, self.cin, self.cout)

    def param_shapes(self):
        shp = OrderedDict(kernel=self.kernel_shape(), bias=(self.cout,))
        if self.kind != 'head':
            for p in ('gamma', 'beta', 'moving_mean', 'moving_var'):
                shp[p] = (self.cout,)
        return shp


class GraphSpec(object):
    """Arguments of the reference's ``unet()`` (unet_2d_summary.py:123-124)."""

    def __init__(self, nb_filters_base=32, prop_dropout_base=0.25, upsampling_or_transpose='transpose'):
        assert upsampling_or_transpose in ('transpose', 'upsampling')
        self.nfb = int(nb_filters_base)
        self.drp = float(prop_dropout_base)
        self.up_mode = upsampling_or_transpose
        n = self.nfb
        b = [Block('enc0a', 'conv', 1, n, 0), Block('enc0b', 'conv', n, n, 0)]
        for lvl in (1, 2, 3):
            w = n << lvl
            b += [Block('enc%da' % lvl, 'conv', w // 2, w, lvl), Block('enc%db' % lvl, 'conv', w, w, lvl)]
        b += [Block('bota', 'conv', 8 * n, 16 * n, 4), Block('botb', 'conv', 16 * n, 16 * n, 4)]
        c = 16 * n
        for lvl in (3, 2, 1, 0):
            w = n << lvl
            if self.up_mode == 'transpose':
                b.append(Block('up%d' % lvl, 'up', c, w, lvl))
                cat = 2 * w
            else:
                cat = c + w
            b += [Block('dec%da' % lvl, 'conv', cat, w, lvl), Block('dec%db' % lvl, 'conv', w, w, lvl)]
            c = w
        b.append(Block('head', 'head', n, 2, 0))
        self.blocks = b
        self.by_name = OrderedDict((x.name, x) for x in b)

    def dropout_after(self):
        """{tensor name: drop probability}, unet_2d_summary.py:179,185,191,198,204,210,216."""
        d = self.drp
        return OrderedDict([('enc1b', d), ('enc2b', 2 * d), ('enc3b', 2 * d),
                            ('up3', 2 * d), ('up2', 2 * d), ('up1', 2 * d), ('up0', d)])

    def weight_keys(self):
        return ['%s/%s' % (blk.name, p) for blk in self.blocks for p in PARAMS[blk.kind]]

    def weight_shapes(self):
        return OrderedDict(('%s/%s' % (blk.name, p), shp) for blk in self.blocks
                           for p, shp in blk.param_shapes().items())

    def flops_forward(self, H, W):
        """MAC*2 of conv + convT + head for one H x W image."""
        f = 0
        for blk in self.blocks:
            h, w = H >> blk.level, W >> blk.level
            if blk.kind == 'conv':
                f += 2 * h * w * 9 * blk.cin * blk.cout
            elif blk.kind == 'up':      # input is at level+1 resolution, 4 output pixels per input pixel
                f += 2 * (h // 2) * (w // 2) * 4 * blk.cin * blk.cout
            else:
                f += 2 * h * w * blk.cin * blk.cout
        return f

    def flops_train(self, H, W):
        """fwd + dgrad + wgrad GEMM flops actually executed per image (no dgrad for the first layer)."""
        f = 0
        for blk in self.blocks:
            h, w = H >> blk.level, W >> blk.level
            if blk.kind == 'conv':
                g = 2 * h * w * 9 * blk.cin * blk.cout
                f += g * (2 if blk.name == 'enc0a' else 3)
            elif blk.kind == 'up':
                f += 3 * 2 * (h // 2) * (w // 2) * 4 * blk.cin * blk.cout
            else:
                f += 3 * 2 * h * w * blk.cin * blk.cout
        return f


def _truncated_normal(rng, shape):
    """standard normal re-drawn where |z| > 2 (tf.truncated_normal, which Keras 2.0.6's VarianceScaling(distribution=
    'normal') - i.e. he_normal - samples from: the effective standard deviation is ~0.88 of the nominal one)"""
    z = rng.standard_normal(shape)
    bad = np.abs(z) > 2
    while bad.any():
        z[bad] = rng.standard_normal(int(bad.sum()))
        bad = np.abs(z) > 2
    return z


def he_normal_weights(spec, seed=None):
    """Fresh Keras-style initialisation: he_normal kernels (truncated normal, stddev sqrt(2 / fan_in) with fan_in =
    kh*kw*shape[-2], the Keras rule, which is 4*Cout for the transposed kernels), zero biases, BN gamma=1 beta=0
    mean=0 var=1."""
    rng = np.random.RandomState(seed)
    w = OrderedDict()
    for blk in spec.blocks:
        ks = blk.kernel_shape()
        fan_in = ks[0] * ks[1] * ks[2]
        if blk.kind == 'head':      # Keras default for Conv2D(2, 1): glorot_uniform
            limit = np.sqrt(6. / (ks[2] + ks[3]))
            w[blk.name + '/kernel'] = rng.uniform(-limit, limit, ks).astype(np.float32)
        else:
            w[blk.name + '/kernel'] = (_truncated_normal(rng, ks) * np.sqrt(2. / fan_in)).astype(np.float32)
        w[blk.name + '/bias'] = np.zeros(blk.cout, np.float32)
        if blk.kind != 'head':
            w[blk.name + '/gamma'] = np.ones(blk.cout, np.float32)
            w[blk.name + '/beta'] = np.zeros(blk.cout, np.float32)
            w[blk.name + '/moving_mean'] = np.zeros(blk.cout, np.float32)
            w[blk.name + '/moving_var'] = np.ones(blk.cout, np.float32)
    return w


def weights_to_list(spec, w):
    return [np.asarray(w[k]) for k in spec.weight_keys()]


def list_to_weights(spec, lst):
    keys = spec.weight_keys()
    if len(keys) != len(lst):
        raise ValueError('expected %d weight arrays, got %d' % (len(keys), len(lst)))
    shapes = spec.weight_shapes()
    out = OrderedDict()
    for k, a in zip(keys, lst):
        a = np.asarray(a, dtype=np.float32)
        if tuple(a.shape) != tuple(shapes[k]):
            raise ValueError('weight %s: expected shape %s, got %s' % (k, shapes[k], a.shape))
        out[k] = a
    return out
