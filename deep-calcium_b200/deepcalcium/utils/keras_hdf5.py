"""Keras-2.0.6 model files (HDF5) for the UNet2DS graph: read the weights of a model the reference saved
(``ModelCheckpoint`` / ``model.save``, unet_2d_summary.py:423-424; the released ``unet2ds_model.hdf5``, :28) and write
checkpoints in the same layout.  Replaces the ``h5py`` + ``keras.models.load_model`` part of
deepcalcium/utils/keras_helpers.py:24-68; the input-shape rewriting of that helper is unnecessary here because the
graph is fully convolutional.

Layout written by Keras 2.0.6 (``keras/engine/topology.py: save_weights_to_hdf5_group``, ``keras/models.py: save_model``):
  /            attrs  keras_version, backend, model_config (JSON), [training_config (JSON)]
  /model_weights            attrs  layer_names = [every layer of model.layers, in order], backend, keras_version
  /model_weights/<layer>    attrs  weight_names = [b'<layer>/kernel:0', ...]   (empty for layers without weights)
  /model_weights/<layer>/<layer>/kernel:0 ...   float32 datasets in Keras layouts (conv HWIO, transposed conv (kh,kw,Cout,Cin),
                            BatchNormalization gamma, beta, moving_mean, moving_variance)
h5py is used when it is importable, otherwise the pure-Python reader / writer in ``hdf5_lite``.
"""
import json

import numpy as np

from ..engine.graph import GraphSpec, list_to_weights

KERAS_VERSION = b'2.0.6'


def _h5():
    try:
        import h5py
        return h5py
    except ImportError:
        from . import hdf5_lite
        return hdf5_lite


def is_hdf5(path):
    with open(path, 'rb') as fp:
        head = fp.read(8 + 2048)
    sig = b'\x89HDF\r\n\x1a\n'
    return head[:8] == sig or head[512:520] == sig or head[1024:1032] == sig


def _text(v):
    if isinstance(v, bytes):
        return v.decode('utf8')
    if isinstance(v, np.ndarray) and v.shape == ():
        return _text(v[()])
    return str(v)


def _names(v):
    return [_text(x) for x in np.asarray(v).ravel().tolist()]


def keras_layer_names(spec):
    """[(keras layer name, our block name or None)] in ``model.layers`` order for the reference's unet()
    (unet_2d_summary.py:169-222): Keras numbers layers per class in creation order."""
    count = {}

    def name(cls):
        count[cls] = count.get(cls, 0) + 1
        return '%s_%d' % (cls, count[cls])

    out = [(name('input'), None), (name('lambda'), None)]

    def conv(block):
        out.extend([(name('conv2d'), block), (name('batch_normalization'), block), (name('activation'), None)])

    conv('enc0a'); conv('enc0b')
    for lvl in (1, 2, 3):
        out.append((name('max_pooling2d'), None))
        conv('enc%da' % lvl); conv('enc%db' % lvl)
        out.append((name('dropout'), None))
    out.append((name('max_pooling2d'), None))
    conv('bota'); conv('botb')
    for lvl in (3, 2, 1, 0):
        if spec.up_mode == 'transpose':
            out.extend([(name('conv2d_transpose'), 'up%d' % lvl), (name('batch_normalization'), 'up%d' % lvl),
                        (name('activation'), None)])
        else:
            out.append((name('up_sampling2d'), None))
        out.append((name('dropout'), None))
        out.append((name('concatenate'), None))
        conv('dec%da' % lvl); conv('dec%db' % lvl)
    out.append((name('conv2d'), 'head'))
    out.append((name('lambda'), None))
    return out


def read_keras_weights(path):
    """-> (GraphSpec, weights dict {block/param: float32 array}, info dict).  Works for any file written by Keras 2.x for
    the reference's unet(): weight-bearing layers are taken in ``layer_names`` order and classified by their arrays
    (4-D kernel + bias = convolution, four vectors = BatchNormalization), so it does not depend on the layer numbering."""
    h5 = _h5()
    with h5.File(path, 'r') as f:
        g = f['model_weights'] if 'model_weights' in f else f
        layer_names = _names(g.attrs['layer_names'])
        layers = []
        for ln in layer_names:
            lg = g[ln]
            wn = _names(lg.attrs['weight_names']) if 'weight_names' in lg.attrs else []
            if wn:
                layers.append((ln, [np.asarray(lg[w][...], dtype=np.float32) for w in wn], wn))
        info = {'keras_version': _text(f.attrs['keras_version']) if 'keras_version' in f.attrs else None,
                'backend': _text(f.attrs['backend']) if 'backend' in f.attrs else None,
                'layer_names': layer_names}
        if 'model_config' in f.attrs:
            try:
                info['model_config'] = json.loads(_text(f.attrs['model_config']))
            except ValueError:
                info['model_config'] = None
        if 'training_config' in f.attrs:
            try:
                info['training_config'] = json.loads(_text(f.attrs['training_config']))
            except ValueError:
                pass
    convs = [(ln, a) for ln, a, _ in layers if len(a) == 2 and a[0].ndim == 4]
    bns = [(ln, a) for ln, a, _ in layers if len(a) == 4 and all(x.ndim == 1 for x in a)]
    if not convs or convs[0][1][0].shape[:3] != (3, 3, 1):
        raise ValueError('%s: the first convolution is not a 3x3 kernel on one input channel - not a UNet2DS model' % path)
    nfb = int(convs[0][1][0].shape[3])
    transpose = any(a[0].shape[:2] == (2, 2) for _, a in convs)
    drp = 0.25
    cfg = info.get('model_config')
    if cfg and 'deepcalcium_b200' in cfg:
        drp = float(cfg['deepcalcium_b200'].get('prop_dropout_base', drp))
    elif cfg:
        rates = [l['config'].get('rate') for l in cfg.get('config', {}).get('layers', []) if l.get('class_name') == 'Dropout']
        if rates and rates[0] is not None:
            drp = float(rates[0])
    spec = GraphSpec(nfb, drp, 'transpose' if transpose else 'upsampling')
    n_conv_blocks = sum(1 for b in spec.blocks)
    n_bn_blocks = sum(1 for b in spec.blocks if b.kind != 'head')
    if len(convs) != n_conv_blocks or len(bns) != n_bn_blocks:
        raise ValueError('%s: %d convolutions / %d BatchNormalizations, the UNet2DS graph has %d / %d'
                         % (path, len(convs), len(bns), n_conv_blocks, n_bn_blocks))
    flat, ci, bi = [], 0, 0
    for blk in spec.blocks:
        flat.extend(convs[ci][1]); ci += 1
        if blk.kind != 'head':
            flat.extend(bns[bi][1]); bi += 1
    return spec, list_to_weights(spec, flat), info


def read_extra(path, names):
    """optional datasets this package adds under /deepcalcium_b200 (optimizer state); missing ones are skipped"""
    h5 = _h5()
    out = {}
    with h5.File(path, 'r') as f:
        for n in names:
            key = 'deepcalcium_b200/' + n
            if key in f:
                out[n] = np.asarray(f[key][...])
    return out


def write_keras_model(path, spec, weights, window_shape=(128, 128), optimizer=None, loss=None, extra_attrs=None,
                      extra_datasets=None):
    """Keras-2.0.6-layout model file: every layer of the reference graph under the names Keras would have given them,
    ``layer_names`` / ``weight_names`` attributes and float32 datasets, so ``keras_model.load_weights(path)`` (and this
    module's reader) can load it.  ``model_config`` lists the layers (class, name, config) but not the two Lambda
    functions' bytecode, which only the Keras process that built the graph can serialise."""
    h5 = _h5()
    names = keras_layer_names(spec)
    wkeys = {'conv': ['kernel', 'bias'], 'bn': ['gamma', 'beta', 'moving_mean', 'moving_variance']}
    ours = {'moving_variance': 'moving_var'}
    layers_cfg = []
    with h5.File(path, 'w') as f:
        f.attrs['keras_version'] = KERAS_VERSION
        f.attrs['backend'] = b'tensorflow'
        g = f.create_group('model_weights')
        g.attrs['layer_names'] = np.array([n.encode('utf8') for n, _ in names])
        g.attrs['backend'] = b'tensorflow'
        g.attrs['keras_version'] = KERAS_VERSION
        for lname, block in names:
            lg = g.create_group(lname)
            cls = lname.rsplit('_', 1)[0]
            kind = 'conv' if cls in ('conv2d', 'conv2d_transpose') else ('bn' if cls == 'batch_normalization' else None)
            wn = []
            if kind:
                for k in wkeys[kind]:
                    wname = '%s/%s:0' % (lname, k)
                    wn.append(wname.encode('utf8'))
                    lg.create_dataset(wname, data=np.asarray(weights['%s/%s' % (block, ours.get(k, k))], dtype=np.float32))
            lg.attrs['weight_names'] = np.array(wn) if wn else np.zeros((0,), dtype='S1')
            layers_cfg.append({'class_name': ''.join(p.capitalize() for p in cls.split('_')).replace('2d', '2D'),
                               'name': lname, 'config': {'name': lname}})
        layers_cfg[0]['config']['batch_input_shape'] = [None, int(window_shape[0]), int(window_shape[1])]
        cfg = {'class_name': 'Model', 'config': {'name': 'model_1', 'layers': layers_cfg},
               'deepcalcium_b200': {'nb_filters_base': spec.nfb, 'prop_dropout_base': spec.drp,
                                    'upsampling_or_transpose': spec.up_mode, 'window_shape': list(window_shape)}}
        f.attrs['model_config'] = json.dumps(cfg).encode('utf8')
        if optimizer is not None:
            f.attrs['training_config'] = json.dumps({'optimizer_config': {'class_name': 'Adam', 'config': optimizer},
                                                     'loss': loss}).encode('utf8')
        for k, v in (extra_attrs or {}).items():
            f.attrs[k] = v
        for k, v in (extra_datasets or {}).items():
            f.create_dataset('deepcalcium_b200/' + k, data=np.asarray(v))
    return path
