"""Command line of the UNet2DS neurofinder workflow - the counterpart of the reference's examples/neurons/unet2ds_nf.py
(:99-144): the same three actions with the same positional / optional arguments, running on the B200 kernels.

    python examples/neurons/unet2ds_nf.py train    <dataset_name> [-m MODEL] [-c CHECKPOINTS_DIR]
    python examples/neurons/unet2ds_nf.py evaluate <dataset_name>  -m MODEL  [-c CHECKPOINTS_DIR]
    python examples/neurons/unet2ds_nf.py predict  <dataset_name>  -m MODEL  [-c CHECKPOINTS_DIR]
    python examples/neurons/unet2ds_nf.py ingest   <dataset_name>            [-d DATASETS_DIR]      (new: TIFF frames -> dataset.hdf5)

<dataset_name>: a neurofinder name, a comma-separated list, or all / all_train / all_test (datasets/nf.py:37-56).
Datasets must already be on disk (download needs network access): `ingest` turns an unzipped neurofinder directory
(images/*.tiff [+ regions/regions.json]) into the reference's dataset.hdf5 through the streaming GPU projection.
MODEL: a Keras-2.x HDF5 model file of the reference (e.g. the released unet2ds_model.hdf5) or a checkpoint of this package.
"""
import argparse
import logging
import os
import sys
from time import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import numpy as np  # noqa: E402

from deepcalcium.datasets.nf import nf_load_hdf5, nf_submit, nf_ingest  # noqa: E402
from deepcalcium.models.neurons.unet_2d_summary import UNet2DSummary  # noqa: E402
from deepcalcium.utils.config import CHECKPOINTS_DIR, DATASETS_DIR  # noqa: E402
from deepcalcium.utils.runtime import funcname  # noqa: E402

DEFAULT_CPDIR = '%s/neurons_unet2ds_nf' % CHECKPOINTS_DIR


def training(dataset_name, model_path, checkpoints_dir, datasets_dir=None):
    """unet2ds_nf.py:23-44 - the canonical training configuration of the reference: 128x128 crops, batch 20,
    100 steps x 10 epochs, top 75 % of the rows for training and the bottom 25 % for validation"""
    paths = nf_load_hdf5(dataset_name, datasets_dir)
    return UNet2DSummary(cpdir=checkpoints_dir).fit(
        paths, model_path=model_path, shape_trn=(128, 128), shape_val=(512, 512), batch_size_trn=20, nb_steps_trn=100,
        nb_epochs=10, keras_callbacks=[], prop_trn=0.75, prop_val=0.25)


def evaluation(dataset_name, model_path, checkpoints_dir, datasets_dir=None):
    """unet2ds_nf.py:47-64 - neurofinder scores with and without test-time augmentation"""
    log = logging.getLogger(funcname())
    paths = nf_load_hdf5(dataset_name, datasets_dir)
    model = UNet2DSummary(cpdir=checkpoints_dir)
    for tta in (True, False):
        log.info('Evaluation with%s.' % (' TTA' if tta else 'out TTA'))
        model.predict(paths, model_path=model_path, window_shape=(512, 512), save=True, print_scores=True, augmentation=tta)


def prediction(dataset_name, model_path, checkpoints_dir, datasets_dir=None):
    """unet2ds_nf.py:67-96 - masks and neurofinder submission files with and without test-time augmentation"""
    log = logging.getLogger(funcname())
    paths = nf_load_hdf5(dataset_name, datasets_dir)
    model = UNet2DSummary(cpdir=checkpoints_dir)
    stamp = int(time())
    for tta in (True, False):
        log.info('Prediction with%s.' % (' TTA' if tta else 'out TTA'))
        masks, names = model.predict(paths, model_path=model_path, window_shape=(512, 512), save=False, augmentation=tta)
        masks = [m.round() for m in masks]
        suffix = '_TTA' if tta else ''
        for json_path in ('%s/submission_%d%s.json' % (model.cpdir, stamp, suffix),
                          '%s/submission_latest%s.json' % (model.cpdir, suffix)):
            nf_submit(masks, names, json_path)


def ingest(dataset_name, model_path=None, checkpoints_dir=None, datasets_dir=None):
    """datasets/nf.py:99-148 for directories that are already unzipped: TIFF frames -> series/mean, series/max (+ masks)"""
    log = logging.getLogger(funcname())
    root = datasets_dir or '%s/neurons_nf' % DATASETS_DIR
    for name in dataset_name.split(','):
        log.info('%s -> %s' % (name, nf_ingest(name, root)))


ACTIONS = {'train': (training, 'all_train', False), 'evaluate': (evaluation, 'all_train', True),
           'predict': (prediction, 'all', True), 'ingest': (ingest, None, False)}


def main(argv=None):
    np.random.seed(865)                 # the reference seeds numpy (and TF) at import, unet2ds_nf.py:18-19
    logging.basicConfig(level=logging.INFO)
    ap = argparse.ArgumentParser(description='CLI for UNet2DS model.')
    sp = ap.add_subparsers(title='actions', description='Choose an action.', dest='which')
    sp.required = True
    for action, (_, default_ds, model_required) in ACTIONS.items():
        p = sp.add_parser(action, help='CLI for %s.' % action)
        p.add_argument('dataset_name', help='dataset name', type=str, **({} if default_ds is None else {'nargs': '?', 'default': default_ds}))
        if action != 'ingest':
            p.add_argument('-m', '--model_path', help='path to model', required=model_required)
            p.add_argument('-c', '--checkpoints_dir', help='checkpoint directory', default=DEFAULT_CPDIR)
        p.add_argument('-d', '--datasets_dir', help='directory holding <name>/dataset.hdf5 (default: the configured datasets_dir)')
    args = vars(ap.parse_args(argv))
    fn = ACTIONS[args.pop('which')][0]
    return fn(**args)


if __name__ == '__main__':
    main()
