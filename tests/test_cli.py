"""Command line (examples/neurons/unet2ds_nf.py, counterpart of the reference file of the same name :99-144) and the
outlined-figure helper - host logic only."""
import importlib.util
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cli():
    spec = importlib.util.spec_from_file_location('unet2ds_nf_cli', os.path.join(ROOT, 'deep-calcium_b200', 'examples', 'neurons', 'unet2ds_nf.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_cli_actions_and_arguments(monkeypatch):
    cli = _cli()
    calls = []
    for action in ('train', 'evaluate', 'predict', 'ingest'):
        fn, default_ds, needs_model = cli.ACTIONS[action]
        monkeypatch.setitem(cli.ACTIONS, action, ((lambda a: (lambda **kw: calls.append((a, kw))))(action), default_ds, needs_model))
    cli.main(['train', 'neurofinder.00.00', '-c', '/tmp/cp'])
    cli.main(['train'])                                                   # reference default: all_train
    cli.main(['evaluate', 'all_train', '-m', 'model.hdf5'])
    cli.main(['predict', '-m', 'model.hdf5', '-d', '/data'])              # reference default: all
    cli.main(['ingest', 'neurofinder.00.00,neurofinder.00.01'])
    assert calls[0] == ('train', dict(dataset_name='neurofinder.00.00', model_path=None, checkpoints_dir='/tmp/cp', datasets_dir=None))
    assert calls[1][1]['dataset_name'] == 'all_train' and calls[1][1]['checkpoints_dir'].endswith('neurons_unet2ds_nf')
    assert calls[2][1]['model_path'] == 'model.hdf5'
    assert calls[3] == ('predict', dict(dataset_name='all', model_path='model.hdf5', checkpoints_dir=cli.DEFAULT_CPDIR, datasets_dir='/data'))
    assert calls[4][1] == dict(dataset_name='neurofinder.00.00,neurofinder.00.01', datasets_dir=None)
    import pytest
    with pytest.raises(SystemExit):
        cli.main(['evaluate', 'all_train'])                               # -m is required (unet2ds_nf.py:116-117)


def test_mask_outlines():
    from deepcalcium.utils.neurons import mask_outlines
    img = np.linspace(0, 10, 20 * 30, dtype=np.float32).reshape(20, 30)
    m = np.zeros((20, 30)); m[5:10, 5:12] = 1
    mp = np.zeros((20, 30)); mp[12:16, 20:25] = 1
    out = mask_outlines(img, [m, mp, np.zeros((20, 30))], ['blue', 'red', 'green'])
    assert out.shape == (20, 30, 3) and out.dtype == np.uint8
    assert tuple(out[5, 5]) == (0, 0, 255) and tuple(out[9, 11]) == (0, 0, 255)       # boundary of the truth mask
    assert tuple(out[12, 20]) == (255, 0, 0)                                            # boundary of the prediction
    assert out[7, 8, 0] == out[7, 8, 1] == out[7, 8, 2]                                 # interior keeps the grey image
    assert out[0, 0, 0] == 0 and out[19, 29, 0] == 255                                  # clipped at the 99th percentile, scaled to [0, 1]
