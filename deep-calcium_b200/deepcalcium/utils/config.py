"""Same on-disk config as the reference (deepcalcium/utils/config.py:6-38):
~/.deep-calcium/deep-calcium.json with datasets_dir / checkpoints_dir, created on first use.
Unlike the reference the directories are created lazily (get_config()), not at import time of
every module, but importing this module keeps the reference's exported names."""
from platform import system
import json
import os


def get_config():
    base_dir_windows = '%s/Documents/deep-calcium' % os.path.expanduser('~')
    base_dir_unix = '%s/.deep-calcium' % os.path.expanduser('~')
    base_dir = base_dir_windows if system() == 'Windows' else base_dir_unix
    base_dir = os.environ.get('DEEP_CALCIUM_HOME', base_dir)
    config_path = '%s/deep-calcium.json' % base_dir
    os.makedirs(base_dir, exist_ok=True)
    if not os.path.exists(config_path):
        config = {'datasets_dir': '%s/datasets' % base_dir, 'checkpoints_dir': '%s/checkpoints' % base_dir}
        with open(config_path, 'w') as fp:
            json.dump(config, fp)
    else:
        with open(config_path, 'r') as fp:
            config = json.load(fp)
    os.makedirs(config['datasets_dir'], exist_ok=True)
    os.makedirs(config['checkpoints_dir'], exist_ok=True)
    return config


config = get_config()
DATASETS_DIR = config['datasets_dir']
CHECKPOINTS_DIR = config['checkpoints_dir']
