"""The C-ABI library builds, loads and exports every symbol include/dcb200.h declares
(no compute calls here - this runs without a GPU)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, 'include', 'dcb200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(dcb_[A-Za-z0-9_]+)\s*\(', src)))


def test_header_declares_the_hot_path_entry_points():
    names = _declared()
    for must in ('dcb_proj_mean_max_f32', 'dcb_standardize_f32', 'dcb_conv3x3_fwd', 'dcb_convT2x2_fwd',
                 'dcb_conv3x3_wgrad', 'dcb_bn_bwd_reduce', 'dcb_head_loss_bwd', 'dcb_adam_step',
                 'dcb_tta_combine'):
        assert must in names


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as ge
    ge.build()
    from deepcalcium import _native
    lib = ctypes.CDLL(_native.LIB_PATH)
    missing = [n for n in _declared() if not hasattr(lib, n)]
    assert not missing, missing
    lib.dcb_version.restype = ctypes.c_int
    assert lib.dcb_version() >= 100


def test_no_cpu_fallback_without_gpu():
    import pytest
    import torch
    from deepcalcium import _native
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    with pytest.raises(_native.DcbError):
        _native.require_cuda()
