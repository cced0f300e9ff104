"""Training crop sampler (reference: unet_2d_summary.py:434-530, _batch_gen).  The device sampler splits it into a host
half that keeps the reference's numpy RNG stream (crop descriptors) and a device half that does the pixel work
(dcb_crop_batch).  CPU: the descriptors + a numpy statement of the kernel reproduce the host sampler of oracle/sampler.py
exactly.  GPU: the kernel reproduces it too."""
import numpy as np
import pytest

from oracle.sampler import host_batches


def _case():
    rng = np.random.default_rng(0)
    S = [rng.standard_normal((90, 120)).astype(np.float32), rng.standard_normal((140, 100)).astype(np.float32),
         rng.standard_normal((64, 64)).astype(np.float32)]
    M = [(rng.random(s.shape) < 0.05).astype(np.uint8) for s in S]
    yc = [(0, 70), (20, 140), (0, 64)]
    return S, M, ['a', 'b', 'c'], yc


def _api():
    from deepcalcium.models.neurons.unet_2d_summary import UNet2DSummary
    return UNet2DSummary.__new__(UNet2DSummary)


@pytest.mark.parametrize('window,max_aug', [(64, 15), (32, 0), (96, 3)])
def test_descriptors_reproduce_the_host_sampler(window, max_aug):
    api = _api()
    S, M, names, yc = _case()
    np.random.seed(865)
    ref = host_batches(S, M, names, yc, 16, 10, (window, window), max_aug)
    ref = [next(ref) for _ in range(4)]
    np.random.seed(865)
    gen = api._crop_descriptors(S, M, names, yc, 16, 10, (window, window), max_aug)
    for k in range(4):
        desc = next(gen)
        assert desc.dtype == np.int32 and desc.shape == (16, 12)
        xs, ys = api._apply_descriptors_host(S, M, desc, window)
        assert np.array_equal(xs, ref[k][0]) and np.array_equal(ys, ref[k][1])


def test_choice_over_the_function_table_draws_like_choice_over_its_length():
    """the reference picks augmentations with rng.choice(augment_funcs, n) (:525); the restatements draw
    rng.choice(6, n) - the same stream"""
    funcs = [lambda a: a, lambda a: a + 1, lambda a: a + 2, lambda a: a + 3, lambda a: a + 4, lambda a: a + 5]
    for seed in (1, 865, 7535):
        np.random.seed(seed)
        picked = [funcs.index(f) for f in np.random.choice(funcs, 11)]
        after = np.random.randint(0, 1 << 30)
        np.random.seed(seed)
        assert picked == [int(k) for k in np.random.choice(len(funcs), 11)]
        assert after == np.random.randint(0, 1 << 30)


def test_augmentation_maps_are_the_numpy_flips_and_rotations():
    api = _api()
    n = 7
    a = np.arange(n * n, dtype=np.float32).reshape(n, n)
    funcs = [lambda v: v, lambda v: v[:, ::-1], lambda v: v[::-1, :], lambda v: np.rot90(v, 1), lambda v: np.rot90(v, 2),
             lambda v: np.rot90(v, 3)]
    ii, jj = np.meshgrid(np.arange(n), np.arange(n), indexing='ij')
    for f, ((m00, m01, m10, m11), (t0, t1)) in zip(funcs, api._aug_maps(n)):
        assert np.array_equal(f(a), a[m00 * ii + m01 * jj + t0, m10 * ii + m11 * jj + t1])


@pytest.mark.gpu
def test_device_sampler_equals_the_host_sampler(cuda):
    import torch
    api = _api()
    S, M, names, yc = _case()
    for window, max_aug in ((64, 15), (128, 6)):
        np.random.seed(7535)
        ref = host_batches(S, M, names, yc, 32, 10, (window, window), max_aug)
        ref = [next(ref) for _ in range(3)]
        np.random.seed(7535)
        gen = api._batch_gen(S, M, names, yc, 32, 10, (window, window), max_aug)
        for k in range(3):
            xb, yb = next(gen)
            assert xb.is_cuda and yb.dtype == torch.uint8
            assert np.array_equal(xb.cpu().numpy(), ref[k][0]) and np.array_equal(yb.cpu().numpy(), ref[k][1])
