"""K1 / K1b parity on the GPU: projection max bit-exact, mean <= 1e-6 relative, standardise."""
import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu


def _proj(movie, **kw):
    from deepcalcium.datasets.nf import summarize_movie_device
    mean, mx = summarize_movie_device(torch.from_numpy(movie).cuda(), **kw)
    torch.cuda.synchronize()
    return mean.cpu().numpy(), mx.cpu().numpy()


def test_golden_projection(cuda, golden_dir):
    z = np.load(golden_dir + '/projection_small.npz')
    mean, mx = _proj(z['movie'])
    assert np.array_equal(mx, z['max'])
    assert np.max(np.abs(mean - z['mean']) / np.abs(z['mean'])) <= 1e-6


@pytest.mark.parametrize('variant', [0, 1, 2, 3])
@pytest.mark.parametrize('t_splits', [1, 3, 8])
def test_projection_variants_and_splits(cuda, variant, t_splits):
    rng = np.random.default_rng(variant * 10 + t_splits)
    movie = (rng.random((211, 64, 96), dtype=np.float32) * 4096).astype(np.float32)
    mean, mx = _proj(movie, variant=variant, t_splits=t_splits)
    omean, omx = oracle.project_mean_max(movie)
    assert np.array_equal(mx, omx)
    assert np.max(np.abs(mean - omean) / np.abs(omean)) <= 1e-6


@pytest.mark.parametrize('shape', [(1, 8, 8), (5, 3, 7), (300, 16, 20), (17, 512, 512), (64, 31, 33)])
def test_projection_ragged_shapes(cuda, shape):
    rng = np.random.default_rng(sum(shape))
    movie = (rng.standard_normal(shape) * 100).astype(np.float32)        # signed data
    mean, mx = _proj(movie)
    omean, omx = oracle.project_mean_max(movie)
    assert np.array_equal(mx, omx)
    scale = np.abs(movie).mean(0, dtype=np.float64)
    assert np.max(np.abs(mean - omean) / scale) <= 1e-6


def test_projection_floor_and_nan_and_int16_like(cuda):
    rng = np.random.default_rng(1)
    movie = rng.integers(-500, 3000, size=(100, 32, 32)).astype(np.float32)
    movie[:, 0, 0] = -7.0
    mean, mx = _proj(movie, floor_max_at_zero=True)
    omean, omx = oracle.project_mean_max(movie, floor_max_at_zero=True)
    assert np.array_equal(mx, omx) and mx[0, 0] == 0.0
    movie[3, 5, 5] = np.nan
    mean, mx = _proj(movie)
    assert np.isnan(mx[5, 5]) and np.isnan(mean[5, 5])          # numpy semantics: NaN propagates
    assert np.isfinite(np.delete(mx.ravel(), 5 * 32 + 5)).all()


def test_projection_bad_arguments_raise(cuda):
    from deepcalcium.datasets.nf import summarize_movie_device
    with pytest.raises(ValueError):
        summarize_movie_device(torch.zeros(4, 4, device='cuda'))
    with pytest.raises(ValueError):
        summarize_movie_device(torch.zeros(0, 4, 4, device='cuda'))
    with pytest.raises(TypeError):
        summarize_movie_device(torch.zeros(2, 4, 4, device='cuda', dtype=torch.float64))


def test_projection_full_size_properties(cuda):
    """BASELINE config C2 size (3000x512x512 fp32 = 3.1 GB): checked through size-independent
    properties - constant movie, linearity in a per-frame offset, max of a planted spike - plus a
    strided oracle check on a subset of pixels."""
    from deepcalcium.datasets.nf import summarize_movie_device
    T, H, W = 3000, 512, 512
    g = torch.Generator(device='cuda'); g.manual_seed(7535)
    movie = torch.rand((T, H, W), device='cuda', generator=g) * 4096
    movie[1234, 100, 200] = 5000.0
    mean, mx = summarize_movie_device(movie)
    assert mx[100, 200].item() == 5000.0
    sub = movie[:, ::64, ::64].cpu().numpy()
    omean, omx = oracle.project_mean_max(sub)
    assert np.array_equal(mx[::64, ::64].cpu().numpy(), omx)
    assert np.max(np.abs(mean[::64, ::64].cpu().numpy() - omean) / omean) <= 1e-6
    # idempotence / determinism: a second run is bit-identical
    mean2, mx2 = summarize_movie_device(movie)
    assert torch.equal(mean, mean2) and torch.equal(mx, mx2)
    del movie
    const = torch.full((T, 64, 64), 3.25, device='cuda')
    mean, mx = summarize_movie_device(const)
    assert torch.all(mean == 3.25) and torch.all(mx == 3.25)


def test_standardize_matches_summarize_series(cuda, golden_dir):
    from deepcalcium.engine import ops
    z = np.load(golden_dir + '/projection_small.npz')
    x = z['mean'].astype(np.float16).astype(np.float32)
    xd = torch.from_numpy(x).cuda()
    out = torch.empty_like(xd)
    stats = torch.zeros(2, dtype=torch.float64, device='cuda')
    ops.standardize(xd, out, stats)
    assert np.allclose(out.cpu().numpy(), z['summary'], atol=2e-6, rtol=1e-6)
    assert abs(stats[0].item() - x.astype(np.float64).mean()) < 1e-9 * abs(x.mean()) + 1e-9
    img = np.random.default_rng(0).standard_normal((512, 512)).astype(np.float32) * 37 + 500
    out = torch.empty(512, 512, device='cuda')
    ops.standardize(torch.from_numpy(img).cuda(), out)
    assert np.allclose(out.cpu().numpy(), oracle.summarize_series(img), atol=2e-6, rtol=1e-5)


@pytest.mark.parametrize('shape', [(1, 8, 8), (5, 3, 7), (300, 16, 24), (17, 512, 512), (64, 31, 33), (1000, 64, 64)])
def test_projection_int16_movie_is_exact(cuda, shape):
    """int16 frames (the reference's TIFFs, nf.py:121): integer sums are exact, so the mean is the correctly rounded
    float32 of sum/T and the max is bit-exact; signed data, extreme values, ragged shapes (scalar fallback)."""
    rng = np.random.default_rng(sum(shape))
    movie = rng.integers(-32768, 32768, size=shape, dtype=np.int16)
    movie[0, 0, 0] = -32768; movie[-1, -1, -1] = 32767
    mean, mx = _proj(movie)
    exact = movie.astype(np.int64).sum(axis=0) / np.float64(shape[0])
    assert np.array_equal(mx, movie.max(axis=0).astype(np.float32))
    assert np.array_equal(mean, exact.astype(np.float32))
    _, mx0 = _proj(-np.abs(movie.astype(np.int32)).clip(0, 32767).astype(np.int16), floor_max_at_zero=True)
    assert np.all(mx0 >= 0)
    # same answer through the host-buffer entry
    from deepcalcium.datasets.nf import summarize_movie
    mean2, mx2 = summarize_movie(movie)
    assert np.array_equal(mean2, mean) and np.array_equal(mx2, mx)


def test_streaming_tiff_ingest(cuda, tmp_path):
    """datasets/nf.py:104-148 for a dataset directory of 16-bit TIFF frames + regions.json, streamed through
    the device-resident running sum / max in chunks (chunk size not dividing the frame count)."""
    import json
    from PIL import Image
    from deepcalcium.datasets.nf import summarize_tiff_dir, nf_ingest, open_dataset
    rng = np.random.default_rng(11)
    T, H, W = 37, 24, 40
    movie = rng.integers(0, 4096, size=(T, H, W)).astype(np.int16)
    root = tmp_path / 'neurofinder.00.00'
    (root / 'images').mkdir(parents=True); (root / 'regions').mkdir()
    for t in range(T):
        Image.fromarray(movie[t].astype(np.uint16)).save(str(root / 'images' / ('image%05d.tiff' % t)))
    regions = [{'coordinates': [[2, 3], [2, 4], [3, 3]]}, {'coordinates': [[10, 20], [11, 20]]}]
    json.dump(regions, open(str(root / 'regions' / 'regions.json'), 'w'))
    mean, mx, n = summarize_tiff_dir(str(root / 'images'), chunk=8)
    assert n == T
    assert np.array_equal(mx, movie.max(axis=0).astype(np.float32))
    assert np.array_equal(mean, (movie.astype(np.int64).sum(axis=0) / np.float64(T)).astype(np.float32))
    path = nf_ingest('neurofinder.00.00', str(tmp_path), chunk=16)
    ds = open_dataset(path)
    assert np.array_equal(np.asarray(ds['series/max']), movie.max(axis=0))
    assert np.allclose(np.asarray(ds['series/mean'], dtype=np.float32), mean, rtol=1e-3)       # stored as float16
    mm = np.asarray(ds['masks/max'])
    assert mm.sum() == 5 and mm[2, 3] == 1 and mm[11, 20] == 1
    # unsigned 16-bit pixels above 32767 (saturated neurofinder frames) must not wrap: the reference computes the mean and
    # the max from the unwrapped values (nf.py:129-130); series/max saturates when stored as int16
    big = rng.integers(0, 65536, size=(9, H, W)).astype(np.uint16)
    (tmp_path / 'u16').mkdir()
    for t in range(9):
        Image.fromarray(big[t]).save(str(tmp_path / 'u16' / ('image%05d.tiff' % t)))
    mean, mx, n = summarize_tiff_dir(str(tmp_path / 'u16'), chunk=4)
    assert np.array_equal(mx, big.max(axis=0).astype(np.float32))
    assert np.array_equal(mean, (big.astype(np.int64).sum(axis=0) / np.float64(9)).astype(np.float32))
