"""Oracle: movie -> summary image projection and summary standardisation.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  PARITY UNPINNED.

Follows
  * deepcalcium/datasets/nf.py:115-130  (streaming per-frame mean / max loop;
    twin copy at examples/neurons/unet2ds_sj.py:67-85)
  * deepcalcium/models/neurons/unet_2d_summary.py:227-241 (_summarize_series)
"""
import numpy as np


def project_mean_max(movie, floor_max_at_zero=False):
    """Mathematical per-pixel temporal mean and max of a [T,H,W] movie.

    The reference accumulates ``ds_mean[...] += img * 1. / T`` frame by frame
    (nf.py:129) and ``ds_max[...] = np.maximum(ds_max[...], img)`` starting
    from zeros (nf.py:125,130).  The north star asks for the true mean of a
    float32 movie within 1e-6 relative, so the truth here is the float64
    accumulate rounded once to float32; ``floor_max_at_zero=True`` reproduces
    the reference's zero-initialised running max.
    """
    movie = np.asarray(movie)
    mean = movie.mean(axis=0, dtype=np.float64).astype(np.float32)
    mx = movie.max(axis=0).astype(np.float32)
    if floor_max_at_zero:
        mx = np.maximum(mx, np.float32(0))
    return mean, mx


def project_streaming_fp16(movie):
    """Bit-faithful emulation of the reference loop (nf.py:121-130): the mean is
    a float16 HDF5 dataset that is read, added to in float64 and rounded back to
    float16 once per frame; the max is an int16 dataset initialised to zero.
    Only used to document how far the reference's stored mean is from the true
    mean; the GPU path is checked against ``project_mean_max``.
    """
    movie = np.asarray(movie)
    T = movie.shape[0]
    ds_mean = np.zeros(movie.shape[1:], dtype=np.float16)
    ds_max = np.zeros(movie.shape[1:], dtype=np.int16)
    for t in range(T):
        img = movie[t]
        ds_mean[...] = (ds_mean.astype(np.float64) + img * 1. / T).astype(np.float16)
        ds_max[...] = np.maximum(ds_max, img).astype(np.int16)
    return ds_mean, ds_max


def summarize_series(mean_image):
    """unet_2d_summary.py:238-239: float32 cast then (x - mean) / std with the
    population standard deviation (numpy default ddof=0), all in float32."""
    summ = np.asarray(mean_image).astype(np.float32)
    return ((summ - np.mean(summ)) / np.std(summ)).astype(np.float32)
