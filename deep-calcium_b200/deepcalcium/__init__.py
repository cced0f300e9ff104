"""deepcalcium - B200-native drop-in for the UNet2DS hot path of alexklibisz/deep-calcium.

Same import surface as the reference for that path:
  deepcalcium.models.neurons.unet_2d_summary : UNet2DSummary (fit / predict / evaluate), unet
  deepcalcium.datasets.nf                    : summarize_movie (the mean/max projection), nf_load_hdf5
  deepcalcium.utils.neurons                  : losses, metrics, INVERTIBLE_2D_AUGMENTATIONS
All arithmetic runs in hand-written sm_100a CUDA kernels behind the C ABI in include/dcb200.h.
"""
__version__ = '0.1.0'
