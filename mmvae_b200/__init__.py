"""mmvae_b200 -- B200-native (sm_100a) implementation of the VAE training step of
praateekmahajan/moving-mnist-vae, behind the reference's own `VAE` nn.Module API.

    from mmvae_b200 import VAE           # drop-in for reference model.VAE (pixelcnn=False)
    from mmvae_b200 import NotebookVAE   # the encoder / decoder pair and loop body of vae-kl.ipynb
    from mmvae_b200 import FusedAdam     # optim.Adam(model.parameters()) as one kernel over the flat arena

The numerical work lives in libmmvae_b200.so (C ABI: include/mmvae.h); importing this package
fails if that library has not been built (`python -m mmvae_b200.build`).
"""
from . import _lib
from ._lib import MMVAEError
from .graph import GraphedTrainStep
from .model import VAE
from .notebook import NotebookVAE
from .optim import FusedAdam

__all__ = ["VAE", "NotebookVAE", "GraphedTrainStep", "FusedAdam", "MMVAEError", "_lib"]
