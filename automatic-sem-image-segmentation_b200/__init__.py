"""B200-native conv-stack hot path of BAMresearch/automatic-sem-image-segmentation (Release 1.2.0).

Python host code holding torch tensors and calling hand-written sm_100a CUDA kernels through the C ABI of
libsemb200.so (include/semb200.h).  Importable as `sem_b200` (see sem_b200.py at the repo root; the
directory name contains hyphens).  There is no CPU fallback: constructing a model without an sm_100
device raises.
"""
from . import build, _lib, engine, nets, model, gan_nets, cyclegan_model, wgan_nets, wgan_model, dp, keras_compat, keras_io  # noqa: F401
from . import Measurements, Scores  # noqa: F401
from . import HelperFunctions, UNet_Segmentation, CycleGAN, WassersteinGAN, StartProcess  # noqa: F401  (drop-in module names)
from .model import UNetModel, load_model  # noqa: F401
from .cyclegan_model import CycleGanModel, ImagePool  # noqa: F401
from .wgan_model import WganGpModel  # noqa: F401

__all__ = ["UNetModel", "load_model", "build", "engine", "nets", "model"]
