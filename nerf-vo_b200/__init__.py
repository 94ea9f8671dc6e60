"""nvo_b200 — B200-native (sm_100a) kernels for NeRF-VO's NeRF mapping hot path, behind the reference's own operator
API (tinycudann modules; nerfstudio fields / samplers / renderers / losses).  See DESIGN.md and include/nvo_b200.h."""
from . import _lib, ops, sharding  # noqa: F401
from . import tcnn_api  # noqa: F401
from . import checkpoint  # noqa: F401
from . import exporter  # noqa: F401
from .data import (CameraOptimizer, CameraOptimizerConfig, Cameras, DynamicDataManager, DynamicDataManagerConfig, DynamicDataset, PixelSampler,  # noqa: F401
                   PixelSamplerConfig, RayGenerator)
from .field_components import MLP, Embedding, HashEncoding, MLPWithHashEncoding, NeRFEncoding, SceneContraction, SHEncoding, trunc_exp  # noqa: F401
from .fields import FieldHeadNames, HashMLPDensityField, NerfactoField  # noqa: F401
from .model import DepthNerfactoModel, ExtendedNerfactoModel, NerfactoModel, NerfactoModelConfig  # noqa: F401
from .nerf_renderer import NerfstudioRenderer  # noqa: F401
from .ray_samplers import PDFSampler, ProposalNetworkSampler, UniformLinDispPiecewiseSampler  # noqa: F401
from .rays import Frustums, RayBundle, RaySamples  # noqa: F401
from .renderers import AccumulationRenderer, DepthRenderer, NormalsRenderer, NormalsShader, RGBRenderer, render_all  # noqa: F401

__version__ = "0.1.0"
