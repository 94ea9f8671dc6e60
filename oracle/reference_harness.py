"""TEST INFRASTRUCTURE ONLY — imports the UNMODIFIED reference (nerfstudio fork under
/root/reference) on CPU so golden vectors can be generated and the oracle restatement
(`oracle/nerfacto_oracle.py`) can be validated against the real thing.

Only usable in the build container: /root/reference does not exist on the GPU box, so nothing
under tests/ -m gpu, smoke() or bench.py imports this file.

Eight third-party packages that the reference imports at module scope but never executes on
the nerfacto mapping path are not installed here; they are replaced by inert
stub modules through a sys.meta_path finder.
"""
from __future__ import annotations

import importlib.abc
import importlib.machinery
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("NVO_REFERENCE_ROOT", "/root/reference")
NS_ROOT = os.path.join(REFERENCE_ROOT, "nerf_vo", "thirdparty", "nerfstudio")

_STUBBED = ("viser", "nerfacc", "comet_ml", "torchmetrics", "lietorch", "wandb", "tensorboard", "matplotlib")


class _Anything:
    """Attribute sink: any attribute / call / subscript returns another sink, usable as a base class."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _Anything()

    def __getitem__(self, k):
        return _Anything()

    def __iter__(self):
        return iter(())

    def __mro_entries__(self, bases):
        return (object,)


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _Anything()


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        top = fullname.split(".")[0]
        if top in _STUBBED:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _StubModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


_installed = False


def available() -> bool:
    return os.path.isdir(NS_ROOT)


def install() -> None:
    """Make `import nerfstudio...` resolve to the reference tree (idempotent)."""
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError(f"reference tree not found at {NS_ROOT}; the harness only runs in the build container")
    sys.meta_path.append(_StubFinder())  # appended: real packages win when installed
    sys.path.insert(0, NS_ROOT)
    _installed = True



# ----------------------------------------------------------------------------------------
# helpers to run the reference model on CPU with recorded randomness
# ----------------------------------------------------------------------------------------


class record_rand:
    """Context manager: records every tensor torch.rand returns (the samplers' jitters) in call order."""

    def __enter__(self):
        import torch

        self._torch = torch
        self._orig = torch.rand
        self.values = []

        def rand(*a, **k):
            v = self._orig(*a, **k)
            self.values.append(v.clone())
            return v

        torch.rand = rand
        return self

    def __exit__(self, *exc):
        self._torch.rand = self._orig
        return False


class record_searchsorted:
    """Context manager: records (args, result) of every torch.searchsorted call."""

    def __enter__(self):
        import torch

        self._torch = torch
        self._orig = torch.searchsorted
        self.calls = []

        def ss(*a, **k):
            r = self._orig(*a, **k)
            self.calls.append((a, k, r.clone()))
            return r

        torch.searchsorted = ss
        return self

    def __exit__(self, *exc):
        self._torch.searchsorted = self._orig
        return False


def build_reference_model(main_log2=19, prop_log2=17, num_images=192, predict_normals=True, seed=0):
    """The unmodified DepthNerfactoModel(implementation='torch') with NeRF-VO's loss multipliers
    (nerf_vo/mapping/nerfstudio.py:71-82); camera optimizer off."""
    install()
    import torch
    from nerfstudio.cameras.camera_optimizers import CameraOptimizerConfig
    from nerfstudio.data.scene_box import SceneBox
    from nerfstudio.models.depth_nerfacto import DepthNerfactoModel, DepthNerfactoModelConfig

    torch.manual_seed(seed)
    cfg = DepthNerfactoModelConfig(
        implementation="torch",
        predict_normals=predict_normals,
        camera_optimizer=CameraOptimizerConfig(mode="off"),
        log2_hashmap_size=main_log2,
        proposal_net_args_list=[
            {"hidden_dim": 16, "log2_hashmap_size": prop_log2, "num_levels": 5, "max_res": 128, "use_linear": False},
            {"hidden_dim": 16, "log2_hashmap_size": prop_log2, "num_levels": 5, "max_res": 256, "use_linear": False},
        ],
        interlevel_loss_mult=1.0,
        distortion_loss_mult=0.002,
        orientation_loss_mult=0,
        pred_normal_loss_mult=0,
        depth_loss_mult=0.001,
        is_euclidean_depth=False,
        depth_sigma=0.001,
        should_decay_sigma=False,
    )
    sb = SceneBox(aabb=torch.tensor([[-1.0, -1, -1], [1, 1, 1]]))
    model = DepthNerfactoModel(cfg, scene_box=sb, num_train_data=num_images)
    return model


def reference_state(model):
    """state_dict restricted to the trainable tensors of the hot path (oracle parameter keys)."""
    return {k: v.detach().clone() for k, v in model.state_dict().items() if k.startswith(("field.", "proposal_networks.")) and v.dtype.is_floating_point and v.ndim > 0
            and not k.endswith(".aabb") and ".mlp_base.0.hash_table" not in k}


def reference_step(model, rays, targets, normal_loss_mult=0.000005, step=0):
    """One training forward + losses (+ backward) through the reference's own public methods.
    Returns (outputs, loss_dict, jitters, searchsorted_calls)."""
    import torch
    from nerfstudio.cameras.rays import RayBundle
    from nerfstudio.model_components.losses import monosdf_normal_loss

    model.train()
    for cb_step in (step,):
        # anneal callback (NS/models/nerfacto.py:260-270)
        import numpy as np

        frac = float(np.clip(cb_step / model.config.proposal_weights_anneal_max_num_iters, 0, 1))
        b = model.config.proposal_weights_anneal_slope
        model.proposal_sampler.set_anneal(b * frac / ((b - 1) * frac + 1))
    rb = RayBundle(
        origins=rays["origins"].clone(),
        directions=rays["directions"].clone(),
        pixel_area=rays["pixel_area"].clone(),
        camera_indices=rays["camera_indices"].clone(),
        metadata={"directions_norm": rays["directions_norm"].clone()},
    )
    model.zero_grad()
    with record_rand() as rr, record_searchsorted() as rs:
        out = model(rb)
        batch = {"image": targets["rgb"], "depth_image": targets["depth"]}
        md = model.get_metrics_dict(out, batch)
        ld = model.get_loss_dict(out, batch, md)
        if "normal" in targets and normal_loss_mult > 0:
            # nerf_vo/mapping/nerfstudio_utils.py:337-349 (ExtendedNerfactoModel)
            ld["normal_loss"] = normal_loss_mult * monosdf_normal_loss(normal_pred=out["normals"], normal_gt=targets["normal"])
    total = sum(ld.values())
    total.backward()
    jit = [v for v in rr.values if v.ndim == 2 and v.shape[1] == 1][:3]
    return out, ld, jit, rs.calls
