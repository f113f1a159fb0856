"""Ray containers, cameras and the camera optimiser (mirror of the reference's `cameras` package for
the hot path).  ref: src/cameras/rays.py:35-349, src/cameras/cameras.py:60-130,
src/cameras/camera_optimizers.py:34-133, src/model_components/ray_generators.py:34-81
"""
from dataclasses import dataclass, field
from typing import Callable, Dict, Optional, Type

import torch
from torch import nn

from . import ops
from .configs import InstantiateConfig


@dataclass
class Frustums:
    """ref: rays.py:35-81"""
    origins: torch.Tensor       # [R, 1 or S, 3]
    directions: torch.Tensor    # [R, 1 or S, 3]
    starts: torch.Tensor        # [R, S, 1]
    ends: torch.Tensor          # [R, S, 1]
    pixel_area: Optional[torch.Tensor] = None
    up_directions: Optional[torch.Tensor] = None

    def get_positions(self):
        return self.origins + self.directions * (self.starts + self.ends) / 2

    def get_start_positions(self):
        return self.origins + self.directions * self.starts


@dataclass
class RaySamples:
    """ref: rays.py:117-237"""
    frustums: Frustums
    camera_indices: Optional[torch.Tensor] = None
    deltas: Optional[torch.Tensor] = None
    spacing_starts: Optional[torch.Tensor] = None
    spacing_ends: Optional[torch.Tensor] = None
    spacing_to_euclidean_fn: Optional[Callable] = None

    @property
    def shape(self):
        return self.deltas.shape[:-1]

    def slice_rays(self, a: int, b: int) -> "RaySamples":
        """Rays [a, b) of the batch (views)."""
        def sl(t):
            return t[a:b] if t is not None else None
        f = self.frustums
        return RaySamples(
            frustums=Frustums(origins=sl(f.origins), directions=sl(f.directions), starts=sl(f.starts), ends=sl(f.ends),
                              pixel_area=sl(f.pixel_area), up_directions=sl(f.up_directions)),
            camera_indices=sl(self.camera_indices), deltas=sl(self.deltas), spacing_starts=sl(self.spacing_starts),
            spacing_ends=sl(self.spacing_ends), spacing_to_euclidean_fn=None)

    def get_weights_from_densities(self, densities):
        """alphas (rays.py:138-151) followed by get_weights_from_alphas (rays.py:201-217), fused."""
        return ops.DensityWeightsFn.apply(densities[..., 0], self.deltas[..., 0])[..., None]


@dataclass
class RayBundle:
    """ref: rays.py:240-349"""
    camera_indices: Optional[torch.Tensor]
    origins: torch.Tensor
    directions: torch.Tensor
    up_directions: Optional[torch.Tensor] = None
    pixel_area: Optional[torch.Tensor] = None
    directions_norm: Optional[torch.Tensor] = None
    nears: Optional[torch.Tensor] = None
    fars: Optional[torch.Tensor] = None

    @property
    def shape(self):
        return self.origins.shape[:-1]

    def __len__(self):
        return self.origins.shape[0]

    def get_ray_samples(self, bin_starts, bin_ends, spacing_starts=None, spacing_ends=None,
                        spacing_to_euclidean_fn=None) -> RaySamples:
        frustums = Frustums(
            origins=self.origins[..., None, :], directions=self.directions[..., None, :],
            up_directions=self.up_directions[..., None, :] if self.up_directions is not None else None,
            starts=bin_starts, ends=bin_ends,
            pixel_area=self.pixel_area[..., None, :] if self.pixel_area is not None else None)
        return RaySamples(
            frustums=frustums,
            camera_indices=self.camera_indices[..., None] if self.camera_indices is not None else None,
            deltas=bin_ends - bin_starts, spacing_starts=spacing_starts, spacing_ends=spacing_ends,
            spacing_to_euclidean_fn=spacing_to_euclidean_fn)


class Cameras:
    """Perspective cameras of one modality (the subset of cameras.py:60-130 the hot path reads).
    camera_to_worlds [n,3,4]; fx, fy, cx, cy [n] or scalars; distortion_params [n,6] or None."""

    def __init__(self, camera_to_worlds, fx, fy, cx, cy, width=None, height=None, distortion_params=None):
        n = camera_to_worlds.shape[0]
        self.camera_to_worlds = camera_to_worlds.float()

        def _b(v):
            v = torch.as_tensor(v, dtype=torch.float32).reshape(-1)
            return v.expand(n) if v.numel() == 1 else v

        self.intrinsics = torch.stack([_b(fx), _b(fy), _b(cx), _b(cy)], dim=-1).contiguous()
        self.distortion_params = distortion_params.float() if distortion_params is not None else None
        self.width, self.height = width, height

    def __len__(self):
        return self.camera_to_worlds.shape[0]

    def to(self, device):
        self.camera_to_worlds = self.camera_to_worlds.to(device)
        self.intrinsics = self.intrinsics.to(device)
        if self.distortion_params is not None:
            self.distortion_params = self.distortion_params.to(device)
        return self


@dataclass
class CameraOptimizerConfig(InstantiateConfig):
    """ref: camera_optimizers.py:34-46"""
    _target: Type = field(default_factory=lambda: CameraOptimizer)
    mode: str = "off"
    modalities_to_optimize: Dict[str, bool] = field(default_factory=dict)
    shared_optimization: bool = False


class CameraOptimizer(nn.Module):
    """Holds `pose_adjustment[mod]` ([1,6] shared or [n_cam,6]); the SO3xR3 exponential map and the pose
    composition are evaluated inside the ray-generation kernel.  ref: camera_optimizers.py:48-119"""

    def __init__(self, config: CameraOptimizerConfig, num_cameras: int, **kwargs) -> None:
        super().__init__()
        self.config = config
        self.num_cameras = num_cameras
        if self.config.mode not in ("off", "SO3xR3"):
            raise ValueError(f"Camera optimization mode {self.config.mode} not supported.")
        self.pose_adjustment = nn.ParameterDict()
        if self.config.mode == "SO3xR3":
            for mod in self.config.modalities_to_optimize.keys():
                rows = 1 if self.config.shared_optimization else self.num_cameras
                self.pose_adjustment[mod] = nn.Parameter(torch.zeros((rows, 6)))

    def parameters_for(self, mod):
        if self.config.mode == "off" or mod not in self.pose_adjustment:
            return None
        p = self.pose_adjustment[mod]
        return p if self.config.modalities_to_optimize.get(mod, False) else p.detach()


class RayGenerator(nn.Module):
    """ref: ray_generators.py:34-81 — `forward({mod: int[R,3] (cam, y, x)}) -> {mod: RayBundle}`."""

    def __init__(self, data: Dict, pose_optimizer: CameraOptimizer, pixel_offset: float) -> None:
        super().__init__()
        self.cameras = {mod: mod_data["cameras"] for mod, mod_data in data.items()}
        self.pose_optimizer = pose_optimizer
        self.pixel_offset = pixel_offset

    def forward(self, ray_indices: Dict[str, torch.Tensor]) -> Dict[str, RayBundle]:
        ray_bundles = {}
        for mod, indices in ray_indices.items():
            if indices is None:
                ray_bundles[mod] = None
                continue
            cams = self.cameras[mod].to(indices.device)
            pa = self.pose_optimizer.parameters_for(mod) if self.pose_optimizer is not None else None
            o, d, up, area, dn = ops.RayGenFn.apply(indices, cams.camera_to_worlds, cams.intrinsics,
                                                    cams.distortion_params, pa, float(self.pixel_offset))
            ray_bundles[mod] = RayBundle(camera_indices=indices[:, 0:1].long(), origins=o, directions=d,
                                         up_directions=up, pixel_area=area, directions_norm=dn)
        return ray_bundles
