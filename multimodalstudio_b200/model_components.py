"""Mirror of the reference's `model_components` package: colliders, samplers, surface / radiance /
background models, volume rendering, renderers and the loss manager — same class and config names and
argument meaning; every named stage of the hot path runs in libmms_b200.so.

Difference in mechanics (not in results): the reference compacts the rays that hit the sphere with a
boolean mask (host sync, models/base_model.py:88-93) and scatters the rendered values back
(renderers.py:105-135).  Here every ray keeps its slot; rays outside the sphere get zero weights inside
the NeuS-weights kernel, which yields the same colours (pure background), zero normals / depth /
accumulation, and the geometry losses average over the unmasked samples only.  No host sync, static
shapes, CUDA-graph capturable.

ref: src/model_components/{scene_colliders,ray_samplers,surface_model,volume_rendering,radiance_model,
     background_model,renderers,losses}.py
"""
from collections import defaultdict
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Type, Union

import numpy as np
import torch
from torch import nn

from . import ops
from .cameras import RayBundle, RaySamples
from .configs import (InstantiateConfig, TrainingCallback, TrainingCallbackAttributes,
                      TrainingCallbackLocation)
from .field_components import (EncodingConfig, FieldComponentConfig, ModalityHeadConfig, NeRFEncodingConfig,
                               SingleVarianceNetwork, SpatialDistortionConfig)
from .fields import NeRFFieldConfig, RadianceFieldConfig, BaseRadianceFieldConfig, SDFFieldConfig, SurfaceFieldConfig


# ---------------------------------------------------------------------------------------------
# colliders  (ref: scene_colliders.py:46-113)
# ---------------------------------------------------------------------------------------------
class SphereCollider(nn.Module):
    def __init__(self, radius: float = 1.0, **kwargs) -> None:
        super().__init__()
        self.radius = radius

    def forward(self, ray_bundle: RayBundle):
        nears, fars, mask = ops.sphere_collide(ray_bundle.origins, ray_bundle.directions, self.radius)
        ray_bundle.nears, ray_bundle.fars = nears, fars
        return ray_bundle, mask.bool()


def _collide_differentiable(o, d, radius):
    """scene_colliders.py:62-79,112-113 as autograd ops (only used when the poses are being optimised)."""
    b = (d * o).sum(dim=-1, keepdim=True)
    under = b ** 2 - (o.norm(p=2, dim=-1, keepdim=True) ** 2 - radius ** 2)
    hit = under > 0.01
    sq = torch.sqrt(under.clamp_min(0.01))
    nears, fars = (-sq - b).clamp_min(0.01), (sq - b).clamp_min(0.01)
    return nears, fars, torch.where(hit, fars, nears), fars + 3.0


class ColliderInstancer:
    def __init__(self, scene_box):
        if scene_box.collider_type == "sphere":
            self.collider = SphereCollider(scene_box.radius)
        else:
            raise ValueError(f"No collider of type {scene_box.collider_type}. Exiting.")

    def update_ray_bundles(self, ray_bundles: Dict[str, RayBundle]):
        masks = {}
        for mod, rb in ray_bundles.items():
            if rb is None:
                masks[mod] = None
                continue
            nears, fars, mask, bgn, bgf = ops.sphere_collide(rb.origins, rb.directions, self.collider.radius, True)
            if rb.origins.requires_grad or rb.directions.requires_grad:
                # pose refinement: near / far must stay differentiable functions of the ray (the kernel
                # above still provides the mask); same formula, evaluated by autograd on [R] tensors.
                nears, fars, bgn, bgf = _collide_differentiable(rb.origins, rb.directions, self.collider.radius)
            rb.nears, rb.fars = nears, fars
            rb._bg_nears, rb._bg_fars = bgn, bgf
            masks[mod] = mask       # uint8 [R]
        return masks

    def update_ray_bundles_for_background(self, ray_bundles: Dict[str, RayBundle]):
        for rb in ray_bundles.values():
            if rb is not None:
                if not hasattr(rb, "_bg_nears"):
                    _, _, _, rb._bg_nears, rb._bg_fars = ops.sphere_collide(rb.origins, rb.directions,
                                                                            self.collider.radius, True)
                rb.nears, rb.fars = rb._bg_nears, rb._bg_fars


@dataclass
class SceneBox:
    """ref: data/scene_box.py:27-44 (the fields the hot path reads)"""
    aabb: Any = None
    near: float = 0.1
    far: float = 6.0
    radius: float = 1.0
    collider_type: str = "sphere"


# ---------------------------------------------------------------------------------------------
# samplers  (ref: ray_samplers.py)
# ---------------------------------------------------------------------------------------------
@dataclass
class SamplerConfig(InstantiateConfig):
    _target: Type = field(default_factory=lambda: Sampler)
    num_samples: int = 32
    train_stratified: bool = True
    single_jitter: bool = False


@dataclass
class UniformSamplerConfig(SamplerConfig):
    _target: Type = field(default_factory=lambda: UniformSampler)


@dataclass
class LinearDisparitySamplerConfig(SamplerConfig):
    _target: Type = field(default_factory=lambda: LinearDisparitySampler)


@dataclass
class PDFSamplerConfig(SamplerConfig):
    _target: Type = field(default_factory=lambda: PDFSampler)
    num_samples: int = 4
    include_original: bool = True
    histogram_padding: float = 0.01


@dataclass
class NeuSSamplerConfig(SamplerConfig):
    _target: Type = field(default_factory=lambda: NeuSSampler)
    num_samples_importance: int = 64
    num_upsample_steps: int = 4
    base_variance: float = 64
    single_jitter: bool = True


class Sampler(nn.Module):
    def __init__(self, config: SamplerConfig, train_stratified=None, single_jitter=None) -> None:
        super().__init__()
        self.config = config
        self.train_stratified = train_stratified if train_stratified is not None else self.config.train_stratified
        self.single_jitter = single_jitter if single_jitter is not None else self.config.single_jitter

    def forward(self, *args, **kwargs):
        return self.generate_ray_samples(*args, **kwargs)

    def get_param_groups(self):
        params = list(self.parameters())
        return {"ray_sampler": params} if len(params) != 0 else {}

    def get_training_callbacks(self, training_callback_attributes):
        return []


def _samples_from_bins(ray_bundle: RayBundle, sbins, ebins, nears, fars, spacing_fn) -> RaySamples:
    """RaySamples from spacing / euclidean bin edges.  With pose refinement the euclidean bins must stay
    a differentiable function of (near, far) — recomputed with torch from the detached spacing bins, the
    kernel's values are used otherwise."""
    if nears.requires_grad or fars.requires_grad:
        ebins = spacing_fn(sbins, nears, fars)
    return ray_bundle.get_ray_samples(
        bin_starts=ebins[..., :-1, None], bin_ends=ebins[..., 1:, None],
        spacing_starts=sbins[..., :-1, None], spacing_ends=sbins[..., 1:, None], spacing_to_euclidean_fn=spacing_fn)


class SpacedSampler(Sampler):
    """ref: ray_samplers.py:156-233.  `rand`: optional {mod: tensor} with the jitter the reference would
    draw with torch.rand ([R,1] single-jitter, [R,N+1] otherwise); drawn here when missing."""
    spacing = ops.SPACING_UNIFORM

    def spacing_to_euclidean(self, x, nears, fars):
        if self.spacing == ops.SPACING_DISPARITY:
            return 1 / ((1 / fars) * x + (1 / nears) * (1 - x))
        return fars * x + nears * (1 - x)

    def generate_ray_samples(self, ray_bundles: Dict[str, RayBundle] = None, num_samples: Optional[int] = None,
                             rand: Optional[Dict[str, torch.Tensor]] = None) -> Dict[str, RaySamples]:
        out = {}
        for mod, rb in ray_bundles.items():
            if rb is None:
                out[mod] = None
                continue
            assert rb.nears is not None and rb.fars is not None
            num_samples = num_samples or self.config.num_samples
            t_rand = None
            if self.train_stratified and self.training:
                if rand is not None and rand.get(mod) is not None:
                    t_rand = rand[mod]
                else:
                    cols = 1 if self.single_jitter else num_samples + 1
                    t_rand = torch.rand((len(rb), cols), dtype=torch.float32, device=rb.origins.device)
            sbins, ebins = ops.spaced_bins(rb.nears, rb.fars, num_samples, self.spacing, t_rand)
            out[mod] = _samples_from_bins(rb, sbins, ebins, rb.nears, rb.fars, self.spacing_to_euclidean)
        return out


class UniformSampler(SpacedSampler):
    spacing = ops.SPACING_UNIFORM


class LinearDisparitySampler(SpacedSampler):
    spacing = ops.SPACING_DISPARITY


class PDFSampler(Sampler):
    """ref: ray_samplers.py:298-422.  Only the stratified `u` construction lives here; the pdf / cdf /
    searchsorted / lerp arithmetic is inside the fused up-sampling kernel."""

    def make_u(self, num_rays, num_samples, device, rand=None):
        num_bins = num_samples + 1
        u = ops.const_tensor(("pdf_u", num_bins), lambda: torch.linspace(0.0, 1.0 - (1.0 / num_bins), steps=num_bins), device)
        if self.config.train_stratified and self.training:
            u = u.expand(num_rays, num_bins)
            if rand is None:
                cols = 1 if self.config.single_jitter else num_bins
                rand = torch.rand((num_rays, cols), device=device)
            u = u + rand / num_bins
        else:
            u = (u + 1.0 / (2 * num_bins)).expand(num_rays, num_bins)
        return u.contiguous()


class NeuSSampler(Sampler):
    """ref: ray_samplers.py:424-551"""

    def __init__(self, config: NeuSSamplerConfig, train_stratified=None, single_jitter=None) -> None:
        super().__init__(config, train_stratified, single_jitter)
        self.config = config
        self.uniform_sampler = UniformSamplerConfig().setup(train_stratified=self.config.train_stratified,
                                                            single_jitter=self.config.single_jitter)
        self.pdf_sampler = PDFSamplerConfig(include_original=False, single_jitter=self.config.single_jitter,
                                            histogram_padding=1e-5).setup()

    def generate_ray_samples(self, ray_bundles: Dict[str, RayBundle] = None, **kwargs):
        sdf_fn = kwargs.get("sdf_fn", None)
        rand = kwargs.get("rand", None) or {}
        assert ray_bundles is not None and sdf_fn is not None
        uniform = kwargs.get("uniform_ray_samples_per_modality", None)
        if uniform is None:
            uniform = self.uniform_sampler(ray_bundles, num_samples=self.config.num_samples, rand=rand.get("uniform"))
        k = self.config.num_samples_importance // self.config.num_upsample_steps
        out = {}
        for mod, rb in ray_bundles.items():
            if rb is None:
                out[mod] = None
                continue
            samples = uniform[mod]
            sbins = torch.cat([samples.spacing_starts[..., 0], samples.spacing_ends[..., -1:, 0]], dim=-1)
            nears, fars = rb.nears.detach(), rb.fars.detach()
            if (rand.get("bins") or {}).get(mod) is not None:
                # test hook: externally supplied final spacing bins [R, S+1] (isolates the sampler from the rest)
                sbins = rand["bins"][mod]
                eb = self.uniform_sampler.spacing_to_euclidean(sbins, nears, fars)
                out[mod] = _samples_from_bins(rb, sbins, eb, rb.nears, rb.fars, self.uniform_sampler.spacing_to_euclidean)
                continue
            new_samples, sdf = samples, None
            pdf_rand = (rand.get("pdf") or {}).get(mod)
            for it in range(self.config.num_upsample_steps):
                with torch.no_grad():
                    new_sdf = sdf_fn(new_samples)[..., 0]
                    sdf = new_sdf if sdf is None else ops.merge_rows(sdf, new_sdf, index)
                    u = self.pdf_sampler.make_u(len(rb), k, sbins.device, None if pdf_rand is None else pdf_rand[it])
                    new_bins, sbins, index = ops.neus_upsample(
                        sbins, sdf, u, nears, fars, inv_s=self.config.base_variance * 2 ** it,
                        histogram_padding=self.pdf_sampler.config.histogram_padding)
                    nb_e = self.uniform_sampler.spacing_to_euclidean(new_bins, nears, fars)
                    new_samples = rb.get_ray_samples(bin_starts=nb_e[..., :-1, None], bin_ends=nb_e[..., 1:, None])
            eb = self.uniform_sampler.spacing_to_euclidean(sbins, nears, fars)
            out[mod] = _samples_from_bins(rb, sbins, eb, rb.nears, rb.fars, self.uniform_sampler.spacing_to_euclidean)
        return {"ray_samples_per_modality": out}


# ---------------------------------------------------------------------------------------------
# volume rendering  (ref: volume_rendering.py)
# ---------------------------------------------------------------------------------------------
@dataclass
class DensityConfig(InstantiateConfig):
    _target: Type = field(default_factory=lambda: NeuSDensity)
    init_val: float = 0.3


@dataclass
class NeuSDensityConfig(DensityConfig):
    _target: Type = field(default_factory=lambda: NeuSDensity)


class NeuSDensity(nn.Module):
    def __init__(self, config: NeuSDensityConfig):
        super().__init__()
        self.config = config
        self.variance_network = SingleVarianceNetwork(init_val=self.config.init_val)

    def get_param_groups(self):
        return {"density_fn": list(self.parameters())}


@dataclass
class VolumeRenderingConfig(InstantiateConfig):
    _target: Type = field(default_factory=lambda: NeuSVolumeRendering)
    density_fn: DensityConfig = field(default_factory=lambda: NeuSDensityConfig)


@dataclass
class NeuSVolumeRenderingConfig(VolumeRenderingConfig):
    _target: Type = field(default_factory=lambda: NeuSVolumeRendering)
    anneal_end_ratio: float = 0.05


class NeuSVolumeRendering(nn.Module):
    """ref: volume_rendering.py:161-239 — alphas + cumprod transmittance are one warp-scan kernel."""

    def __init__(self, config: NeuSVolumeRenderingConfig):
        super().__init__()
        self.config = config
        self.density_fn = self.config.density_fn.setup()
        self._cos_anneal_ratio = 1.0

    def forward(self, ray_samples: RaySamples, sdf, gradients, mask=None):
        s = self.density_fn.variance_network.get_inv_variance()
        dirs = ray_samples.frustums.directions[..., 0, :]
        w = ops.NeusWeightsFn.apply(sdf[..., 0], gradients, dirs, ray_samples.deltas[..., 0], s, mask,
                                    float(self._cos_anneal_ratio))
        return w[..., None]

    def set_cos_anneal_ratio(self, anneal: float) -> None:
        self._cos_anneal_ratio = anneal

    def get_param_groups(self):
        return self.density_fn.get_param_groups()

    def get_training_callbacks(self, training_callback_attributes):
        callbacks = []
        if self.config.anneal_end_ratio > 0:
            def set_anneal(step):
                anneal_end = int(training_callback_attributes.trainer.max_num_iterations * self.config.anneal_end_ratio)
                self.set_cos_anneal_ratio(min([1.0, step / anneal_end]))

            callbacks.append(TrainingCallback(where_to_run=[TrainingCallbackLocation.BEFORE_TRAIN_ITERATION],
                                              update_every_num_iters=1, func=set_anneal))
        return callbacks


# ---------------------------------------------------------------------------------------------
# surface model  (ref: surface_model.py)
# ---------------------------------------------------------------------------------------------
@dataclass
class SurfaceModelConfig(InstantiateConfig):
    _target: Type = field(default_factory=lambda: SurfaceModel)
    surface_field: SurfaceFieldConfig = field(default_factory=lambda: SDFFieldConfig)
    volume_rendering: VolumeRenderingConfig = field(default_factory=lambda: NeuSVolumeRenderingConfig)
    spatial_distortion: Union[None, SpatialDistortionConfig] = None
    use_numerical_gradients: bool = False
    numerical_gradient_taps: int = 4
    compute_hessian: bool = False


class SurfaceModel(nn.Module):
    def __init__(self, config: SurfaceModelConfig):
        super().__init__()
        self.config = config
        self.surface_field = self.config.surface_field.setup()
        self.volume_rendering = self.config.volume_rendering.setup()
        self.spatial_distortion = self.config.spatial_distortion.setup() \
            if self.config.spatial_distortion is not None else None
        self.numerical_gradients_delta = None
        if not self.config.use_numerical_gradients and self.config.compute_hessian:
            raise NotImplementedError("Hessians of the autograd SDF gradient (a third-order backward) are not on the B200 path; "
                                      "the shipped `mlp*` presets set compute_hessian=False")
        if self.config.use_numerical_gradients and self.config.numerical_gradient_taps != 4:
            raise ValueError("Invalid number of taps for numerical gradients. Must be 4.")

    def forward(self, ray_samples: RaySamples, return_weights=True, mask=None):
        shape = ray_samples.shape
        inputs = ray_samples.frustums.get_start_positions().reshape(-1, 3)
        if self.spatial_distortion is not None:
            inputs = self.spatial_distortion(inputs)
        n = inputs.shape[0]
        if not self.config.use_numerical_gradients:
            return self._forward_analytic(ray_samples, inputs, shape, return_weights, mask)
        # 4 tetrahedron taps, sdf only (surface_model.py:138-146)
        delta = self.numerical_gradients_delta / np.sqrt(3)
        k = ops.const_tensor("taps4", lambda: torch.tensor([[1, -1, -1], [-1, -1, 1], [-1, 1, -1], [1, 1, 1]], dtype=torch.float32),
                             inputs.device)
        if hasattr(self.surface_field, "_fused") and self.surface_field._fused():
            # centre + taps as one batch through the network (geometry features only for the centre rows), a sample's
            # five evaluations in adjacent rows: row 5 i = centre (offset 0: x + 0 is x), rows 5 i + 1..4 = taps
            offs = ops.const_tensor("taps4c", lambda: torch.tensor([[0, 0, 0], [1, -1, -1], [-1, -1, 1], [-1, 1, -1], [1, 1, 1]],
                                                                   dtype=torch.float32), inputs.device)
            rows = (inputs[:, None, :] + offs[None] * delta).reshape(-1, 3)
            sdf_all, geo_feature = self.surface_field.forward_split(rows, n, group=5)
            sdf_all = sdf_all.view(n, 5)
            sdf, sdf_t = sdf_all[:, :1], sdf_all[:, 1:].t()
        else:
            taps = (inputs[None] + k[:, None, :] * delta).reshape(-1, 3)
            sdf, geo_feature = self.surface_field(inputs)
            sdf_t = self.surface_field.single_output(taps).reshape(4, n)
        want_h = bool(self.training and self.config.compute_hessian)
        gradients, hessians, normals = ops.SdfTapsFn.apply(sdf[..., 0], sdf_t, float(delta), want_h)
        sdf = sdf.view(*shape, -1)
        gradients = gradients.view(*shape, -1)
        normals = normals.view(*shape, -1)
        hessians = hessians.view(*shape, -1) if want_h else None
        outputs = {
            "sdf": sdf, "normals": normals, "gradients": gradients, "geo_feature": geo_feature,
            "hessians": hessians, "inputs": inputs, "sampled_sdf": sdf_t.view(4, *shape).permute(1, 2, 0),
            "inv_s": 1.0 / self.volume_rendering.density_fn.variance_network.get_inv_variance(),
        }
        if return_weights:
            outputs["weights"] = self.volume_rendering(ray_samples, sdf, gradients=gradients, mask=mask)
        return outputs

    def _forward_analytic(self, ray_samples, inputs, shape, return_weights, mask):
        """use_numerical_gradients=False (presets `mlp*`, surface_model.py:193-203): d sdf / d x by a forward-mode pass
        through the MLP (SDFField.forward_with_gradient) instead of autograd.grad(create_graph=True)."""
        sdf, geo_feature, gradients = self.surface_field.forward_with_gradient(inputs)
        normals = torch.nn.functional.normalize(gradients, p=2, dim=-1)
        sdf = sdf.reshape(*shape, -1)
        gradients = gradients.reshape(*shape, -1)
        outputs = {
            "sdf": sdf, "normals": normals.reshape(*shape, -1), "gradients": gradients, "geo_feature": geo_feature,
            "hessians": None, "inputs": inputs, "sampled_sdf": None,
            "inv_s": 1.0 / self.volume_rendering.density_fn.variance_network.get_inv_variance(),
        }
        if return_weights:
            outputs["weights"] = self.volume_rendering(ray_samples, sdf, gradients=gradients, mask=mask)
        return outputs

    def get_sdf(self, ray_samples: RaySamples):
        shape = ray_samples.shape
        inputs = ray_samples.frustums.get_start_positions().reshape(-1, 3)
        if self.spatial_distortion is not None:
            inputs = self.spatial_distortion(inputs)
        sdf = self.surface_field.single_output(inputs)
        return sdf.view(*shape, -1)

    def get_param_groups(self):
        groups = {"surface_field": list(self.surface_field.parameters())}
        groups.update(self.volume_rendering.get_param_groups())
        return groups

    def set_numerical_gradients_delta(self, delta: float) -> None:
        self.numerical_gradients_delta = delta

    def get_training_callbacks(self, training_callback_attributes):
        callbacks = self.volume_rendering.get_training_callbacks(training_callback_attributes) + \
            self.surface_field.get_training_callbacks(training_callback_attributes)
        if self.config.use_numerical_gradients and hasattr(self.surface_field.field, "feature_grid"):
            fg = self.surface_field.field.feature_grid
            enc = fg.config.encoding
            max_it = training_callback_attributes.trainer.max_num_iterations
            steps_per_level = min(int(max_it * fg.config.steps_per_level_ratio), int(max_it / enc.num_levels))
            growth = np.exp((np.log(enc.max_res) - np.log(enc.min_res)) / (enc.num_levels - 1))

            def set_delta(step):
                delta = 1.0 / (enc.min_res * growth ** int(step / steps_per_level))
                delta = max(1.0 / enc.max_res, delta)
                self.set_numerical_gradients_delta(delta * (fg.radius * 2.0))

            callbacks.append(TrainingCallback(where_to_run=[TrainingCallbackLocation.BEFORE_TRAIN_ITERATION],
                                              update_every_num_iters=1, func=set_delta))
        return callbacks

    def get_model_parameters(self):
        return self.surface_field.get_model_parameters()


# ---------------------------------------------------------------------------------------------
# radiance model  (ref: radiance_model.py)
# ---------------------------------------------------------------------------------------------
@dataclass
class RadianceModelConfig(InstantiateConfig):
    _target: Type = field(default_factory=lambda: RadianceModel)
    spatial_distortion: Union[None, SpatialDistortionConfig] = None
    radiance_field: BaseRadianceFieldConfig = field(default_factory=lambda: RadianceFieldConfig)
    modality_heads: Optional[Dict[str, FieldComponentConfig]] = field(default_factory=lambda: {})
    use_direction_encoding: bool = True
    direction_encoding: EncodingConfig = field(default_factory=lambda: NeRFEncodingConfig)
    use_n_dot_v: bool = False
    use_reflection_direction: bool = False
    geo_feature_dim: int = 256
    radiance_feature_dim: int = 256


class RadianceModel(nn.Module):
    def __init__(self, config: RadianceModelConfig, modalities: Dict[str, int]):
        super().__init__()
        self.config = config
        self.modalities = modalities
        self.spatial_distortion = self.config.spatial_distortion.setup() \
            if self.config.spatial_distortion is not None else None
        self.direction_encoding = self.config.direction_encoding.setup(in_dim=3)
        direction_input_dim = self.direction_encoding.get_out_dim() if self.config.use_direction_encoding else 3
        additional_input_dim = self.config.geo_feature_dim + (1 if self.config.use_n_dot_v else 0)
        self.radiance_field = self.config.radiance_field.setup(
            position_dim=3, view_direction_dim=direction_input_dim, additional_input_dim=additional_input_dim,
            output_dim=self.config.radiance_feature_dim)
        self.modality_heads = nn.ParameterDict({
            mod: self.config.modality_heads.get(mod, ModalityHeadConfig()).setup(
                input_dim=self.config.radiance_feature_dim, output_dim=self.modalities[mod])
            for mod in self.modalities})

    def forward(self, ray_samples: RaySamples, normals, geo_feature, heads=None, bounds=None):
        """`heads`: subset of modality heads to evaluate (default: all, like the reference).
        `bounds` ({mod: (first ray, last ray + 1)}): the batch holds the rays of several modalities; the trunk runs
        once over all rows, the heads over each modality's rows -> {mod: {head: [R_mod, S, C]}}; `heads` is then
        a dict {mod: [heads...]} or None (all heads for every modality)."""
        shape = ray_samples.shape
        position_input = ray_samples.frustums.get_start_positions().reshape(-1, 3)
        directions = ray_samples.frustums.directions.expand(*shape, 3).reshape(-1, 3)
        direction_input = directions
        normals = normals.reshape(-1, 3)
        if self.spatial_distortion is not None:
            position_input = self.spatial_distortion(position_input)
        additional_input = [geo_feature]
        n_dot_v = None
        if self.config.use_n_dot_v:
            n_dot_v = torch.sum(normals * -directions, dim=-1, keepdim=True)
            additional_input.append(n_dot_v)
        if self.config.use_reflection_direction:
            if n_dot_v is None:
                n_dot_v = torch.sum(normals * -directions, dim=-1, keepdim=True)
            direction_input = 2 * (n_dot_v * normals) + direction_input
        direction_piece = None
        if self.config.use_direction_encoding:
            direction_piece, direction_input = self.direction_encoding.piece(direction_input), None
        radiance_feature = self.radiance_field(positions=position_input, view_directions=direction_input,
                                               additional_inputs=additional_input, view_direction_piece=direction_piece)
        outputs = {}
        up_directions = ray_samples.frustums.up_directions.expand(*shape, 3).reshape(-1, 3)
        if bounds is not None:
            s_ = shape[-1]
            # torch.split: ONE cat in backward (per-modality indexing would zero-fill and add a full-size gradient each)
            one_head = heads is not None and all(len(heads[m]) == 1 for m in bounds)
            feats = ops.split_rows(radiance_feature, [(b - a) * s_ for a, b in bounds.values()], single_consumer=one_head)
            for (mod, (a, b)), feat in zip(bounds.items(), feats):
                rows = slice(a * s_, b * s_)
                outputs[mod] = {}
                for head in (heads[mod] if heads is not None else self.modalities):
                    out = self.modality_heads[head](feat, directions=directions[rows], up_directions=up_directions[rows])
                    outputs[mod][head] = out.view(b - a, s_, -1)
            return outputs
        for mod in (heads if heads is not None else self.modalities):
            out = self.modality_heads[mod](radiance_feature, directions=directions, up_directions=up_directions)
            outputs[mod] = out.view(*shape, -1)
        return outputs

    def get_param_groups(self):
        return {"radiance_field": list(self.radiance_field.parameters()) + list(self.modality_heads.parameters())}

    def get_training_callbacks(self, training_callback_attributes):
        return self.radiance_field.get_training_callbacks(training_callback_attributes)

    def get_model_parameters(self):
        return self.radiance_field.get_model_parameters()


# ---------------------------------------------------------------------------------------------
# background model  (ref: background_model.py)
# ---------------------------------------------------------------------------------------------
@dataclass
class BackgroundModelConfig(InstantiateConfig):
    _target: Type = field(default_factory=lambda: BackgroundModel)
    background_field: NeRFFieldConfig = field(default_factory=lambda: NeRFFieldConfig)
    modality_heads: Optional[Dict[str, FieldComponentConfig]] = field(default_factory=lambda: {})
    spatial_distortion: Union[None, SpatialDistortionConfig] = None
    radiance_feature_dim: int = 256


class BackgroundModel(nn.Module):
    def __init__(self, config: BackgroundModelConfig, modalities: Dict[str, int]):
        super().__init__()
        self.config = config
        self.modalities = modalities
        self.spatial_distortion = self.config.spatial_distortion.setup() \
            if self.config.spatial_distortion is not None else None
        self.background_field = self.config.background_field.setup(radiance_output_dim=self.config.radiance_feature_dim)
        self.modality_heads = nn.ParameterDict({
            mod: self.config.modality_heads.get(mod, ModalityHeadConfig()).setup(
                input_dim=self.config.radiance_feature_dim, output_dim=self.modalities[mod])
            for mod in self.modalities})

    def forward(self, ray_samples: RaySamples, heads=None, bounds=None):
        """`bounds` / dict `heads`: see RadianceModel.forward -> {mod: {head: [R_mod, C]}}."""
        shape = ray_samples.shape
        inputs = ray_samples.frustums.get_start_positions().reshape(-1, 3)
        directions = ray_samples.frustums.directions.expand(*shape, 3).reshape(-1, 3)
        if self.spatial_distortion is not None:
            inputs = self.spatial_distortion(inputs)
        density, radiance_feature = self.background_field(inputs, directions)
        density = density.view(*shape, -1)
        weights = ray_samples.get_weights_from_densities(density)
        outputs = {}
        up_directions = ray_samples.frustums.up_directions.expand(*shape, 3).reshape(-1, 3)
        if bounds is not None:
            s_ = shape[-1]
            feats = torch.split(radiance_feature, [(b - a) * s_ for a, b in bounds.values()], dim=0)
            wsplit = torch.split(weights[..., 0], [b - a for a, b in bounds.values()], dim=0)
            for (mod, (a, b)), feat, w in zip(bounds.items(), feats, wsplit):
                rows = slice(a * s_, b * s_)
                outputs[mod] = {}
                for head in (heads[mod] if heads is not None else self.modalities):
                    radiance = self.modality_heads[head](feat, directions=directions[rows], up_directions=up_directions[rows])
                    outputs[mod][head] = ops.CompositeFn.apply(w, radiance.view(b - a, s_, -1), None)
            return outputs
        for mod in (heads if heads is not None else self.modalities):
            radiance = self.modality_heads[mod](radiance_feature, directions=directions, up_directions=up_directions)
            outputs[mod] = ops.CompositeFn.apply(weights[..., 0], radiance.view(*shape, -1), None)
        return outputs

    def get_param_groups(self):
        return {"background_field": list(self.background_field.parameters()) + list(self.modality_heads.parameters())}

    def get_training_callbacks(self, training_callback_attributes):
        return []

    def get_model_parameters(self):
        return {}


# ---------------------------------------------------------------------------------------------
# renderers  (ref: renderers.py)
# ---------------------------------------------------------------------------------------------
class RadianceRenderer(nn.Module):
    """ref: renderers.py:149-174"""

    @classmethod
    def render(cls, radiance_values, weights, background_color):
        return ops.CompositeFn.apply(weights[..., 0], radiance_values, background_color)

    def forward(self, *args, **kwargs):
        return self.render(*args, **kwargs)


@dataclass
class RendererConfig(InstantiateConfig):
    _target: Type = field(default_factory=lambda: Renderer)
    renderers: Dict[str, Any] = field(default_factory=lambda: {"rgb": RadianceRenderer})
    background_color: Any = "None"


class Renderer:
    """ref: renderers.py:48-136.  `weights` already carries zeros for rays outside the sphere, so no masked
    scatter is needed; `mask` is accepted for signature compatibility."""

    def __init__(self, config):
        self.config = config
        for element, renderer_class in self.config.renderers.items():
            setattr(self, element, renderer_class())

    def prepare_background(self, background_samples, n_rays, n_channels, device):
        if self.config.background_color == "None" and background_samples is not None:
            return background_samples
        if self.config.background_color == "white":
            return torch.ones((n_rays, n_channels), device=device)
        if self.config.background_color == "black":
            return torch.zeros((n_rays, n_channels), device=device)
        if self.config.background_color == "random":
            return torch.rand((n_rays, n_channels), device=device)
        raise ValueError(f"Background color {self.config.background_color} not supported.")

    def render(self, weights, data_fields: Dict[str, Any], mask=None) -> Dict[str, torch.Tensor]:
        outputs = {}
        n_rays = weights.shape[0]
        normals, depth_samples = None, None
        for mod, value in data_fields.items():
            if mod == "background":
                continue
            if mod in self.config.renderers:
                n_channels = value.shape[-1]
                bg = data_fields.get("background")
                bg = self.prepare_background(bg[mod] if bg is not None else None, n_rays, n_channels, weights.device)
                outputs[mod] = getattr(self, mod)(value, weights, bg)
            elif mod == "normals":
                normals = value
            elif mod == "depth":
                depth_samples = value
            else:
                outputs[mod] = ops.CompositeFn.apply(weights[..., 0], value, None)
        if normals is not None and depth_samples is not None:
            starts, ends = depth_samples.frustums.starts[..., 0], depth_samples.frustums.ends[..., 0]
            rn, rd, ra = ops.composite_aux(weights[..., 0], normals, starts, ends)
            outputs["normals"] = rn
            # global clip of renderers.py:214-215 over the rays that hit the sphere
            steps = (starts + ends).detach() / 2
            if mask is not None:
                m = mask.bool()[:, None]
                big = torch.finfo(steps.dtype).max
                lo = torch.where(m, steps, torch.full_like(steps, big)).min()
                hi = torch.where(m, steps, torch.full_like(steps, -big)).max()
                outputs["depth"] = torch.where(m, torch.minimum(torch.maximum(rd, lo), hi), torch.zeros_like(rd))
            else:
                outputs["depth"] = torch.clip(rd, steps.min(), steps.max())
            outputs["accumulation"] = ra
        else:
            outputs["accumulation"] = weights.detach().sum(dim=-2)
        return outputs


# ---------------------------------------------------------------------------------------------
# losses  (ref: losses.py, engine/schedulers.py:320-346)
# ---------------------------------------------------------------------------------------------
@dataclass
class CurvatureLossWarmUpSchedulerConfig(InstantiateConfig):
    _target: Type = field(default_factory=lambda: CurvatureLossWarmUpScheduler)
    warm_up_ratio: float = 0.1


class CurvatureLossWarmUpScheduler:
    def __init__(self, config, num_iterations, grow_factor, level_init, num_levels, steps_per_level, optimizer=None):
        self.config = config
        self.warm_up_end = int(num_iterations * config.warm_up_ratio)

        def func(step):
            if step < self.warm_up_end:
                return step / self.warm_up_end
            level = int(step / steps_per_level) + 1
            level = min(max(level, level_init), num_levels)
            return np.reciprocal(grow_factor ** (level - 1))

        self.get_update_factor = func


@dataclass
class LossConfig(InstantiateConfig):
    _target: Type = field(default_factory=lambda: Loss)
    loss: str = "L1"
    weight: float = 1.0
    scheduler: Any = None
    per_channel_probability: List[float] = None


@dataclass
class EikonalLossConfig(LossConfig):
    _target: Type = field(default_factory=lambda: EikonalLoss)
    loss: str = "MSE"
    weight: float = 0.1


@dataclass
class CurvatureLossConfig(LossConfig):
    _target: Type = field(default_factory=lambda: CurvatureLoss)
    loss: str = "L1"
    weight: float = 5e-4


@dataclass
class SkipSaturationLossConfig(LossConfig):
    _target: Type = field(default_factory=lambda: SkipSaturationLoss)
    saturation_threshold: float = 0.9999


class Loss(nn.Module):
    """Radiance L1 loss; with `pixel_coords` + `mosaick_pattern` the channel select of
    raw_pipeline.py:112-122 is fused into the same kernel.  ref: losses.py:77-105"""
    saturation_threshold = float("inf")

    def __init__(self, config: LossConfig, reduction: str = "mean", **kwargs):
        super().__init__()
        self.config = config
        if self.config.loss != "L1" or reduction != "mean":
            raise ValueError("the B200 radiance loss implements L1 with mean reduction")
        self.channel_probability = None
        if self.config.per_channel_probability is not None:
            self.channel_probability = torch.tensor(self.config.per_channel_probability)      # CPU, like the reference
        if self.config.scheduler is not None and "num_iterations" in kwargs:
            self.scheduler = self.config.scheduler.setup(num_iterations=kwargs["num_iterations"])

    def _weight(self, step):
        weight = self.config.weight
        if self.config.scheduler is not None:
            weight *= self.scheduler.get_update_factor(step)
        return weight

    def _decimated(self, output, target, channel_draw=None):
        """per_channel_probability (preset grid_decimated, losses.py:87-105).  The reference draws one channel index
        per pixel with torch.multinomial and indexes `output[arange(n), indexes.view(-1, 1)]`: index shapes [n] and [n, 1]
        BROADCAST to [n, n], so element (i, j) is pixel j at the channel drawn for pixel i, and the mean over the n x n
        matrix is  (1/n) sum_j sum_c f_c |out[j, c] - tgt[j, c]|  with f_c = the fraction of this batch's draws that
        picked channel c.  That is what is computed here (small [R, C] tensors: element-wise torch ops; the draw itself is
        the reference's CPU torch.multinomial call, same generator state -> same indices).  `channel_draw`: optional
        externally supplied indices [n] (tests)."""
        n, c = output.shape
        if len(self.channel_probability) != c:
            raise ValueError(f"per_channel_probability has {len(self.channel_probability)} entries for {c} channels")
        if channel_draw is None:
            channel_draw = torch.multinomial(self.channel_probability, n, replacement=True)
        freq = torch.bincount(channel_draw.reshape(-1).cpu(), minlength=c).to(torch.float32) / float(n)
        freq = freq.to(output.device)
        if self.saturation_threshold != float("inf"):
            # SkipSaturationLoss.forward (losses.py:158-164) without a host sync: first saturated target by index
            flat = target.reshape(-1)
            idx = ops.first_saturated(target, self.saturation_threshold)[0]
            has = idx < flat.numel()
            value = flat[idx.clamp(max=flat.numel() - 1)]
            output = torch.where((target > self.saturation_threshold) & has, value, output)
        return ((output - target).abs() * freq[None, :]).sum() / float(n)

    def forward(self, output, target, step, pixel_coords=None, mosaick_pattern=None, channel_draw=None, **kwargs):
        if self.channel_probability is not None:
            if mosaick_pattern is not None:
                raise ValueError("per_channel_probability applies to demosaicked targets (preset grid_decimated)")
            return self._decimated(output, target, channel_draw), self._weight(step)
        sat_index = None
        if self.saturation_threshold != float("inf"):
            sat_index = ops.first_saturated(target, self.saturation_threshold)
        if mosaick_pattern is not None:
            ph, pw = mosaick_pattern.shape
            pat = mosaick_pattern.to(device=output.device, dtype=torch.int32).contiguous()
            loss, _ = ops.MosaickL1Fn.apply(output, target.reshape(-1), pixel_coords, pat, ph, pw,
                                            self.saturation_threshold, sat_index)
        else:
            loss, _ = ops.MosaickL1Fn.apply(output, target, None, None, 1, 1, self.saturation_threshold, sat_index)
        return loss, self._weight(step)


class SkipSaturationLoss(Loss):
    """ref: losses.py:152-164"""

    def __init__(self, config: SkipSaturationLossConfig, num_iterations: int = 0, **kwargs):
        super().__init__(config, num_iterations=num_iterations)
        self.saturation_threshold = float(self.config.saturation_threshold)


class _GeometryLoss(nn.Module):
    def __init__(self, config, **kwargs):
        super().__init__()
        self.config = config

    def _weight(self, step):
        weight = self.config.weight
        if self.config.scheduler is not None:
            weight *= self.scheduler.get_update_factor(step)
        return weight


class EikonalLoss(_GeometryLoss):
    def __init__(self, config: EikonalLossConfig, num_iterations: int, **kwargs):
        super().__init__(config)
        if self.config.scheduler is not None:
            self.scheduler = self.config.scheduler.setup(num_iterations=num_iterations)


class CurvatureLoss(_GeometryLoss):
    """ref: losses.py:121-150"""

    def __init__(self, config: CurvatureLossConfig, num_iterations: int, **kwargs):
        super().__init__(config)
        mp = kwargs.get("model").get_model_parameters()
        steps_per_level = int(num_iterations * mp["steps_per_level_ratio"])
        self.steps_per_level = min(steps_per_level, int(num_iterations / mp["num_levels"]))
        self.grow_factor = np.exp((np.log(mp["max_res"]) - np.log(mp["min_res"])) / (mp["num_levels"] - 1))
        if self.config.scheduler is not None:
            self.scheduler = self.config.scheduler.setup(
                num_iterations=num_iterations, grow_factor=self.grow_factor, level_init=mp["level_init"],
                num_levels=mp["num_levels"], steps_per_level=self.steps_per_level)


@dataclass
class LossManagerConfig(InstantiateConfig):
    _target: Type = field(default_factory=lambda: LossManager)
    radiance_losses: Dict[str, LossConfig] = field(default_factory=lambda: {"rgb": LossConfig()})
    geometry_losses: Dict[str, LossConfig] = field(default_factory=lambda: {"eikonal_loss": EikonalLossConfig()})
    additional_losses: Dict[str, LossConfig] = field(default_factory=lambda: {})


class LossManager:
    """ref: losses.py:180-265.  `mosaick_patterns` ({mod: int tensor [ph,pw]}) switches on the fused
    channel select for raw pipelines; `outputs[mod][mod]` is then the full [R,C] rendering."""

    def __init__(self, config: LossManagerConfig, modalities: List[str], num_iterations: int, **kwargs):
        self.config = config
        self.modalities = modalities
        for element in self.modalities:
            setattr(self, element, self.config.radiance_losses[element].setup(num_iterations=num_iterations, **kwargs))
        for element, cfg in self.config.geometry_losses.items():
            setattr(self, element, cfg.setup(num_iterations=num_iterations, **kwargs))

    @property
    def graph_capturable(self) -> bool:
        """False when a loss draws random channels on the host every step (per_channel_probability): such a step cannot
        be replayed from a CUDA graph."""
        return all(getattr(getattr(self, mod), "channel_probability", None) is None for mod in self.modalities)

    def weights(self, step):
        """Loss weights in effect at `step` (the step-dependent scalars compute_loss multiplies in)."""
        out = [getattr(self, mod)._weight(step) for mod in self.modalities]
        out += [getattr(self, name)._weight(step) for name in self.config.geometry_losses]
        return out

    def compute_loss(self, outputs, targets, pixel_coords, step, eval_step=False, mosaick_patterns=None,
                     loss_scales=None, geometry_count=None):
        """`loss_scales` ({mod: n_batch / n_global}) and `geometry_count` (device fp32 [1], the global number of
        in-sphere samples): set when the batch is one shard of a step's global batch (pipelines.ShardPlan) — every term
        of the total is then `local sum / global count`, so shards add up to the reference's full-batch loss.  The
        entries of `losses` stay the batch's own means."""
        losses = {}
        total_loss = 0.0
        for mod in self.modalities:
            loss_func = getattr(self, mod)
            pattern = mosaick_patterns[mod] if mosaick_patterns is not None else None
            loss, weight = loss_func(outputs[mod][mod], targets[mod], step,
                                     pixel_coords=pixel_coords[mod] if pixel_coords is not None else None,
                                     mosaick_pattern=pattern, eval_step=eval_step)
            losses[mod] = loss
            if weight != 1:
                losses[mod + "_weight"] = weight
            if loss_scales is not None:
                weight = weight * loss_scales[mod]
            total_loss = total_loss + weight * loss
        if not eval_step and self.config.geometry_losses:
            grads = [outputs[mod]["gradients"] for mod in self.modalities]
            hess = [outputs[mod].get("hessians") for mod in self.modalities]
            masks = [outputs[mod]["ray_mask"] for mod in self.modalities]
            g = torch.cat(grads, dim=0)
            h = torch.cat(hess, dim=0) if all(x is not None for x in hess) else None
            m = torch.cat(masks, dim=0) if all(x is not None for x in masks) else None
            eik, curv = ops.GeometryLossFn.apply(g, h, m, geometry_count)
            for loss_name in self.config.geometry_losses:
                loss_fn = getattr(self, loss_name)
                if loss_name == "eikonal_loss":
                    loss = eik
                elif loss_name == "curvature_loss":
                    loss = curv
                else:
                    raise NotImplementedError
                weight = loss_fn._weight(step)
                losses[loss_name] = loss
                losses[loss_name + "_weight"] = weight
                total_loss = total_loss + weight * loss
        return losses, total_loss
