"""Mirror of `models/base_model.py` and of the method presets the hot path is quoted on.
ref: src/models/base_model.py:34-199, src/configs/method_configs.py:63-445, confs/*.yaml
"""
import copy
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Type

import torch

from .cameras import RayBundle
from .configs import InstantiateConfig, update_config
from .field_components import (FeatureGridAndMLPConfig, FeatureGridConfig, HashEncodingConfig, MLPConfig,
                               ModalityHeadConfig, NeRFEncodingConfig, PolarizationHeadConfig, SceneContractionConfig,
                               SHEncodingConfig)
from .fields import NeRFFieldConfig, RadianceFieldConfig, SDFFieldConfig
from .model_components import (BackgroundModelConfig, ColliderInstancer, CurvatureLossConfig,
                               CurvatureLossWarmUpSchedulerConfig, EikonalLossConfig, LinearDisparitySamplerConfig,
                               LossConfig, LossManagerConfig, NeuSDensityConfig, NeuSSamplerConfig,
                               NeuSVolumeRenderingConfig, RadianceModelConfig, RadianceRenderer, RendererConfig,
                               SamplerConfig, SceneBox, SkipSaturationLossConfig, SurfaceModelConfig)

MODALITY_CHANNELS = {"rgb": 3, "infrared": 1, "mono": 1, "polarization": 4, "multispectral": 9}
# ref: preprocessing/preprocess_mmsdata.py:43-47
MOSAICK_PATTERNS = {
    "rgb": [[1, 2], [0, 1]],
    "infrared": [[0]],
    "mono": [[0]],
    "polarization": [[2, 1], [3, 0]],
    "multispectral": [[4, 5, 6], [2, 1, 0], [3, 8, 7]],
}


@dataclass
class BaseModelConfig(InstantiateConfig):
    _target: Type = field(default_factory=lambda: BaseModel)
    ray_sampler: SamplerConfig = field(default_factory=lambda: SamplerConfig)
    background_ray_sampler: SamplerConfig = field(default_factory=lambda: SamplerConfig)
    surface_model: SurfaceModelConfig = field(default_factory=lambda: SurfaceModelConfig)
    radiance_model: RadianceModelConfig = field(default_factory=lambda: RadianceModelConfig)
    background_model: BackgroundModelConfig = field(default_factory=lambda: BackgroundModelConfig)
    renderer: RendererConfig = field(default_factory=lambda: RendererConfig)
    use_background_model: bool = True
    render_all_heads: bool = True
    """True: every modality's rays go through all M heads like the reference (radiance_model.py:143-149).
    False: only the head of the bundle's own modality is evaluated — the only one the losses read
    (losses.py:225); identical loss and gradients."""
    batch_modalities: bool = True
    """True: the rays of all modalities run through the shared networks (sampler's SDF queries, SDF field + taps,
    radiance trunk, background field) as ONE batch; only the modality heads, the compositing and the losses see
    per-modality row ranges.  Every operator is row-independent, so the outputs equal those of the reference's
    per-modality loop (base_model.py:102-159) up to the summation order of the parameter gradients.
    False: the reference's loop."""


class BaseModel(torch.nn.Module):
    """ref: base_model.py:54-199.  forward(ray_bundles, rand=None) -> {mod: {head..., normals, depth,
    accumulation, [gradients, hessians, inv_s, ray_mask]}}.  `rand` optionally carries the random draws
    the reference would make ({"uniform": {mod: [R,1]}, "pdf": {mod: [U x [R,1]]}, "background": {mod: [R,S_bg+1]}})."""

    def __init__(self, config: BaseModelConfig, scene_box: SceneBox, modalities: Dict[str, int]):
        super().__init__()
        self.config = config
        self.modalities = modalities
        self.ray_sampler = self.config.ray_sampler.setup()
        self.collider = ColliderInstancer(scene_box)
        self.surface_model = self.config.surface_model.setup()
        self.radiance_model = self.config.radiance_model.setup(modalities=modalities)
        if self.config.use_background_model:
            self.background_ray_sampler = self.config.background_ray_sampler.setup()
            self.background_model = self.config.background_model.setup(modalities=modalities)
        self.renderer = self.config.renderer.setup()

    def forward(self, ray_bundles, rand: Optional[dict] = None):
        rand = rand or {}
        if self.config.batch_modalities and sum(rb is not None for rb in ray_bundles.values()) > 1:
            return self._forward_batched(ray_bundles, rand)
        masks = self.collider.update_ray_bundles(ray_bundles)
        sampler_out = self.ray_sampler(ray_bundles, sdf_fn=self.surface_model.get_sdf, rand=rand)
        samples_per_modality = sampler_out["ray_samples_per_modality"]
        background_samples = {}
        if self.config.use_background_model:
            self.collider.update_ray_bundles_for_background(ray_bundles)
            background_samples = self.background_ray_sampler(ray_bundles, rand=rand.get("background"))

        outputs = {}
        for mod in samples_per_modality.keys():
            samples = samples_per_modality.get(mod)
            if samples is None:
                outputs[mod] = None
                continue
            heads = None if self.config.render_all_heads else [mod]
            mask = masks.get(mod)
            background_outputs = None
            if self.config.use_background_model:
                background_outputs = self.background_model(background_samples[mod], heads=heads)
            geometry = self.surface_model(samples, mask=mask)
            radiance = self.radiance_model(ray_samples=samples, normals=geometry["normals"].detach(),
                                           geo_feature=geometry["geo_feature"], heads=heads)
            renderer_input = dict(radiance)
            renderer_input.update({"normals": geometry["normals"], "depth": samples, "background": background_outputs})
            modality_outputs = self.renderer.render(geometry["weights"], renderer_input, mask)
            if self.training:
                modality_outputs.update({"gradients": geometry["gradients"], "hessians": geometry["hessians"],
                                         "inv_s": geometry["inv_s"], "ray_mask": mask})
            outputs[mod] = modality_outputs
        return outputs

    def _forward_batched(self, ray_bundles, rand):
        """All modalities as one ray batch through the shared networks (see BaseModelConfig.batch_modalities)."""
        ALL = "__all__"
        mods = [m for m, rb in ray_bundles.items() if rb is not None]
        bundles = [ray_bundles[m] for m in mods]
        counts = [len(rb) for rb in bundles]
        bounds, off = {}, 0
        for m, c in zip(mods, counts):
            bounds[m] = (off, off + c)
            off += c

        def cat_attr(name):
            vals = [getattr(rb, name) for rb in bundles]
            return torch.cat(vals, 0) if all(v is not None for v in vals) else None

        big = RayBundle(camera_indices=cat_attr("camera_indices"), origins=cat_attr("origins"),
                        directions=cat_attr("directions"), up_directions=cat_attr("up_directions"),
                        pixel_area=cat_attr("pixel_area"), directions_norm=cat_attr("directions_norm"))

        def cat_rand(key):
            d = rand.get(key)
            if not d or all(d.get(m) is None for m in mods):
                return None
            if any(d.get(m) is None for m in mods):
                raise ValueError(f"rand['{key}'] must be given for every modality or for none")
            if isinstance(d[mods[0]], (list, tuple)):
                return {ALL: [torch.cat([d[m][i] for m in mods], 0) for i in range(len(d[mods[0]]))]}
            return {ALL: torch.cat([d[m] for m in mods], 0)}

        rand_all = {k: cat_rand(k) for k in ("uniform", "pdf", "bins")}
        mask = self.collider.update_ray_bundles({ALL: big})[ALL]
        samples = self.ray_sampler({ALL: big}, sdf_fn=self.surface_model.get_sdf, rand=rand_all)["ray_samples_per_modality"][ALL]
        background = None
        if self.config.use_background_model:
            self.collider.update_ray_bundles_for_background({ALL: big})
            bg_samples = self.background_ray_sampler({ALL: big}, rand=cat_rand("background"))[ALL]
            heads = None if self.config.render_all_heads else {m: [m] for m in mods}
            background = self.background_model(bg_samples, heads=heads, bounds=bounds)
        geometry = self.surface_model(samples, mask=mask)
        heads = None if self.config.render_all_heads else {m: [m] for m in mods}
        radiance = self.radiance_model(ray_samples=samples, normals=geometry["normals"].detach(),
                                       geo_feature=geometry["geo_feature"], heads=heads, bounds=bounds)
        outputs = {m: None for m in ray_bundles}
        # torch.split: one cat in backward instead of a zero-filled full-size gradient (+ add) per modality
        sizes = [bounds[m][1] - bounds[m][0] for m in mods]
        split = {k: torch.split(geometry[k], sizes, dim=0) if geometry.get(k) is not None else None
                 for k in ("weights", "normals", "gradients", "hessians")}
        for i, m in enumerate(mods):
            a, b = bounds[m]
            renderer_input = dict(radiance[m])
            renderer_input.update({"normals": split["normals"][i], "depth": samples.slice_rays(a, b),
                                   "background": background[m] if background is not None else None})
            out = self.renderer.render(split["weights"][i], renderer_input, mask[a:b] if mask is not None else None)
            if self.training:
                out.update({"gradients": split["gradients"][i],
                            "hessians": split["hessians"][i] if split["hessians"] is not None else None,
                            "inv_s": geometry["inv_s"], "ray_mask": mask[a:b] if mask is not None else None})
            outputs[m] = out
        return outputs

    def get_param_groups(self):
        groups = {}
        groups.update(self.surface_model.get_param_groups())
        groups.update(self.radiance_model.get_param_groups())
        groups.update(self.ray_sampler.get_param_groups())
        if self.config.use_background_model:
            groups.update(self.background_model.get_param_groups())
        return groups

    def get_training_callbacks(self, training_callback_attributes):
        callbacks = self.surface_model.get_training_callbacks(training_callback_attributes) + \
            self.radiance_model.get_training_callbacks(training_callback_attributes) + \
            self.ray_sampler.get_training_callbacks(training_callback_attributes)
        if self.config.use_background_model:
            callbacks += self.background_model.get_training_callbacks(training_callback_attributes)
        return callbacks

    def get_model_parameters(self):
        parameters = {}
        parameters.update(self.surface_model.get_model_parameters())
        if self.config.use_background_model:
            parameters.update(self.background_model.get_model_parameters())
        return parameters

    def set_schedule_state(self, level: int = 16, delta: float = 2.0 / 1024, anneal: float = 1.0):
        """What the BEFORE_TRAIN_ITERATION callbacks set each step (feature_structures.py:97-108,
        surface_model.py:266-271, volume_rendering.py:227-230), for callers that drive the model directly."""
        self.surface_model.volume_rendering.set_cos_anneal_ratio(anneal)
        self.surface_model.set_numerical_gradients_delta(delta)
        for m in self.modules():
            if hasattr(m, "update_mask"):
                m.update_mask(level)


# ---------------------------------------------------------------------------------------------
# method presets (the model / loss part of method_configs.py)
# ---------------------------------------------------------------------------------------------
def _head(hidden, out_act):
    return MLPConfig(num_layers=3, hidden_dim=hidden, out_activation=out_act, weight_norm=True)


def grid_model_config() -> BaseModelConfig:
    """ref: method_configs.py:85-240 (preset `grid`; `grid_raw` deep-copies it, :360-378)"""
    heads = {
        "rgb": ModalityHeadConfig(field=_head(64, "Sigmoid")),
        "infrared": ModalityHeadConfig(field=_head(64, "Sigmoid")),
        "mono": ModalityHeadConfig(field=_head(64, "Sigmoid")),
        "polarization": PolarizationHeadConfig(field=_head(256, "None")),
        "multispectral": ModalityHeadConfig(field=_head(64, "Sigmoid")),
    }
    pe6 = lambda: NeRFEncodingConfig(num_frequencies=6, min_freq_exp=0.0, max_freq_exp=5, include_input=True)
    return BaseModelConfig(
        ray_sampler=NeuSSamplerConfig(num_samples=32, num_samples_importance=32),
        background_ray_sampler=LinearDisparitySamplerConfig(),
        surface_model=SurfaceModelConfig(
            use_numerical_gradients=True,
            surface_field=SDFFieldConfig(
                field=FeatureGridAndMLPConfig(
                    feature_grid=FeatureGridConfig(encoding=HashEncodingConfig(max_res=1024), coarse_to_fine=True, radius=1),
                    mlp_head=MLPConfig(num_layers=3, activation="Softplus", activation_params={"beta": 100},
                                       out_activation="None", geometric_init=True, weight_norm=True)),
                use_position_encoding=True, position_encoding=pe6()),
            volume_rendering=NeuSVolumeRenderingConfig(density_fn=NeuSDensityConfig()),
            compute_hessian=True),
        radiance_model=RadianceModelConfig(
            radiance_field=RadianceFieldConfig(
                base_field=FeatureGridAndMLPConfig(
                    feature_grid=FeatureGridConfig(encoding=HashEncodingConfig(max_res=1024), coarse_to_fine=True, radius=1),
                    mlp_head=MLPConfig(num_layers=3, hidden_dim=256, out_activation="ReLU", weight_norm=True))),
            radiance_feature_dim=256, modality_heads=heads, use_direction_encoding=True,
            direction_encoding=SHEncodingConfig(degree=4), use_reflection_direction=True, use_n_dot_v=True),
        background_model=BackgroundModelConfig(
            background_field=NeRFFieldConfig(
                base_field=MLPConfig(activation="ReLU", hidden_dim=256, num_layers=4, out_activation="ReLU", weight_norm=True),
                head_field=MLPConfig(num_layers=4, out_activation="ReLU", weight_norm=True),
                use_position_encoding=True, position_encoding=pe6(), use_direction_encoding=True,
                direction_encoding=NeRFEncodingConfig(num_frequencies=4, min_freq_exp=0.0, max_freq_exp=3, include_input=True)),
            radiance_feature_dim=128, modality_heads={"polarization": PolarizationHeadConfig()},
            spatial_distortion=SceneContractionConfig(order=float("inf"))),
        renderer=RendererConfig(renderers={m: RadianceRenderer for m in MODALITY_CHANNELS}),
    )


def grid_loss_config() -> LossManagerConfig:
    """ref: method_configs.py:241-259"""
    return LossManagerConfig(
        radiance_losses={"rgb": LossConfig(), "mono": LossConfig(), "multispectral": LossConfig(),
                         "infrared": LossConfig(), "polarization": SkipSaturationLossConfig(saturation_threshold=0.9980)},
        geometry_losses={"eikonal_loss": EikonalLossConfig(),
                         "curvature_loss": CurvatureLossConfig(scheduler=CurvatureLossWarmUpSchedulerConfig(warm_up_ratio=0.1))})


# the `pipeline.model` part of confs/grid.yaml == confs/grid_raw.yaml (lines 62-107)
GRID_YAML_MODEL = {
    "ray_sampler": {"num_samples": 32, "num_samples_importance": 32},
    "background_ray_sampler": {"num_samples": 16},
    "surface_model": {
        "use_numerical_gradients": True, "numerical_gradient_taps": 4,
        "surface_field": {
            "field": {"feature_grid": {"encoding": {"max_res": 1024}, "coarse_to_fine": True, "radius": 1.0},
                      "mlp_head": {"hidden_dim": 256, "geometric_init": True, "weight_norm": True, "geometric_init_bias": 0.4}},
            "use_position_encoding": True}},
    "radiance_model": {
        "radiance_field": {"base_field": {"feature_grid": {"encoding": {"max_res": 1024}, "coarse_to_fine": True, "radius": 1.0},
                                          "mlp_head": {"hidden_dim": 256, "weight_norm": True}}},
        "use_reflection_direction": False, "use_n_dot_v": True},
    "background_model": {"background_field": {"base_field": {"output_dim": 256, "weight_norm": True},
                                              "head_field": {"hidden_dim": 256, "weight_norm": True}}},
}


# the `pipeline.model` part of confs/grid_raw_rgb_all_views_pol_10_views.yaml (lines 62-101): as above, but the
# background base field is the preset's hash grid (no `base_field` override)
GRID_BG_YAML_MODEL = {k: v for k, v in GRID_YAML_MODEL.items() if k != "background_model"}
GRID_BG_YAML_MODEL["background_model"] = {"background_field": {"head_field": {"hidden_dim": 256, "weight_norm": True}}}

# the `pipeline.model` part of confs/mlp.yaml == confs/mlp_raw.yaml (lines 59-88)
MLP_YAML_MODEL = {
    "ray_sampler": {"num_samples": 32, "num_samples_importance": 32},
    "background_ray_sampler": {"num_samples": 16},
    "surface_model": {"use_numerical_gradients": False,
                      "surface_field": {"field": {"weight_norm": True, "geometric_init_bias": 0.4}, "use_position_encoding": True}},
    "radiance_model": {"radiance_field": {"base_field": {"weight_norm": True}}, "use_reflection_direction": False, "use_n_dot_v": True},
    "background_model": {"background_field": {"base_field": {"output_dim": 256, "weight_norm": True},
                                              "head_field": {"hidden_dim": 256, "weight_norm": True}}},
}

GRID_PRESETS = ("grid", "grid_raw", "grid_unbalanced", "grid_raw_unbalanced", "grid_decimated")
MLP_PRESETS = ("mlp", "mlp_raw")
GRID_BG_PRESETS = ("grid_raw_grid_bg_unbalanced",)


def grid_bg_model_config() -> BaseModelConfig:
    """ref: method_configs.py:428-445 (preset `grid_raw_grid_bg_unbalanced`): `grid_raw` with the background estimated by
    a multi-resolution hash grid of radius 2 (+ 3-layer MLP) instead of an MLP, the background modality heads copied from
    the radiance model's, and 256 background radiance features."""
    cfg = grid_model_config()
    cfg.background_model.background_field.base_field = FeatureGridAndMLPConfig(
        output_dim=256,
        feature_grid=FeatureGridConfig(encoding=HashEncodingConfig(max_res=1024), coarse_to_fine=True, radius=2),
        mlp_head=MLPConfig(num_layers=3, out_activation="ReLU"))
    cfg.background_model.modality_heads = copy.deepcopy(cfg.radiance_model.modality_heads)
    cfg.background_model.radiance_feature_dim = 256
    return cfg


def mlp_model_config() -> BaseModelConfig:
    """ref: method_configs.py:302-356 (preset `mlp`; `mlp_raw` deep-copies it, :382-400): 8 x 256 MLPs with a skip
    connection at layer 4 for the SDF (Softplus beta = 100, geometric init) and the radiance trunk, SDF gradients by
    autograd (use_numerical_gradients=False), no Hessian."""
    cfg = grid_model_config()
    pe6 = lambda: NeRFEncodingConfig(num_frequencies=6, min_freq_exp=0.0, max_freq_exp=5, include_input=True)
    cfg.surface_model = SurfaceModelConfig(
        use_numerical_gradients=False,
        surface_field=SDFFieldConfig(
            field=MLPConfig(activation="Softplus", num_layers=8, hidden_dim=256, activation_params={"beta": 100},
                            out_activation="None", skip_connections=(4,), geometric_init=True, weight_norm=True),
            use_position_encoding=True, position_encoding=pe6()),
        volume_rendering=NeuSVolumeRenderingConfig(density_fn=NeuSDensityConfig()),
        compute_hessian=False)
    heads = copy.deepcopy(cfg.radiance_model.modality_heads)
    cfg.radiance_model = RadianceModelConfig(
        radiance_field=RadianceFieldConfig(
            base_field=MLPConfig(activation="ReLU", num_layers=8, hidden_dim=256, out_activation="ReLU", skip_connections=(4,),
                                 weight_norm=True)),
        radiance_feature_dim=256, modality_heads=heads, use_direction_encoding=True,
        direction_encoding=SHEncodingConfig(degree=4), use_reflection_direction=True, use_n_dot_v=True)
    return cfg


def mlp_loss_config() -> LossManagerConfig:
    """ref: method_configs.py:357-359 (the `mlp*` presets keep the eikonal loss only)"""
    cfg = grid_loss_config()
    cfg.geometry_losses = {"eikonal_loss": EikonalLossConfig()}
    return cfg


def decimated_loss_config() -> LossManagerConfig:
    """ref: method_configs.py:410-424 (preset `grid_decimated`: `grid` supervising one random channel per pixel)"""
    cfg = grid_loss_config()
    cfg.radiance_losses["rgb"].per_channel_probability = [0.25, 0.5, 0.25]
    cfg.radiance_losses["multispectral"].per_channel_probability = [0.1111] * 9
    cfg.radiance_losses["polarization"].per_channel_probability = [0.25, 0.25, 0.25, 0.25]
    return cfg


def loss_config_for(preset: str) -> LossManagerConfig:
    if preset in MLP_PRESETS:
        return mlp_loss_config()
    return decimated_loss_config() if preset == "grid_decimated" else grid_loss_config()


def build_model(preset: str = "grid_raw", modalities: Optional[Dict[str, int]] = None, yaml_model: Optional[dict] = None,
                interpolation: str = "Linear", direction_encoding: str = "nerf", log2_hashmap_size: Optional[int] = None,
                num_samples: Optional[int] = None, num_samples_importance: Optional[int] = None,
                bg_samples: Optional[int] = None, render_all_heads: bool = True, seed: Optional[int] = 654824,
                batch_modalities: bool = True):
    """Builds the BaseModel of a preset after the YAML overrides, with the tcnn-free substitutions the
    pinned oracle uses (Linear interpolation, NeRF direction encoding).  Construction order and
    RNG consumption match the reference, so the same torch seed gives the same initial parameters."""
    if preset in GRID_PRESETS:
        cfg, default_yaml = grid_model_config(), GRID_YAML_MODEL
    elif preset in GRID_BG_PRESETS:
        cfg, default_yaml = grid_bg_model_config(), GRID_BG_YAML_MODEL
    elif preset in MLP_PRESETS:
        cfg, default_yaml = mlp_model_config(), MLP_YAML_MODEL
    else:
        raise ValueError(f"preset '{preset}' is not on the B200 hot path")
    update_config_dict = copy.deepcopy(default_yaml if yaml_model is None else yaml_model)

    class _Holder:  # update_config works on an object with a `model` attribute
        pass

    holder = _Holder()
    holder.model = cfg
    update_config(holder, {"model": update_config_dict})
    grids = [f.feature_grid for f in (cfg.surface_model.surface_field.field, cfg.radiance_model.radiance_field.base_field,
                                      cfg.background_model.background_field.base_field) if hasattr(f, "feature_grid")]
    for fg in grids:
        fg.encoding.interpolation = interpolation
        if log2_hashmap_size is not None:
            fg.encoding.log2_hashmap_size = log2_hashmap_size
    if direction_encoding == "nerf":
        cfg.radiance_model.direction_encoding = NeRFEncodingConfig(num_frequencies=4, max_freq_exp=3)
    if num_samples is not None:
        cfg.ray_sampler.num_samples = num_samples
    if num_samples_importance is not None:
        cfg.ray_sampler.num_samples_importance = num_samples_importance
    if bg_samples is not None:
        cfg.background_ray_sampler.num_samples = bg_samples
    cfg.render_all_heads = render_all_heads
    cfg.batch_modalities = batch_modalities
    if modalities is None:
        modalities = dict(MODALITY_CHANNELS)
    if seed is not None:
        torch.manual_seed(seed)
    scene_box = SceneBox(radius=1.0, collider_type="sphere")
    return cfg.setup(scene_box=scene_box, modalities=modalities)
