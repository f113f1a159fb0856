"""TEST INFRASTRUCTURE — not shipped, not imported by the product.

Imports the *unmodified* reference (`/root/reference/src`) in this container on CPU, following the
recipe pinned in SURVEY.md §8(c): six import shims (`oracle/refshim`), `configs.configs` imported
first, model built from `method_configs[preset].pipeline.model` after the YAML overrides, tcnn-free
substitutions (Linear interpolation, NeRF direction encoding), schedule state set by hand.

Used only by `oracle/make_golden.py` (writes `tests/golden/*.npz`) and by
`tests/test_oracle_vs_reference.py` (skipped when `/root/reference` is absent, i.e. on the GPU box).
"""
import copy
import os
import sys
import warnings

REF_ROOT = os.environ.get("MMS_REFERENCE_ROOT", "/root/reference")
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "refshim")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "src", "field_components"))


def import_reference():
    """Put the shims and the reference on sys.path (idempotent) and import configs first."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found under {REF_ROOT}")
    src = os.path.join(REF_ROOT, "src")
    for p in (src, _SHIM):
        if p not in sys.path:
            sys.path.insert(0, p)
    warnings.filterwarnings("ignore")
    import configs.configs  # noqa: F401  (must be first: circular imports otherwise)
    from configs import method_configs  # noqa: F401
    return sys.modules["configs.configs"]


MODALITY_CHANNELS = {"rgb": 3, "infrared": 1, "mono": 1, "polarization": 4, "multispectral": 9}


def load_yaml(name: str) -> dict:
    import yaml
    with open(os.path.join(REF_ROOT, "confs", name)) as f:
        return yaml.safe_load(f)


def build_reference_model(preset="grid_raw", yaml_name="grid_raw.yaml", modalities=None,
                          log2_hashmap_size=None, num_samples=None, num_samples_importance=None,
                          bg_samples=None, seed=654824, direction_encoding="nerf"):
    """Returns (model, trainer_config). `modalities`: dict name->channels (default: the 5 of MMS)."""
    import torch
    cfgmod = import_reference()
    from configs.method_configs import method_configs
    from field_components.encodings import NeRFEncodingConfig
    from data.scene_box import SceneBox

    trainer_cfg = copy.deepcopy(method_configs[preset])
    y = load_yaml(yaml_name)
    y.pop("method", None)
    cfgmod.Config.update_config(trainer_cfg, y)
    m = trainer_cfg.pipeline.model
    # tcnn-free substitutions
    for f in (m.surface_model.surface_field.field, m.radiance_model.radiance_field.base_field):
        if not hasattr(f, "feature_grid"):          # presets `mlp*`: no hash grids
            continue
        fg = f.feature_grid
        fg.encoding.interpolation = "Linear"
        fg.encoding.implementation = "torch"
        if log2_hashmap_size is not None:
            fg.encoding.log2_hashmap_size = log2_hashmap_size
    bgf = m.background_model.background_field.base_field
    if hasattr(bgf, "feature_grid"):
        bgf.feature_grid.encoding.interpolation = "Linear"
        bgf.feature_grid.encoding.implementation = "torch"
        if log2_hashmap_size is not None:
            bgf.feature_grid.encoding.log2_hashmap_size = log2_hashmap_size
    if direction_encoding == "nerf":
        m.radiance_model.direction_encoding = NeRFEncodingConfig(num_frequencies=4, max_freq_exp=3)
    if num_samples is not None:
        m.ray_sampler.num_samples = num_samples
    if num_samples_importance is not None:
        m.ray_sampler.num_samples_importance = num_samples_importance
    if bg_samples is not None:
        m.background_ray_sampler.num_samples = bg_samples
    if modalities is None:
        modalities = dict(MODALITY_CHANNELS)
    torch.manual_seed(seed)
    scene_box = SceneBox(aabb=torch.tensor([[-1., -1, -1], [1, 1, 1]]), radius=1.0, collider_type="sphere")
    model = m.setup(scene_box=scene_box, modalities=modalities)
    return model, trainer_cfg


def set_schedule_state(model, level=16, delta=2.0 / 1024, anneal=1.0):
    """Hand-set what the BEFORE_TRAIN_ITERATION callbacks would set."""
    model.surface_model.volume_rendering.set_cos_anneal_ratio(anneal)
    model.surface_model.set_numerical_gradients_delta(delta)
    for f in (model.surface_model.surface_field.field, model.radiance_model.radiance_field.base_field):
        if hasattr(f, "feature_grid"):
            f.feature_grid.update_mask(level)
    bgf = model.background_model.background_field.base_field
    if hasattr(bgf, "feature_grid"):
        bgf.feature_grid.update_mask(level)
