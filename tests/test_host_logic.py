"""CPU: host-side mirror of the reference interface — construction, state-dict layout, schedules."""
import numpy as np
import pytest
import torch

from multimodalstudio_b200 import ops
from multimodalstudio_b200.configs import TrainingCallbackAttributes, TrainingCallbackLocation, update_config
from multimodalstudio_b200.models import MODALITY_CHANNELS, build_model, grid_loss_config


@pytest.fixture(scope="module")
def model():
    return build_model("grid_raw", log2_hashmap_size=10)


def test_state_dict_layout_matches_reference(model):
    sd = model.state_dict()
    assert len(sd) == 108
    t = "surface_model.surface_field.field.feature_grid.encoding.hash_table"
    assert sd[t].shape == (16 * 1024, 2)
    assert sd["radiance_model.radiance_field.base_field.feature_grid.encoding.hash_table"].shape == (16 * 1024, 2)
    assert sd["surface_model.surface_field.field.mlp_head.layers.0.parametrizations.weight.original1"].shape == (256, 71)
    assert sd["surface_model.surface_field.field.mlp_head.layers.2.parametrizations.weight.original1"].shape == (257, 256)
    assert sd["radiance_model.radiance_field.base_field.mlp_head.layers.0.parametrizations.weight.original1"].shape == (256, 319)
    assert sd["radiance_model.modality_heads.polarization.field.layers.2.bias"].shape == (3,)
    assert sd["background_model.background_field.head_field.layers.0.parametrizations.weight.original1"].shape == (256, 283)
    assert float(sd["surface_model.volume_rendering.density_fn.variance_network.s"]) == pytest.approx(0.3)
    assert float(sd[t].abs().max()) <= 1e-3


def test_resolutions_and_freqs():
    assert ops.hash_resolutions(16, 1024, 16) == [16, 21, 27, 36, 48, 64, 84, 111, 147, 194, 256, 337, 445, 588, 776, 1024]
    assert ops.nerf_freqs(0.0, 5, 6) == [1, 2, 4, 8, 16, 32]


def test_schedule_callbacks(model):
    class T:
        max_num_iterations = 1600
    cbs = model.get_training_callbacks(TrainingCallbackAttributes(model=model, trainer=T()))
    for step, level in ((0, 1), (100, 2), (1599, 16)):
        for cb in cbs:
            cb.run_callback_at_location(step, TrainingCallbackLocation.BEFORE_TRAIN_ITERATION)
        fg = model.surface_model.surface_field.field.feature_grid
        assert int(fg.hash_encoding_mask.sum()) == 2 * level
    assert model.surface_model.numerical_gradients_delta == pytest.approx(2.0 / 1024)
    assert model.surface_model.volume_rendering._cos_anneal_ratio == 1.0
    lm = grid_loss_config().setup(modalities=list(MODALITY_CHANNELS), num_iterations=1600, model=model)
    assert lm.curvature_loss._weight(80) == pytest.approx(5e-4 * 0.5)
    g = np.exp((np.log(1024) - np.log(16)) / 15)
    assert lm.curvature_loss._weight(1599) == pytest.approx(5e-4 / g ** 15)


def test_update_config_and_errors():
    from multimodalstudio_b200.field_components import HashEncodingConfig, MLPConfig
    class H:
        pass
    h = H(); h.enc = HashEncodingConfig()
    update_config(h, {"enc": {"max_res": 512, "num_levels": 8}})
    assert h.enc.max_res == 512 and h.enc.num_levels == 8
    with pytest.raises(ValueError):
        HashEncodingConfig(interpolation="Nearest").setup(in_dim=3)
    with pytest.raises(ValueError):
        HashEncodingConfig().setup(in_dim=0)
    with pytest.raises(ValueError):
        MLPConfig(activation="Tanh").setup(input_dim=4, output_dim=4)
    with pytest.raises(ValueError):
        build_model("no_such_preset")
    from multimodalstudio_b200.models import loss_config_for
    assert loss_config_for("grid_decimated").radiance_losses["rgb"].per_channel_probability == [0.25, 0.5, 0.25]
    mlp = build_model("mlp_raw", modalities={"rgb": 3, "mono": 1})
    assert not mlp.surface_model.config.use_numerical_gradients and not mlp.surface_model.config.compute_hessian
    assert [tuple(l.weight.shape) for l in mlp.surface_model.surface_field.field.layers][4] == (256, 295)   # skip at layer 4
    bg = build_model("grid_raw_grid_bg_unbalanced", modalities={"rgb": 3, "polarization": 4}, log2_hashmap_size=8)
    assert bg.background_model.background_field.base_field.feature_grid.radius == 2


def test_split_rows_falls_back_to_autograd_semantics_on_cpu():
    """ops.split_rows without sink-writing consumers (plain torch ops on CPU tensors): the backward is autograd's zero-fill +
    concatenation, block by block — same gradient as torch.split, also for an unused and an empty block."""
    import torch
    from multimodalstudio_b200 import ops
    torch.manual_seed(0)
    x = torch.randn(10, 4, requires_grad=True)
    sizes = [3, 0, 5, 2]

    def loss_of(blocks):
        return (blocks[0] ** 2).sum() + (blocks[2] * 3.0).sum()          # block 3 unused, block 1 empty

    g1, = torch.autograd.grad(loss_of(ops.split_rows(x * 1.0, sizes)), x)
    g2, = torch.autograd.grad(loss_of(torch.split(x * 1.0, sizes, dim=0)), x)
    assert torch.equal(g1, g2)
    # no autograd: plain views
    with torch.no_grad():
        blocks = ops.split_rows(x, sizes)
    assert [b.shape[0] for b in blocks] == sizes and not hasattr(blocks[0], "_mmsb_sink")
