"""Host-side mirror of the reference's `field_components` package (same class / config names, argument
meaning, state-dict keys and error behaviour); the arithmetic runs in libmms_b200.so through `ops`.

ref: src/field_components/{base_field_component,encodings,mlp,feature_structures,field_heads,
     spatial_distortions,single_variance}.py and src/model_components/polarizer.py
"""
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple, Type, Union

import numpy as np
import torch
from torch import nn

from . import ops
from .configs import (InstantiateConfig, TrainingCallback, TrainingCallbackAttributes,
                      TrainingCallbackLocation)


# ---------------------------------------------------------------------------------------------
# base  (ref: base_field_component.py:33-94)
# ---------------------------------------------------------------------------------------------
@dataclass
class FieldComponentConfig(InstantiateConfig):
    _target: Type = field(default_factory=lambda: FieldComponent)
    input_dim: int = None
    output_dim: int = None


class FieldComponent(nn.Module):
    def __init__(self, config: FieldComponentConfig, input_dim: Optional[int] = None, output_dim: Optional[int] = None):
        super().__init__()
        self.config = config
        self.input_dim = input_dim if input_dim is not None else self.config.input_dim
        self.output_dim = output_dim if output_dim is not None else self.config.output_dim

    def set_in_dim(self, input_dim: int) -> None:
        if input_dim <= 0:
            raise ValueError("Input dimension should be greater than zero")
        self.input_dim = input_dim

    def get_out_dim(self) -> int:
        if self.output_dim is None:
            raise ValueError("Output dimension has not been set")
        return self.output_dim

    def forward(self, input_tensor):
        raise NotImplementedError

    def get_training_callbacks(self, training_callback_attributes: TrainingCallbackAttributes) -> List[TrainingCallback]:
        return []


# ---------------------------------------------------------------------------------------------
# encodings  (ref: encodings.py)
# ---------------------------------------------------------------------------------------------
@dataclass
class EncodingConfig(FieldComponentConfig):
    _target: Type = field(default_factory=lambda: Encoding)


@dataclass
class HashEncodingConfig(EncodingConfig):
    """ref: encodings.py:47-67.  `implementation`: the reference selects "tcnn" / "torch"; here every
    value runs the sm_100a kernels ("b200"), there is no other backend."""
    _target: Type = field(default_factory=lambda: HashEncoding)
    num_levels: int = 16
    features_per_level: int = 2
    min_res: int = 16
    max_res: int = 2048
    log2_hashmap_size: int = 19
    hash_init_scale: float = 0.001
    interpolation: Optional[str] = "Smoothstep"
    implementation: str = "b200"


@dataclass
class NeRFEncodingConfig(EncodingConfig):
    _target: Type = field(default_factory=lambda: NeRFEncoding)
    num_frequencies: int = 6
    min_freq_exp: float = 0.0
    max_freq_exp: int = 5
    include_input: bool = True


@dataclass
class SHEncodingConfig(EncodingConfig):
    _target: Type = field(default_factory=lambda: SHEncoding)
    degree: int = 4


class Encoding(FieldComponent):
    def __init__(self, config: EncodingConfig, in_dim: int = 3) -> None:
        if in_dim <= 0:
            raise ValueError("Input dimension should be greater than zero")
        super().__init__(config, input_dim=in_dim)
        self.config = config


class NeRFEncoding(Encoding):
    """ref: encodings.py:131-182"""

    def __init__(self, config: NeRFEncodingConfig, in_dim: int = 3) -> None:
        super().__init__(config, in_dim=in_dim)
        self.num_frequencies = self.config.num_frequencies
        self.min_freq = self.config.min_freq_exp
        self.max_freq = self.config.max_freq_exp
        self.include_input = self.config.include_input
        self.freqs = ops.nerf_freqs(self.min_freq, self.max_freq, self.num_frequencies)

    def get_out_dim(self) -> int:
        if self.input_dim is None:
            raise ValueError("Input dimension has not been set")
        out_dim = self.input_dim * self.num_frequencies * 2
        if self.include_input:
            out_dim += self.input_dim
        return out_dim

    def forward(self, input_tensor):
        return ops.NerfEncodingFn.apply(input_tensor, self.freqs, self.include_input)

    def piece(self, input_tensor):
        """This encoding as a column range of an `ops.assemble` row (written in place by the kernel)."""
        return ops.nerf_piece(input_tensor.reshape(-1, input_tensor.shape[-1]), self.freqs, self.include_input)


class HashEncoding(Encoding):
    """ref: encodings.py:184-310 (torch path semantics; table layout [L * 2^log2, F] fp32)."""

    def __init__(self, config: HashEncodingConfig, in_dim: int = 3) -> None:
        super().__init__(config, in_dim=in_dim)
        self.growth_factor = np.exp(
            (np.log(self.config.max_res) - np.log(self.config.min_res)) / (self.config.num_levels - 1))
        self.implementation = "b200"
        self.hash_table_size = 2 ** self.config.log2_hashmap_size
        levels = torch.arange(self.config.num_levels)
        self.scalings = torch.floor(self.config.min_res * self.growth_factor ** levels)
        self.hash_offset = levels * self.hash_table_size
        table = torch.rand(size=(self.hash_table_size * self.config.num_levels, self.config.features_per_level)) * 2 - 1
        table *= self.config.hash_init_scale
        self.hash_table = nn.Parameter(table)
        if self.config.interpolation not in (None, "Linear", "Smoothstep"):
            raise ValueError(f"interpolation '{self.config.interpolation}' is not supported")
        self._descs: Dict[float, ops.MmsbHashGridDesc] = {}

    def get_out_dim(self) -> int:
        return self.config.num_levels * self.config.features_per_level

    def desc(self, radius: float = 0.0):
        if radius not in self._descs:
            self._descs[radius] = ops.make_hashgrid_desc(
                self.config.num_levels, self.config.features_per_level, self.config.log2_hashmap_size,
                self.scalings.tolist(), radius=radius, interpolation=self.config.interpolation or "Linear")
        return self._descs[radius]

    def forward(self, input_tensor, radius: float = 0.0, mask: Optional[torch.Tensor] = None):
        assert input_tensor.shape[-1] == 3
        return ops.HashGridFn.apply(input_tensor, self.hash_table, mask, self.desc(radius))

    def hash_indices(self, input_tensor):
        """int64 [..., L, 8] flat rows of the 8 corners (hashed_0..7 of encodings.py:274-281)."""
        idx, _ = ops.hashgrid_indices(self.desc(0.0), input_tensor, self.hash_table)
        return idx.reshape(*input_tensor.shape[:-1], self.config.num_levels, 8)


class SHEncoding(Encoding):
    """ref: encodings.py:368-392 — `degree + 1` levels; the basis is the in-tree definition
    utils/math.py:21-82 (tcnn's op is not in the tree)."""

    def __init__(self, config: SHEncodingConfig, in_dim: int = 3):
        super().__init__(config, in_dim=in_dim)

    def get_out_dim(self) -> int:
        return (self.config.degree + 1) ** 2

    def forward(self, input_tensor):
        return ops.SHEncodingFn.apply(input_tensor, self.config.degree + 1)

    def piece(self, input_tensor):
        return ops.copy_piece(self.forward(input_tensor.reshape(-1, 3)))


# ---------------------------------------------------------------------------------------------
# MLP  (ref: mlp.py:32-209)
# ---------------------------------------------------------------------------------------------
@dataclass
class MLPConfig(FieldComponentConfig):
    _target: Type = field(default_factory=lambda: MLP)
    num_layers: int = 8
    hidden_dim: int = 128
    weight_norm: bool = True
    activation: str = "ReLU"
    activation_params: dict = field(default_factory=dict)
    out_activation: Optional[str] = "Sigmoid"
    skip_connections: Optional[Tuple[int]] = field(default_factory=lambda: [])
    geometric_init: bool = False
    geometric_init_bias: float = 0.5


class MLP(FieldComponent):
    """Same parameters as the reference (`layers.N` = weight-norm-parametrized nn.Linear, so state
    dicts interchange); forward / backward run the fused-layer kernels."""

    def __init__(self, config: MLPConfig, input_dim: int = None, output_dim: int = None):
        self.config = config
        super().__init__(config, input_dim=input_dim, output_dim=output_dim)
        if self.output_dim is None:
            self.output_dim = self.config.hidden_dim
        if self.config.activation not in ("ReLU", "Softplus"):
            raise ValueError(f"activation '{self.config.activation}' not supported")
        if self.config.out_activation not in ("None", None, "ReLU", "Sigmoid", "Softplus"):
            raise ValueError(f"out_activation '{self.config.out_activation}' not supported")

        dims = []
        for i in range(self.config.num_layers - 1):
            if i + 1 in self.config.skip_connections:
                dims.append(self.config.hidden_dim + self.input_dim)
            else:
                dims.append(self.config.hidden_dim)
        dims = [self.input_dim] + dims + [self.output_dim]
        layers = []
        for i in range(0, len(dims) - 1):
            out_dim = dims[i + 1] - dims[0] if i + 1 in self.config.skip_connections else dims[i + 1]
            layers.append(nn.Linear(dims[i], out_dim))
        self.layers = nn.ModuleList(layers)

        if self.config.geometric_init:
            self.geometric_init(bias=self.config.geometric_init_bias, additional_input=self.input_dim > 3)
        else:
            self.standard_init()
        if self.config.weight_norm:
            self.weight_norm()
        self.act_param = float(self.config.activation_params.get("beta", 1.0)) if self.config.activation == "Softplus" else 1.0

    def forward(self, input_tensor, n_out_used: Optional[int] = None, first_weight=None):
        """`first_weight`: replaces layers[0].weight (a column-permuted view of it for reordered input rows)."""
        out_act = self.config.out_activation
        if out_act == "Softplus" and self.config.activation != "Softplus":
            # nn.Softplus() default beta for the output (density head) while hidden is ReLU: one act_param suffices
            act_param = 1.0
        else:
            act_param = self.act_param
        weights = [layer.weight for layer in self.layers]
        if first_weight is not None:
            weights[0] = first_weight
        biases = [layer.bias for layer in self.layers]
        return ops.mlp_forward(input_tensor, weights, biases, self.config.activation, out_act, act_param,
                               tuple(self.config.skip_connections), n_out_used)

    def geometric_init(self, bias=0.5, inside_outside=False, additional_input=True):
        """ref: mlp.py:173-198"""
        nl = len(self.layers)
        for l in range(nl):
            in_dim, out_dim = self.layers[l].in_features, self.layers[l].out_features
            if l == nl - 1:
                sign = -1.0 if inside_outside else 1.0
                torch.nn.init.normal_(self.layers[l].weight, mean=sign * np.sqrt(np.pi) / np.sqrt(in_dim), std=0.0001)
                torch.nn.init.constant_(self.layers[l].bias, -sign * bias)
            elif additional_input and l == 0:
                torch.nn.init.constant_(self.layers[l].bias, 0.0)
                torch.nn.init.constant_(self.layers[l].weight[:, 3:], 0.0)
                torch.nn.init.normal_(self.layers[l].weight[:, :3], 0.0, np.sqrt(2) / np.sqrt(out_dim))
            elif additional_input and l in self.config.skip_connections:
                torch.nn.init.constant_(self.layers[l].bias, 0.0)
                torch.nn.init.normal_(self.layers[l].weight, 0.0, np.sqrt(2) / np.sqrt(out_dim))
                torch.nn.init.constant_(self.layers[l].weight[:, -(self.layers[0].in_features - 3):], 0.0)
            else:
                torch.nn.init.constant_(self.layers[l].bias, 0.0)
                torch.nn.init.normal_(self.layers[l].weight, 0.0, np.sqrt(2) / np.sqrt(out_dim))

    def standard_init(self):
        for layer in self.layers:
            torch.nn.init.kaiming_uniform_(layer.weight.data)
            torch.nn.init.zeros_(layer.bias.data)

    def weight_norm(self):
        for l in range(len(self.layers)):
            self.layers[l] = torch.nn.utils.parametrizations.weight_norm(self.layers[l])


# ---------------------------------------------------------------------------------------------
# feature structures  (ref: feature_structures.py)
# ---------------------------------------------------------------------------------------------
@dataclass
class FeatureGridConfig(FieldComponentConfig):
    _target: Type = field(default_factory=lambda: FeatureGrid)
    encoding: EncodingConfig = field(default_factory=lambda: EncodingConfig)
    coarse_to_fine: bool = True
    steps_per_level_ratio: float = 1.0
    level_init: int = 1
    radius: float = 1


@dataclass
class FeatureGridAndMLPConfig(FieldComponentConfig):
    _target: Type = field(default_factory=lambda: FeatureGridAndMLP)
    feature_grid: FeatureGridConfig = field(default_factory=lambda: FeatureGridConfig)
    mlp_head: MLPConfig = field(default_factory=lambda: MLPConfig)
    return_features: bool = False


class FeatureGrid(FieldComponent):
    """ref: feature_structures.py:56-128.  Rescale + encode + level mask are one kernel."""

    def __init__(self, config: FeatureGridConfig, input_dim: int = None, output_dim: int = None):
        super().__init__(config, input_dim=input_dim, output_dim=output_dim)
        self.config = config
        self.radius = self.config.radius
        self.encoding = self.config.encoding.setup(in_dim=3)
        self.output_dim = self.encoding.get_out_dim()
        self.hash_encoding_mask = torch.ones(
            self.config.encoding.num_levels * self.config.encoding.features_per_level, dtype=torch.float32)
        self.active_level = self.config.encoding.num_levels

    def _mask_on(self, device):
        if self.hash_encoding_mask.device != device:
            self.hash_encoding_mask = self.hash_encoding_mask.to(device)
        return self.hash_encoding_mask

    def forward(self, input_tensor):
        return self.encoding(input_tensor, radius=float(self.radius), mask=self._mask_on(input_tensor.device))

    def piece(self, positions):
        return ops.hash_piece(positions.reshape(-1, 3), self.encoding.hash_table, self._mask_on(positions.device),
                              self.encoding.desc(float(self.radius)))

    def update_mask(self, level: int):
        self.active_level = int(level)
        self.hash_encoding_mask[:] = 1.0
        self.hash_encoding_mask[level * self.config.encoding.features_per_level:] = 0

    def get_training_callbacks(self, training_callback_attributes):
        callbacks = super().get_training_callbacks(training_callback_attributes)
        if self.config.coarse_to_fine:
            def set_mask(step):
                max_it = training_callback_attributes.trainer.max_num_iterations
                steps_per_level = int(max_it * self.config.steps_per_level_ratio)
                steps_per_level = min(steps_per_level, int(max_it / self.config.encoding.num_levels))
                level = int(step / steps_per_level) + 1
                level = max(level, self.config.level_init)
                level = min(level, self.config.encoding.num_levels)
                self.update_mask(level)

            callbacks.append(TrainingCallback(where_to_run=[TrainingCallbackLocation.BEFORE_TRAIN_ITERATION],
                                              update_every_num_iters=1, func=set_mask))
        return callbacks

    def get_model_parameters(self):
        return {
            "num_levels": self.config.encoding.num_levels,
            "min_res": self.config.encoding.min_res,
            "max_res": self.config.encoding.max_res,
            "steps_per_level_ratio": self.config.steps_per_level_ratio,
            "level_init": self.config.level_init,
        }


class FeatureGridAndMLP(FieldComponent):
    """ref: feature_structures.py:130-173"""

    def __init__(self, config: FeatureGridAndMLPConfig, input_dim: int = None, output_dim: int = None):
        super().__init__(config, input_dim=input_dim, output_dim=output_dim)
        self.config = config
        self.feature_grid = self.config.feature_grid.setup(input_dim=3)
        mlp_input_dim = input_dim + self.feature_grid.encoding.get_out_dim()
        self.mlp_head = self.config.mlp_head.setup(input_dim=mlp_input_dim, output_dim=output_dim)
        self.output_dim = self.mlp_head.get_out_dim()

    def forward(self, input_tensor=None, n_out_used: Optional[int] = None, pieces=None, positions=None):
        """`pieces` + `positions`: the row cat[pieces..., hash features(positions)] is assembled in place by the
        encoder kernels (ops.assemble) instead of torch.cat over materialised parts — same values."""
        if pieces is not None and not self.config.return_features:
            mlp_input, perm = self.assemble_input(pieces, positions)
            w0 = ops.permuted_columns(self.mlp_head.layers[0].weight, perm)
            return self.mlp_head(mlp_input, n_out_used=n_out_used, first_weight=w0)
        if input_tensor is None:
            input_tensor = ops.assemble(list(pieces))
        features = self.feature_grid(input_tensor[..., :3])
        mlp_input = torch.cat([input_tensor, features], dim=-1)   # == cat[x, aux, features]
        output = self.mlp_head(mlp_input, n_out_used=n_out_used)
        if self.config.return_features:
            return output, features
        return output

    def assemble_input(self, pieces, positions):
        """The MLP input row with the hash features FIRST and the other pieces by decreasing width (the hash-grid and
        copy kernels then write 16/32-byte aligned), plus the column permutation (new position -> reference column of
        cat[pieces..., features], feature_structures.py:164) to apply to the first layer's weight."""
        all_pieces = list(pieces) + [self.feature_grid.piece(positions)]
        widths = [ops._piece_width(p) for p in all_pieces]
        starts = [sum(widths[:i]) for i in range(len(widths))]
        last = len(all_pieces) - 1
        order = [last] + sorted(range(last), key=lambda i: -widths[i])
        perm = tuple(c for i in order for c in range(starts[i], starts[i] + widths[i]))
        return ops.assemble([all_pieces[i] for i in order]), perm

    def get_training_callbacks(self, training_callback_attributes):
        return self.feature_grid.get_training_callbacks(training_callback_attributes)

    def get_model_parameters(self):
        return self.feature_grid.get_model_parameters()


# ---------------------------------------------------------------------------------------------
# heads  (ref: field_heads.py, polarizer.py)
# ---------------------------------------------------------------------------------------------
@dataclass
class ModalityHeadConfig(FieldComponentConfig):
    _target: Type = field(default_factory=lambda: ModalityHead)
    field: Optional[FieldComponentConfig] = field(
        default_factory=lambda: MLPConfig(num_layers=1, hidden_dim=64, weight_norm=True, out_activation="Sigmoid"))


@dataclass
class PolarizationHeadConfig(ModalityHeadConfig):
    _target: Type = field(default_factory=lambda: PolarizationHead)
    field: Optional[FieldComponentConfig] = field(
        default_factory=lambda: MLPConfig(num_layers=1, hidden_dim=64, weight_norm=True, out_activation="None"))


class ModalityHead(FieldComponent):
    def __init__(self, config: ModalityHeadConfig, input_dim: int = None, output_dim: int = None):
        super().__init__(config, input_dim=input_dim, output_dim=output_dim)
        self.config = config
        assert input_dim is not None, "input_dim must be provided"
        assert output_dim is not None, "output_dim must be provided"
        self.field = self.config.field.setup(input_dim=input_dim, output_dim=output_dim)

    def forward(self, input_tensor, **kwargs):
        return self.field(input_tensor)


def mueller_rotate(theta):
    """ref: polarizer.py:39-52"""
    c, s = torch.cos(2 * theta), torch.sin(2 * theta)
    one, zero = torch.ones_like(c), torch.zeros_like(c)
    return torch.stack([one, zero, zero, zero, c, s, zero, -s, c], dim=-1).view(-1, 3, 3)


def align_polarization_filters(stokes_vectors, directions, camera_up_directions):
    """ref: polarizer.py:54-82"""
    z = ops.const_tensor("z_axis", lambda: torch.tensor([0.0, 0.0, 1.0]), directions.device)[None].expand(directions.shape)
    normal = torch.nn.functional.normalize(torch.linalg.cross(directions, z), dim=-1)
    cos_theta = torch.clamp(torch.sum(normal * camera_up_directions, dim=-1), min=-1 + 1e-4, max=1 - 1e-4)
    theta = torch.acos(cos_theta) - np.pi / 2
    return (mueller_rotate(theta) @ stokes_vectors[..., None]).squeeze(-1)


def stokes_to_intensity(stokes_vectors):
    """ref: polarizer.py:84-101"""
    m = ops.const_tensor("stokes2int", lambda: 0.5 * torch.tensor([[1.0, 1.0, 0.0], [1.0, 0.0, 1.0], [1.0, -1.0, 0.0],
                                                                     [1.0, 0.0, -1.0]]), stokes_vectors.device)
    return (m[None] @ stokes_vectors[..., None]).squeeze(-1)


class PolarizationHead(ModalityHead):
    """ref: field_heads.py:75-106"""

    def __init__(self, config: PolarizationHeadConfig, input_dim: int = None, output_dim: int = 3):
        super().__init__(config, input_dim=input_dim, output_dim=output_dim)
        self.config = config
        self.field = self.config.field.setup(input_dim=input_dim, output_dim=3)

    def forward(self, input_tensor, directions, up_directions):
        stokes = self.field(input_tensor)
        # leaky_relu(S0) + align_polarization_filters + stokes_to_intensity: one kernel (mmsb_polarization_fwd/bwd)
        return ops.PolarizationFn.apply(stokes, directions, up_directions)


# ---------------------------------------------------------------------------------------------
# spatial distortion / variance  (ref: spatial_distortions.py:65-97, single_variance.py:19-36)
# ---------------------------------------------------------------------------------------------
@dataclass
class SpatialDistortionConfig(InstantiateConfig):
    _target: Type = field(default_factory=lambda: SpatialDistortion)


@dataclass
class SceneContractionConfig(SpatialDistortionConfig):
    _target: Type = field(default_factory=lambda: SceneContraction)
    order: Union[None, int, float] = None


class SpatialDistortion(nn.Module):
    def forward(self, positions):
        raise NotImplementedError


class SceneContraction(SpatialDistortion):
    def __init__(self, config) -> None:
        super().__init__()
        self.order = config.order

    def forward(self, positions):
        mag = torch.linalg.norm(positions, ord=self.order, dim=-1, keepdim=True)
        contracted = (2 - (1 / mag)) * (positions / mag)
        return torch.where(mag >= 1, contracted, positions)   # sync-free form of the masked assignment


class SingleVarianceNetwork(nn.Module):
    def __init__(self, init_val):
        super().__init__()
        self.register_parameter("s", nn.Parameter(init_val * torch.ones(1), requires_grad=True))

    def forward(self, x):
        return torch.ones([len(x), 1], device=x.device) * torch.exp(self.s * 10.0)

    def get_inv_variance(self):
        return torch.exp(self.s * 10.0).clip(1e-6, 1e6)
