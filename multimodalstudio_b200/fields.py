"""Mirror of the reference's `fields` package.
ref: src/fields/surface_field.py:27-116, src/fields/radiance_field.py:25-81, src/fields/nerf_field.py:35-105
"""
from dataclasses import dataclass, field
from typing import Optional, Type

import torch

from . import ops
from .configs import InstantiateConfig
from .field_components import (EncodingConfig, FieldComponent, FieldComponentConfig, MLPConfig, ModalityHeadConfig,
                               NeRFEncodingConfig)


@dataclass
class SurfaceFieldConfig(InstantiateConfig):
    _target: Type = field(default_factory=lambda: SurfaceField)
    use_position_encoding: bool = True
    position_encoding: EncodingConfig = field(default_factory=lambda: NeRFEncodingConfig)
    geo_feature_dim: int = 256
    field: FieldComponentConfig = field(default_factory=lambda: MLPConfig)


@dataclass
class SDFFieldConfig(SurfaceFieldConfig):
    _target: Type = field(default_factory=lambda: SDFField)
    inside_outside: bool = False


class SurfaceField(torch.nn.Module):
    def __init__(self, config: SurfaceFieldConfig):
        super().__init__()
        self.config = config
        self.position_encoding = self.config.position_encoding.setup(in_dim=3)
        self.input_dim = self.position_encoding.get_out_dim() if self.config.use_position_encoding else 3
        self.output_dim = 1 + self.config.geo_feature_dim if self.config.geo_feature_dim is not None else 1

    def single_output(self, x):
        return self.forward(x, sdf_only=True)[0]

    def get_training_callbacks(self, training_callback_attributes):
        return self.field.get_training_callbacks(training_callback_attributes)

    def get_model_parameters(self):
        return self.field.get_model_parameters()


class SDFField(SurfaceField):
    """ref: surface_field.py:86-116.  `sdf_only=True` evaluates just the first output of the last
    layer (what `single_output` keeps) instead of computing and discarding the 256 geometry features."""

    def __init__(self, config: SDFFieldConfig):
        super().__init__(config)
        self.field = self.config.field.setup(input_dim=self.input_dim, output_dim=self.output_dim)

    def _fused(self):
        """The fused SDF network (ops.SdfNetFn) applies to the shipped shape: PE + hash grid, two hidden layers, no skips,
        no output activation, tcgen05 layer path."""
        mlp = getattr(self.field, "mlp_head", None)
        return (ops.MLP_PRECISION != 0 and self.config.use_position_encoding and hasattr(self.field, "feature_grid")
                and mlp is not None and len(mlp.layers) == 3 and not mlp.config.skip_connections
                and mlp.config.out_activation in (None, "None") and self.config.geo_feature_dim is not None
                and mlp.layers[1].out_features <= 256)

    def forward_split(self, x, n_full: int, group: int = 1):
        """x [n, 3] -> sdf [n, 1] for every row, geo_feature [n_full, G] for the full rows: the first n_full ones
        (`group` = 1) or row 0 of every group of `group` rows (one network call for the centre evaluations and the
        finite-difference taps of surface_model.py:129-152; see ops.SdfNetFn for why the grouped layout matters)."""
        mlp = self.field.mlp_head
        rows, perm = self.field.assemble_input([self.position_encoding.piece(x)], x)
        weights = [l.weight for l in mlp.layers]
        weights[0] = ops.permuted_columns(weights[0], perm)
        return ops.sdf_net_forward(rows, n_full, weights, [l.bias for l in mlp.layers], mlp.config.activation, mlp.act_param,
                                   group=group)

    def forward_with_gradient(self, x):
        """x [n, 3] -> sdf [n, 1], geo_feature [n, G], d sdf / d x [n, 3] for an MLP field (presets `mlp*`).

        The reference takes the gradient with `torch.autograd.grad(sdf, x, create_graph=True)` (surface_model.py:193-203)
        and differentiates THROUGH it in the training backward (a double backward).  Here the same quantity is a
        forward-mode pass: three tangent rows per point (d PE / d x_k) go through the same linear layers (no bias) and are
        scaled by the activation derivative of every hidden layer, t <- act'(z) * (W t).  It is built from the ordinary
        differentiable operators (the tensor-core layer kernels + element-wise torch ops), so the training backward is a
        plain first-order backward of a forward computation: no double backward exists on this path.
        act'(z) is evaluated from the stored activation: Softplus_beta' = 1 - exp(-beta * softplus(z)), ReLU' = [y > 0]."""
        import math
        mlp = self.field
        if hasattr(mlp, "feature_grid") or not self.config.use_position_encoding:
            raise NotImplementedError("analytic SDF gradients are implemented for the PE + MLP field of the `mlp*` presets")
        enc = self.position_encoding
        n = x.shape[0]
        pe = enc(x)                                                        # [n, P]
        # tangents of the encoding: d/dx_k of [x, sin(f x_d), sin(f x_d + pi/2)] (encodings.py:161-182)
        fr = ops.const_tensor(("nerf_freqs", tuple(enc.freqs)), lambda: torch.tensor(enc.freqs, dtype=torch.float32), x.device)  # [K]
        k_ = fr.shape[0]
        s = x[:, :, None] * fr                                             # [n, 3, K]
        eye = ops.const_tensor("eye3", lambda: torch.eye(3, dtype=torch.float32), x.device)
        d_sin = (fr * torch.cos(s))[:, None, :, :] * eye[None, :, :, None]               # [n, k, d, K]
        d_cos = (fr * torch.cos(s + math.pi / 2.0))[:, None, :, :] * eye[None, :, :, None]
        parts = ([eye[None].expand(n, 3, 3)] if enc.include_input else []) + [d_sin.reshape(n, 3, 3 * k_), d_cos.reshape(n, 3, 3 * k_)]
        t_pe = torch.cat(parts, dim=-1).reshape(3 * n, -1)                  # [3 n, P], row 3 i + k = d PE(x_i) / d x_k
        act, beta = mlp.config.activation, mlp.act_param
        skips = tuple(mlp.config.skip_connections)
        nl = len(mlp.layers)
        h, t = pe, t_pe
        for i, layer in enumerate(mlp.layers):
            if i in skips:
                h = torch.cat([h, pe], -1) / math.sqrt(2)
                t = torch.cat([t, t_pe], -1) / math.sqrt(2)
            last = i == nl - 1
            h = ops.mlp_forward(h, [layer.weight], [layer.bias], act, mlp.config.out_activation if last else act, beta)
            t = ops.mlp_forward(t, [layer.weight], [None], act, "None", beta, n_out_used=1 if last else None)
            if not last:
                d_act = (1.0 - torch.exp(-beta * h)) if act == "Softplus" else (h > 0).to(h.dtype)
                t = (d_act[:, None, :] * t.view(n, 3, -1)).reshape(3 * n, -1)
        if mlp.config.out_activation not in (None, "None"):
            raise NotImplementedError("analytic SDF gradients need a linear output layer")
        sdf, geo = h[:, :1], h[:, 1:]
        return sdf, geo, t.view(n, 3)

    def forward(self, x, sdf_only: bool = False):
        if self._fused():
            x2 = x.reshape(-1, 3)
            sdf, geo = self.forward_split(x2, 0 if sdf_only else x2.shape[0])
            return sdf.reshape(*x.shape[:-1], 1), (None if sdf_only else geo.reshape(*x.shape[:-1], geo.shape[-1]))
        if self.config.use_position_encoding and hasattr(self.field, "feature_grid"):
            # cat[PE(x), hash(x)] written in place by the two encoders
            kw = dict(pieces=[self.position_encoding.piece(x)], positions=x)
            if sdf_only:
                return self.field(n_out_used=1, **kw), None
            out = self.field(**kw)
        else:
            if self.config.use_position_encoding:
                x = self.position_encoding(x)
            if sdf_only:
                return self.field(x, n_out_used=1), None
            out = self.field(x)
        if self.config.geo_feature_dim is not None:
            sdf, geo_feature = torch.split(out, [1, self.config.geo_feature_dim], dim=-1)
        else:
            sdf, geo_feature = out, None
        return sdf, geo_feature


@dataclass
class BaseRadianceFieldConfig(FieldComponentConfig):
    _target: Type = field(default_factory=lambda: RadianceField)


@dataclass
class RadianceFieldConfig(BaseRadianceFieldConfig):
    _target: Type = field(default_factory=lambda: RadianceField)
    base_field: FieldComponentConfig = field(default_factory=lambda: MLPConfig)


class RadianceField(FieldComponent):
    """ref: radiance_field.py:55-81"""

    def __init__(self, config: RadianceFieldConfig, position_dim=3, view_direction_dim=3, additional_input_dim=0,
                 output_dim: int = 3):
        input_dim = position_dim + view_direction_dim + additional_input_dim
        super().__init__(config, input_dim=input_dim, output_dim=output_dim)
        self.base_field = self.config.base_field.setup(input_dim=self.input_dim, output_dim=self.output_dim)

    def forward(self, positions, view_directions, additional_inputs, view_direction_piece=None):
        if hasattr(self.base_field, "feature_grid"):
            pieces = [ops.copy_piece(positions),
                      view_direction_piece if view_direction_piece is not None else ops.copy_piece(view_directions)]
            pieces += [ops.copy_piece(a) for a in (additional_inputs if isinstance(additional_inputs, (list, tuple))
                                                   else [additional_inputs])]
            return self.base_field(pieces=pieces, positions=positions)
        if isinstance(additional_inputs, (list, tuple)):
            additional_inputs = torch.cat(list(additional_inputs), dim=-1)
        if view_directions is None:
            view_directions = ops.assemble([view_direction_piece])
        inputs = torch.cat([positions, view_directions, additional_inputs], dim=-1)
        return self.base_field(inputs)

    def get_training_callbacks(self, training_callback_attributes):
        return self.base_field.get_training_callbacks(training_callback_attributes)

    def get_model_parameters(self):
        return self.base_field.get_model_parameters()


@dataclass
class NeRFFieldConfig(FieldComponentConfig):
    _target: Type = field(default_factory=lambda: NeRFField)
    base_field: FieldComponentConfig = field(default_factory=lambda: MLPConfig)
    head_field: FieldComponentConfig = field(default_factory=lambda: MLPConfig)
    use_position_encoding: bool = True
    position_encoding: EncodingConfig = field(default_factory=lambda: NeRFEncodingConfig)
    use_direction_encoding: bool = True
    direction_encoding: EncodingConfig = field(default_factory=lambda: NeRFEncodingConfig)


class NeRFField(torch.nn.Module):
    """ref: nerf_field.py:53-105"""

    def __init__(self, config: NeRFFieldConfig, radiance_output_dim: int = 3):
        super().__init__()
        self.config = config
        self.position_encoding = self.config.position_encoding.setup(in_dim=3)
        self.direction_encoding = self.config.direction_encoding.setup(in_dim=3)
        base_input = self.position_encoding.get_out_dim() if self.config.use_position_encoding else 3
        head_input = self.config.base_field.output_dim + self.direction_encoding.get_out_dim() \
            if self.config.use_direction_encoding else 3 + self.config.base_field.output_dim
        self.base_field = self.config.base_field.setup(input_dim=base_input, output_dim=self.config.base_field.output_dim)
        self.head_field = self.config.head_field.setup(input_dim=head_input, output_dim=radiance_output_dim)
        self.density_head = ModalityHeadConfig(
            field=MLPConfig(num_layers=1, hidden_dim=64, weight_norm=True, out_activation="Softplus")
        ).setup(input_dim=self.base_field.output_dim, output_dim=1)

    def forward(self, x, viewing_direction):
        piece = self.position_encoding.piece(x) if self.config.use_position_encoding else ops.copy_piece(x)
        if hasattr(self.base_field, "feature_grid"):
            # hash-grid background (preset grid_raw_grid_bg_unbalanced): cat[x, PE(x)[3:], hash(x)] assembled in place
            feature = self.base_field(pieces=[piece], positions=x)
        else:
            feature = self.base_field(ops.assemble([piece]))
        density = self.density_head(feature)
        head_input = ops.assemble([ops.copy_piece(feature),
                                   self.direction_encoding.piece(viewing_direction) if self.config.use_direction_encoding
                                   else ops.copy_piece(viewing_direction)])
        feature = self.head_field(head_input)
        return density, feature
