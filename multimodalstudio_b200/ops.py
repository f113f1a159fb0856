"""Autograd wrappers over the C ABI (include/mms_b200.h): one `torch.autograd.Function` per operator
of the hot path.  PyTorch only supplies device memory, the current stream and the autograd tape;
all arithmetic runs in libmms_b200.so.  Non-CUDA tensors raise (no fallback).
"""
import ctypes
import math
import os
import weakref
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import MmsbHashGridDesc, call, ptr, stream_ptr

_CONST_CACHE = {}


def const_tensor(key, make, device):
    """Small device constants (tap offsets, Mueller matrices, linspaces): built once per device — a host-to-device
    copy inside the step would break CUDA-graph capture."""
    k = (key, str(device))
    t = _CONST_CACHE.get(k)
    if t is None:
        t = make().to(device)
        _CONST_CACHE[k] = t
    return t


ACT = {"None": 0, None: 0, "ReLU": 1, "Softplus": 2, "Sigmoid": 3}
INTERP = {"Linear": 0, None: 0, "Smoothstep": 1}
SPACING_UNIFORM, SPACING_DISPARITY = 0, 1

_i32, _i64, _f32 = ctypes.c_int32, ctypes.c_int64, ctypes.c_float


def _f(t: torch.Tensor) -> torch.Tensor:
    """fp32 contiguous CUDA view/copy of t."""
    if not t.is_cuda:
        raise ValueError("mms_b200 ops need CUDA tensors (no CPU fallback)")
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _rows(t: torch.Tensor, last: int) -> torch.Tensor:
    """[..., last] -> [n, last] with unit column stride (row-strided 2-D inputs, e.g. the padded rows of
    `assemble`, pass through without a copy: every kernel takes its leading dimension)."""
    if t.shape[-1] != last:
        raise ValueError(f"expected last dimension {last}, got {tuple(t.shape)}")
    if t.dim() == 2 and t.is_cuda and t.dtype == torch.float32 and t.stride(1) == 1 and t.stride(0) >= last:
        return t
    return _f(t).reshape(-1, last)


def _zeros_many(shapes, device):
    """Zero-initialised fp32 tensors of the given shapes carved out of ONE buffer (one fill kernel instead of one per
    gradient accumulator; every tensor starts on a 16-byte boundary)."""
    sizes = [int(np.prod(sh)) if len(sh) else 1 for sh in shapes]
    offs, total = [], 0
    for k in sizes:
        offs.append(total)
        total += (k + 3) // 4 * 4
    flat = torch.zeros((total,), device=device, dtype=torch.float32)
    return [flat[o:o + k].view(sh) for o, k, sh in zip(offs, sizes, shapes)]


def _padded_rows(n: int, cols: int, device) -> torch.Tensor:
    """[n, cols] fp32 view whose rows start on 16-byte boundaries (leading dimension rounded up to 4 floats)."""
    ld = (cols + 3) // 4 * 4
    return torch.empty((n, ld), device=device, dtype=torch.float32)[:, :cols]


# ------------------------------------------------------------------------------------------------
# hash grid (A9/A10)
# ------------------------------------------------------------------------------------------------
def hash_resolutions(min_res: int, max_res: int, num_levels: int) -> List[float]:
    """floor(min_res * growth**l) exactly as the reference evaluates it (encodings.py:195-197,227):
    growth in float64 numpy, the power/product/floor on a float32 torch tensor."""
    growth = np.exp((np.log(max_res) - np.log(min_res)) / (num_levels - 1)) if num_levels > 1 else 1.0
    levels = torch.arange(num_levels)
    return torch.floor(min_res * growth ** levels).to(torch.float32).tolist()


def make_hashgrid_desc(num_levels, features_per_level, log2_hashmap_size, resolutions, radius=0.0,
                       interpolation="Linear") -> MmsbHashGridDesc:
    if interpolation not in INTERP:
        raise ValueError(f"interpolation '{interpolation}' is not supported")
    d = MmsbHashGridDesc()
    d.num_levels = int(num_levels)
    d.features_per_level = int(features_per_level)
    d.log2_hashmap_size = int(log2_hashmap_size)
    d.interpolation = INTERP[interpolation]
    d.radius = float(radius)
    for i, r in enumerate(resolutions):
        d.resolution[i] = float(r)
    return d


def hashgrid_fwd_into(desc, x, table, mask, out, col_offset=0, idx_out=None):
    """Writes the L*F features of x [n,3] into out[:, col_offset:col_offset+L*F] (row stride = out.stride(0))."""
    n = x.shape[0]
    out_view = out[:, col_offset:]
    call("mmsb_hashgrid_fwd", ctypes.byref(desc), ptr(x), _i64(x.stride(0)), ptr(table), ptr(mask),
         ptr(out_view), _i64(out.stride(0)), ptr(idx_out), _i64(n), stream_ptr())


def hashgrid_bwd_from(desc, x, table, mask, dout, col_offset, dtable, dx):
    n = x.shape[0]
    dview = dout[:, col_offset:]
    call("mmsb_hashgrid_bwd", ctypes.byref(desc), ptr(x), _i64(x.stride(0)), ptr(table), ptr(mask),
         ptr(dview), _i64(dout.stride(0)), ptr(dtable), ptr(dx), _i64(dx.stride(0) if dx is not None else 3),
         _i64(n), stream_ptr())


class HashGridFn(torch.autograd.Function):
    """features = HashEncoding(x) [* level mask, after the FeatureGrid rescale when desc.radius > 0]."""

    @staticmethod
    def forward(ctx, x, table, mask, desc):
        x2 = _rows(x, 3)
        table = _f(table)
        nf = desc.num_levels * desc.features_per_level
        out = torch.empty((x2.shape[0], nf), device=x2.device, dtype=torch.float32)
        hashgrid_fwd_into(desc, x2, table, mask, out)
        ctx.save_for_backward(x2, table, mask)
        ctx.desc = desc
        ctx.x_shape = x.shape
        return out.reshape(*x.shape[:-1], nf)

    @staticmethod
    def backward(ctx, dout):
        x2, table, mask = ctx.saved_tensors
        desc = ctx.desc
        nf = desc.num_levels * desc.features_per_level
        dout = _f(dout).reshape(-1, nf)
        dtable = torch.zeros_like(table) if ctx.needs_input_grad[1] else None
        dx = torch.empty_like(x2) if ctx.needs_input_grad[0] else None
        if dtable is not None or dx is not None:
            hashgrid_bwd_from(desc, x2, table, mask, dout, 0, dtable, dx)
        return (dx.reshape(ctx.x_shape) if dx is not None else None), dtable, None, None


def hashgrid_indices(desc, x, table):
    """int64 [n, L, 8] corner rows (reference order hashed_0..7) — parity surface."""
    x2 = _rows(x, 3)
    nf = desc.num_levels * desc.features_per_level
    out = torch.empty((x2.shape[0], nf), device=x2.device, dtype=torch.float32)
    idx = torch.empty((x2.shape[0], desc.num_levels, 8), device=x2.device, dtype=torch.int64)
    hashgrid_fwd_into(desc, x2, _f(table), None, out, 0, idx)
    return idx, out


# ------------------------------------------------------------------------------------------------
# NeRF / SH encodings (A8, A15)
# ------------------------------------------------------------------------------------------------
def nerf_freqs(min_freq_exp, max_freq_exp, num_frequencies):
    return (2 ** torch.linspace(min_freq_exp, max_freq_exp, num_frequencies)).to(torch.float32).tolist()


def nerf_out_dim(in_dim, num_freqs, include_input):
    return in_dim * num_freqs * 2 + (in_dim if include_input else 0)


def nerf_fwd_into(x, freqs, include_input, out, col_offset=0):
    n, d = x.shape
    fr = (_f32 * len(freqs))(*freqs)
    call("mmsb_nerf_encoding_fwd", ptr(x), _i64(x.stride(0)), _i32(d), fr, _i32(len(freqs)), _i32(int(include_input)),
         ptr(out[:, col_offset:]), _i64(out.stride(0)), _i64(n), stream_ptr())


def nerf_bwd_from(x, freqs, include_input, dout, col_offset, dx, accumulate):
    n, d = x.shape
    fr = (_f32 * len(freqs))(*freqs)
    call("mmsb_nerf_encoding_bwd", ptr(x), _i64(x.stride(0)), _i32(d), fr, _i32(len(freqs)), _i32(int(include_input)),
         ptr(dout[:, col_offset:]), _i64(dout.stride(0)), ptr(dx), _i64(dx.stride(0)), _i32(int(accumulate)), _i64(n),
         stream_ptr())


class NerfEncodingFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, freqs, include_input):
        d = x.shape[-1]
        x2 = _rows(x, d)
        od = nerf_out_dim(d, len(freqs), include_input)
        out = torch.empty((x2.shape[0], od), device=x2.device, dtype=torch.float32)
        nerf_fwd_into(x2, freqs, include_input, out)
        ctx.save_for_backward(x2)
        ctx.cfg = (tuple(freqs), include_input, x.shape, od)
        return out.reshape(*x.shape[:-1], od)

    @staticmethod
    def backward(ctx, dout):
        (x2,) = ctx.saved_tensors
        freqs, include_input, shape, od = ctx.cfg
        dx = torch.empty_like(x2)
        nerf_bwd_from(x2, list(freqs), include_input, _f(dout).reshape(-1, od), 0, dx, False)
        return dx.reshape(shape), None, None


class SHEncodingFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, dirs, levels):
        d2 = _rows(dirs, 3)
        nc = levels * levels
        out = torch.empty((d2.shape[0], nc), device=d2.device, dtype=torch.float32)
        call("mmsb_sh_encoding_fwd", ptr(d2), _i64(3), _i32(levels), ptr(out), _i64(nc), _i64(d2.shape[0]), stream_ptr())
        ctx.save_for_backward(d2)
        ctx.cfg = (levels, dirs.shape)
        return out.reshape(*dirs.shape[:-1], nc)

    @staticmethod
    def backward(ctx, dout):
        (d2,) = ctx.saved_tensors
        levels, shape = ctx.cfg
        nc = levels * levels
        dd = torch.empty_like(d2)
        g = _f(dout).reshape(-1, nc)
        call("mmsb_sh_encoding_bwd", ptr(d2), _i64(3), _i32(levels), ptr(g), _i64(nc), ptr(dd), _i64(3),
             _i64(d2.shape[0]), stream_ptr())
        return dd.reshape(shape), None


# ------------------------------------------------------------------------------------------------
# MLP input assembly: cat[x, encodings(x), features, hash features] written in place by the encoders
# (feature_structures.py:164 / radiance_field.py:73 / nerf_field.py:101 build these rows with torch.cat)
# ------------------------------------------------------------------------------------------------
def copy_piece(t):
    return ("copy", t)


def nerf_piece(x, freqs, include_input):
    return ("nerf", x, tuple(freqs), bool(include_input))


def hash_piece(x, table, mask, desc):
    return ("hash", x, table, mask, desc)


def _piece_width(p):
    if p[0] == "copy":
        return p[1].shape[-1]
    if p[0] == "nerf":
        return nerf_out_dim(p[1].shape[-1], len(p[2]), p[3])
    return p[4].num_levels * p[4].features_per_level


# Zero-copy "copy" pieces.  A wide materialised piece (the SDF network's 256 geometry features in the radiance MLP's input
# row, radiance_field.py:90-101) costs a pass over 2 x n x 256 floats when it is copied into the assembled row.  Instead
# the producer writes it where it will be needed: AssembleFn records where a wide copy piece ended up (_ROW_HINT: rows,
# width -> leading dimension, total width, column); on the next call the producer (`row_slot`) allocates the whole row
# buffer, registers it (_ROW_BUFS) and returns the piece's column slice as its output; AssembleFn recognises the slice by
# its storage and ADOPTS the buffer as its output — the other pieces are written around the one that is already there.
# Any mismatch (another preset, a reshaped piece, a stale hint) falls back to the copy.
_ROW_HINT = {}
_ROW_BUFS = {}
ROW_ADOPT = int(os.environ.get("MMSB_ROW_ADOPT", "1"))
ROW_ADOPT_MIN_WIDTH = 64
ROW_ADOPTIONS = 0            # how often AssembleFn adopted a producer's row buffer (tests / dev)


def row_slot(n: int, width: int, device):
    """Output buffer [n, width] for a producer whose result will be a copy piece of an assembled row: a column slice of a
    fresh, registered row buffer when the layout is known from an earlier call, else a plain tensor."""
    hint = _ROW_HINT.get((n, width)) if ROW_ADOPT else None
    if hint is None:
        return torch.empty((n, width), device=device, dtype=torch.float32)
    ld, total, col = hint
    buf = torch.empty((n, ld), device=device, dtype=torch.float32)
    _ROW_BUFS[buf.untyped_storage().data_ptr()] = (n, ld, total, col, width)
    return buf[:, col:col + width]


def _adoptable(t, n, ld, total, col, width):
    if not (ROW_ADOPT and t.dim() == 2 and t.is_cuda and t.dtype == torch.float32 and t.stride(1) == 1 and t.stride(0) == ld):
        return False
    st = t.untyped_storage()
    return (_ROW_BUFS.get(st.data_ptr()) == (n, ld, total, col, width) and t.storage_offset() == col
            and st.nbytes() >= n * ld * 4)


class AssembleFn(torch.autograd.Function):
    """rows = cat(pieces, -1), every piece written by its producer kernel straight into its column range of one
    16-byte-aligned row buffer (no torch.cat, no second pass); backward reads the column ranges of the incoming
    gradient in place.  spec: tuple of ("copy", i) | ("nerf", i, freqs, include_input) | ("hash", i, j, mask, desc)
    where i / j index `tensors` (a tensor feeding several pieces appears once: its gradient is accumulated here)."""

    @staticmethod
    def forward(ctx, spec, *tensors):
        n = tensors[spec[0][1]].reshape(-1, tensors[spec[0][1]].shape[-1]).shape[0]
        widths = []
        for p in spec:
            if p[0] == "copy":
                widths.append(tensors[p[1]].shape[-1])
            elif p[0] == "nerf":
                widths.append(nerf_out_dim(tensors[p[1]].shape[-1], len(p[2]), p[3]))
            else:
                widths.append(p[4].num_levels * p[4].features_per_level)
        total = sum(widths)
        dev = tensors[0].device
        ld = (total + 3) // 4 * 4
        out, in_place, col = None, None, 0
        for i, (p, w) in enumerate(zip(spec, widths)):
            if p[0] == "copy" and w >= ROW_ADOPT_MIN_WIDTH:
                t = tensors[p[1]]
                if out is None and _adoptable(t, n, ld, total, col, w):
                    # the piece already sits in its column range of a row buffer of this very layout: adopt the buffer
                    _ROW_BUFS.pop(t.untyped_storage().data_ptr(), None)
                    global ROW_ADOPTIONS
                    ROW_ADOPTIONS += 1
                    # (a fresh tensor over the same storage, not an autograd view of the piece)
                    out = torch.empty((0,), device=dev, dtype=torch.float32).set_(t.untyped_storage(), 0, (n, total), (ld, 1))
                    in_place = i
                elif t.dim() == 2 and t.shape[0] == n:
                    _ROW_HINT[(n, w)] = (ld, total, col)
            col += w
        if out is None:
            out = _padded_rows(n, total, dev)
        saved, col = {}, 0
        for i, (p, w) in enumerate(zip(spec, widths)):
            if p[0] == "copy":
                if i != in_place:
                    src = _rows(tensors[p[1]].reshape(n, w), w)
                    call("mmsb_copy_rows", ptr(src), _i64(src.stride(0)), ptr(out[:, col:]), _i64(out.stride(0)), _i64(n),
                         _i32(w), stream_ptr())
            elif p[0] == "nerf":
                x2 = saved.setdefault(p[1], _rows(tensors[p[1]], tensors[p[1]].shape[-1]))
                nerf_fwd_into(x2, list(p[2]), p[3], out, col)
            else:
                x2 = saved.setdefault(p[1], _rows(tensors[p[1]], 3))
                tab = saved.setdefault(p[2], _f(tensors[p[2]]))
                hashgrid_fwd_into(p[4], x2, tab, p[3], out, col)
            col += w
        keys = sorted(saved)
        ctx.save_for_backward(*[saved[k] for k in keys])
        ctx.keys, ctx.spec, ctx.widths = keys, spec, widths
        ctx.shapes = [t.shape for t in tensors]
        return out

    @staticmethod
    def backward(ctx, dout):
        saved = dict(zip(ctx.keys, ctx.saved_tensors))
        spec, widths = ctx.spec, ctx.widths
        n, total = dout.shape
        if not (dout.stride(1) == 1 and dout.dtype == torch.float32):
            dout = dout.contiguous().float()
        grads = [None] * len(ctx.shapes)
        written = set()
        cols, col = [], 0
        for w in widths:
            cols.append(col)
            col += w
        # hash pieces first (the kernel writes dx), then the nerf pieces accumulate into the same dx
        order = sorted(range(len(spec)), key=lambda i: {"hash": 0, "nerf": 1, "copy": 2}[spec[i][0]])
        for i in order:
            p, w, c = spec[i], widths[i], cols[i]
            if p[0] == "copy":
                if ctx.needs_input_grad[1 + p[1]]:
                    g = dout[:, c:c + w].reshape(ctx.shapes[p[1]]) if dout[:, c:c + w].is_contiguous() else dout[:, c:c + w]
                    g = g if g.shape == ctx.shapes[p[1]] else g.reshape(ctx.shapes[p[1]])
                    grads[p[1]] = g if grads[p[1]] is None else grads[p[1]] + g
            elif p[0] == "hash":
                x2, tab = saved[p[1]], saved[p[2]]
                need_x, need_t = ctx.needs_input_grad[1 + p[1]], ctx.needs_input_grad[1 + p[2]]
                if not (need_x or need_t):
                    continue
                dtable = torch.zeros_like(tab) if need_t else None
                dx = None
                if need_x:
                    if p[1] in written:
                        raise NotImplementedError("two hash pieces on one input")
                    dx = torch.empty((n, 3), device=dout.device, dtype=torch.float32)
                    written.add(p[1])
                    grads[p[1]] = dx
                hashgrid_bwd_from(p[4], x2, tab, p[3], dout, c, dtable, dx)
                if need_t:
                    grads[p[2]] = dtable if grads[p[2]] is None else grads[p[2]] + dtable
            else:
                if not ctx.needs_input_grad[1 + p[1]]:
                    continue
                x2 = saved[p[1]]
                if grads[p[1]] is None:
                    grads[p[1]] = torch.empty((n, x2.shape[1]), device=dout.device, dtype=torch.float32)
                    nerf_bwd_from(x2, list(p[2]), p[3], dout, c, grads[p[1]], False)
                else:
                    nerf_bwd_from(x2, list(p[2]), p[3], dout, c, grads[p[1]], True)
        for i, g in enumerate(grads):
            if g is not None and tuple(g.shape) != tuple(ctx.shapes[i]):
                grads[i] = g.reshape(ctx.shapes[i])
        return (None, *grads)


def assemble(pieces):
    """pieces: list from copy_piece / nerf_piece / hash_piece -> row-strided [n, sum(widths)] tensor."""
    tensors, index = [], {}

    def idx(t):
        k = id(t)
        if k not in index:
            index[k] = len(tensors)
            tensors.append(t)
        return index[k]

    spec = []
    for p in pieces:
        if p[0] == "copy":
            spec.append(("copy", idx(p[1])))
        elif p[0] == "nerf":
            spec.append(("nerf", idx(p[1]), p[2], p[3]))
        else:
            spec.append(("hash", idx(p[1]), idx(p[2]), p[3], p[4]))
    return AssembleFn.apply(tuple(spec), *tensors)


# ------------------------------------------------------------------------------------------------
# MLP (A11)
# ------------------------------------------------------------------------------------------------
def linear_fwd(x, w, b, act, act_param, out=None):
    n, k = x.shape
    o = w.shape[0]
    if out is None:
        out = torch.empty((n, o), device=x.device, dtype=torch.float32)
    call("mmsb_linear_fwd", ptr(x), _i64(x.stride(0)), ptr(w), ptr(b), ptr(out), _i64(out.stride(0)), _i64(n), _i32(k),
         _i32(o), _i32(act), _f32(act_param), stream_ptr())
    return out


def pack_weight(w, transpose: bool, precision: int):
    """Pre-split / pre-swizzled operand of the tcgen05 layer kernels (mmsb_linear_pack_weight)."""
    o, k = w.shape
    n_dim, k_dim = (k, o) if transpose else (o, k)
    size = int(_lib.load_library().mmsb_linear_packed_size(n_dim, k_dim, precision))
    if size < 0:
        raise ValueError(f"pack_weight: bad arguments {tuple(w.shape)} precision {precision}")
    packed = torch.empty((size,), device=w.device, dtype=torch.float32)
    call("mmsb_linear_pack_weight", ptr(w), _i64(w.stride(0)), _i32(o), _i32(k), _i32(1 if transpose else 0),
         _i32(precision), ptr(packed), stream_ptr())
    return packed


def amax_of(x, out=None):
    """Device scalar max |x| of a row-strided [n, cols] fp32 matrix (mmsb_amax): the scale source of a precision-2
    (fp16 split) product whose streamed operand was not produced by a tensor-core kernel."""
    if out is None:
        out = torch.zeros((1,), device=x.device, dtype=torch.float32)
    x2 = x if x.dim() == 2 else x.reshape(-1, 1)
    call("mmsb_amax", ptr(x2), _i64(x2.stride(0)), _i64(x2.shape[0]), _i32(x2.shape[1]), ptr(out), stream_ptr())
    return out


F16_MIN_ROWS = 2 * 128 * 74      # the CTA-pair kernels need enough 256-row tiles to fill the machine


def f16_eligible(x, out_dim) -> bool:
    """Shapes the precision-2 (fp16 split) kernels cover: see include/mms_b200.h."""
    # (narrow inputs such as the 71-column first SDF layer are epilogue-bound: measured no gain, so they stay on 3xTF32)
    return (out_dim % 256 == 0 and x.shape[1] >= 128 and x.shape[0] >= F16_MIN_ROWS and x.stride(1) == 1
            and x.stride(0) % 4 == 0 and x.data_ptr() % 16 == 0)


def layer_precision(x, out_dim) -> int:
    """Arithmetic of one layer product under the global MLP_PRECISION: mode 2 uses the fp16 split where its kernels
    apply and 3xTF32 elsewhere (both are fp32-accurate)."""
    if MLP_PRECISION == 2:
        return 2 if f16_eligible(x, out_dim) else 3
    return MLP_PRECISION


def linear_fwd_tc(x, packed_w, b, out_dim, act, act_param, precision, out=None, x_amax=None, y_amax=None):
    n, k = x.shape
    if out is None:
        out = torch.empty((n, out_dim), device=x.device, dtype=torch.float32)
    if precision == 2 and x_amax is None:
        x_amax = amax_of(x)
    call("mmsb_linear_fwd_tc", ptr(x), _i64(x.stride(0)), ptr(packed_w), ptr(b), ptr(out), _i64(out.stride(0)), _i64(n),
         _i32(k), _i32(out_dim), _i32(act), _f32(act_param), _i32(precision), ptr(x_amax), ptr(y_amax), stream_ptr())
    return out


def linear_bwd_data_tc(dz, packed_wt, in_dim, y_prev, act_prev, act_param, precision, out=None, accumulate=False):
    n, o = dz.shape
    if out is None:
        out = torch.empty((n, in_dim), device=dz.device, dtype=torch.float32)
    call("mmsb_linear_bwd_data_tc", ptr(dz), _i64(dz.stride(0)), ptr(packed_wt), ptr(out), _i64(out.stride(0)),
         ptr(y_prev), _i64(y_prev.stride(0) if y_prev is not None else 0), _i32(act_prev), _f32(act_param), _i64(n),
         _i32(in_dim), _i32(o), _i32(precision), _i32(int(accumulate)), stream_ptr())
    return out


def linear_bwd_weight_tc(dz, x, dw, db, precision):
    n, o = dz.shape
    k = x.shape[1]
    call("mmsb_linear_bwd_weight_tc", ptr(dz), _i64(dz.stride(0)), ptr(x), _i64(x.stride(0)), ptr(dw), ptr(db), _i64(n),
         _i32(k), _i32(o), _i32(precision), stream_ptr())

# MLP arithmetic mode of the wide layers: 3 = tcgen05 3xTF32 (fp32-accurate, default), 1 = tcgen05 single-pass TF32
# (1e-2 band, the precision class of the reference's fp16-autocast GPU runs), 0 = fp32 SIMT GEMM.
MLP_PRECISION = int(os.environ.get("MMSB_MLP_PRECISION", "3"))
_PACK_CACHE = {}


def set_mlp_precision(precision: int) -> None:
    global MLP_PRECISION
    if precision not in (0, 1, 2, 3):
        raise ValueError("MLP precision must be 0 (fp32 SIMT), 1 (TF32), 2 (fp16 split + 3xTF32) or 3 (3xTF32)")
    MLP_PRECISION = precision
    _PACK_CACHE.clear()


def clear_pack_cache() -> None:
    """Call after every optimiser step (the packed operands are functions of the weights)."""
    _PACK_CACHE.clear()
    _PERM_CACHE.clear()
    _ROW_BUFS.clear()


_PERM_CACHE = {}


def permuted_columns(w, perm):
    """w[:, perm] (perm: tuple of old column indices in the new order), memoised like packed_weight so that the packed
    operands of the permuted weight are reused within a step.  The MLP input rows are assembled with the hash features
    first (sector-aligned stores of the hash-grid kernel); the first layer's weight follows with this column gather,
    whose autograd backward scatters the gradient back to the reference's column order."""
    key = (id(w), perm)
    hit = _PERM_CACHE.get(key)
    if hit is not None and hit[0]() is w and hit[1] == w._version and hit[2].requires_grad == (w.requires_grad and torch.is_grad_enabled()):
        return hit[2]
    idx = const_tensor(("perm", perm), lambda: torch.tensor(perm, dtype=torch.long), w.device)
    out = w.index_select(1, idx)
    if len(_PERM_CACHE) > 128:
        _PERM_CACHE.clear()
    _PERM_CACHE[key] = (weakref.ref(w), w._version, out)
    return out


def _use_tc(w) -> bool:
    # every layer, also the narrow ones (sdf-only column, modality heads): the centre and tap evaluations of the SDF
    # must share one arithmetic, otherwise the finite differences (Hessian ~ 1/delta^2) amplify the path difference
    return MLP_PRECISION != 0


def packed_weight(w, transpose: bool, precision: int, rows=None):
    """Packed operand of w (or of its row range `rows` = (first, last + 1)), memoised per tensor object and version
    (valid while torch.nn.utils.parametrize.cached() keeps the effective weight alive, i.e. for one step)."""
    key = (id(w), bool(transpose), precision, rows)
    hit = _PACK_CACHE.get(key)
    if hit is not None and hit[0]() is w and hit[1] == w._version:
        return hit[2]
    src = _f(w.detach())
    if rows is not None:
        src = src[rows[0]:rows[1]]
    packed = pack_weight(src, transpose, precision)
    if len(_PACK_CACHE) > 512:
        _PACK_CACHE.clear()
    _PACK_CACHE[key] = (weakref.ref(w), w._version, packed)
    return packed


class _OutActToken:
    """Shared between the MLP that produced a tensor y = act(z) and the split_rows that consumes it: when `applied` is
    set, the gradient the producer receives is already dL/dz (the consumers folded act'(y) into their dgrad epilogues)."""
    __slots__ = ("act", "act_param", "applied")

    def __init__(self, act, act_param):
        self.act, self.act_param, self.applied = act, act_param, False


class _GradSink:
    """Shared gradient buffer of the row blocks of one tensor (see split_rows)."""
    __slots__ = ("buf", "rows", "cols", "claimed", "token")

    def __init__(self, rows, cols, token=None):
        self.buf, self.rows, self.cols, self.claimed, self.token = None, rows, cols, set(), token

    def block(self, off, n, device):
        """The rows [off, off + n) of the buffer for the FIRST consumer that asks for them (None for any later one: its
        gradient is then summed by autograd and split_rows falls back to a concatenation)."""
        if off in self.claimed:
            return None
        self.claimed.add(off)
        if self.buf is None:
            self.buf = torch.empty((self.rows, self.cols), device=device, dtype=torch.float32)
        return self.buf[off:off + n]


class SplitRowsFn(torch.autograd.Function):
    """x [N, C] -> consecutive row blocks of the given sizes (torch.split along dim 0).  Backward: when every block's
    consumer (MLPFn.backward, first layer) has written its input gradient straight into the block's rows of the shared
    sink buffer, that buffer IS dL/dx — no concatenation pass over N x C; otherwise torch.cat as autograd would do."""

    @staticmethod
    def forward(ctx, x, sink, *sizes):
        ctx.sink, ctx.sizes = sink, sizes
        if sink.token is not None:
            ctx.save_for_backward(x)
        return tuple(torch.split(x, list(sizes), dim=0))

    @staticmethod
    def backward(ctx, *grads):
        sink, sizes = ctx.sink, ctx.sizes
        buf, sink.buf = sink.buf, None
        tok = sink.token
        in_sink, off = [], 0
        for g, n in zip(grads, sizes):
            in_sink.append(n == 0 or (buf is not None and g is not None and g.dim() == 2 and tuple(g.shape) == (n, sink.cols)
                                      and g.stride(1) == 1 and g.stride(0) == buf.stride(0) and g.data_ptr() == buf[off].data_ptr()))
            off += n
        if all(in_sink):
            if tok is not None:
                tok.applied = True
            return (buf, None) + (None,) * len(sizes)
        dev = next(g.device for g in grads if g is not None)
        parts = [g if g is not None else torch.zeros((n, sink.cols), device=dev) for g, n in zip(grads, sizes)]
        if tok is not None:
            # the blocks that went through the sink already carry act'(y); bring the others to the same state
            (y,) = ctx.saved_tensors
            off = 0
            for i, n in enumerate(sizes):
                if n > 0 and not in_sink[i] and grads[i] is not None:
                    gi, yi = _rows(parts[i], sink.cols), y[off:off + n]
                    out = torch.empty((n, sink.cols), device=dev, dtype=torch.float32)
                    call("mmsb_act_bwd", ptr(gi), _i64(gi.stride(0)), ptr(yi), _i64(yi.stride(0)), ptr(out), _i64(out.stride(0)),
                         _i64(n), _i32(sink.cols), _i32(tok.act), _f32(tok.act_param), stream_ptr())
                    parts[i] = out
                off += n
            tok.applied = True
        return (torch.cat(parts, 0), None) + (None,) * len(sizes)


SPLIT_FOLD = int(os.environ.get("MMSB_SPLIT_FOLD", "1"))     # dev switch: 0 keeps the producer's own activation-derivative pass


def split_rows(x, sizes, single_consumer: bool = False):
    """torch.split(x, sizes, dim=0) for a 2-D x whose blocks feed MLPs (the per-modality heads on the radiance features,
    model_components.RadianceModel): the blocks carry a handle of a shared gradient buffer, MLPFn writes the input
    gradient of its first layer into it, and the backward of the split returns the buffer instead of concatenating.
    `single_consumer` (every block feeds exactly ONE MLP): when x is the output y = act(z) of an MLP, that MLP's
    output-activation derivative is folded into the consumers' dgrad epilogues (dx = (dz W) * act'(y): the epilogue's
    `y_prev` operand) instead of a pass of its own over x.  The caller guarantees that x has no consumer besides this
    split (RadianceModel: the radiance features only feed the heads); without that guarantee pass False."""
    if x.dim() != 2 or not (torch.is_grad_enabled() and x.requires_grad):
        return torch.split(x, list(sizes), dim=0)
    token = getattr(x, "_mmsb_out_act", None) if (single_consumer and SPLIT_FOLD) else None
    sink = _GradSink(x.shape[0], x.shape[1], token)
    outs = SplitRowsFn.apply(x, sink, *[int(n) for n in sizes])
    off = 0
    for o, n in zip(outs, sizes):
        o._mmsb_sink = (sink, off)
        off += int(n)
    return outs


class MLPFn(torch.autograd.Function):
    """y = MLP(x): layers [(W_i [out,in], b_i)], hidden activation, output activation, optional
    skip connections (the layer input becomes cat([h, x]) / sqrt(2), mlp.py:164-165).
    args = (x, hidden_act, act_param, out_act, skips(tuple), n_out_used, W0, b0, W1, b1, ...)
    n_out_used: if not None only the first n_out_used outputs of the last layer are evaluated
    (the sdf-only evaluations of surface_model.py:143-146 / get_sdf)."""

    @staticmethod
    def forward(ctx, x, hidden_act, act_param, out_act, skips, n_out_used, out_token, *params):
        nl = len(params) // 2
        in_dim = x.shape[-1]
        x2 = _rows(x, in_dim)
        # skip-connection networks (`mlp*` presets) run on the tensor cores too: the concatenated input of the skip layer
        # has an odd row stride, which the register-staged operand producers take (no TMA)
        prec = MLP_PRECISION
        bwd_prec = 3 if prec == 2 else prec
        amax = None                  # max |h| of the running activation when a tensor-core epilogue produced it (mode 2)
        ws = [_f(params[2 * i]) for i in range(nl)]
        bs = [_f(params[2 * i + 1]) if params[2 * i + 1] is not None else None for i in range(nl)]
        if n_out_used is not None:
            ws[-1] = ws[-1][:n_out_used].contiguous()
            bs[-1] = bs[-1][:n_out_used].contiguous() if bs[-1] is not None else None
        acts_in = [x2]       # input of every layer
        h = x2
        need_grad = any(ctx.needs_input_grad)
        packed_t = [None] * nl   # W^T operands of the dgrad products (packed once per step, see packed_weight)
        for i in range(nl):
            if i in skips:
                h = torch.cat([h, x2], -1) / math.sqrt(2)
                acts_in[i] = h
                amax = None
            a = hidden_act if i < nl - 1 else out_act
            if prec != 0 and _use_tc(ws[i]):
                src = params[2 * i] if (n_out_used is None or i < nl - 1) else ws[i]
                # mode 2: fp16 split for the forward products it covers (the output's amax comes out of the epilogue and
                # scales the next layer's operand), 3xTF32 for everything else including the backward products
                pf = layer_precision(h, ws[i].shape[0]) if prec == 2 else prec
                y_amax = torch.zeros((1,), device=h.device) if prec == 2 and i < nl - 1 else None
                h = linear_fwd_tc(h, packed_weight(src, False, pf), bs[i], ws[i].shape[0], a, act_param, pf,
                                  x_amax=amax if pf == 2 else None, y_amax=y_amax)
                amax = y_amax
                if need_grad and (i > 0 or ctx.needs_input_grad[0]):
                    packed_t[i] = packed_weight(src, True, bwd_prec)
            else:
                h = linear_fwd(h, ws[i], bs[i], a, act_param)
                amax = None
            acts_in.append(h)
        ctx.save_for_backward(*acts_in, *ws)
        ctx.prec = bwd_prec
        ctx.packed_t = packed_t
        ctx.sink = getattr(x, "_mmsb_sink", None) if x.dim() == 2 else None
        ctx.out_token = out_token
        ctx.cfg = (nl, hidden_act, act_param, out_act, tuple(skips), n_out_used, x.shape, in_dim,
                   [p is not None for p in params], [tuple(params[2 * i].shape) for i in range(nl)])
        return h.reshape(*x.shape[:-1], h.shape[-1])

    @staticmethod
    def backward(ctx, dy):
        nl, hidden_act, act_param, out_act, skips, n_out_used, x_shape, in_dim, has, wshapes = ctx.cfg
        prec = ctx.prec
        saved = ctx.saved_tensors
        acts = saved[: nl + 1]
        ws = saved[nl + 1:]
        x2 = acts[0] if 0 not in skips else None
        n = acts[0].shape[0]
        out_dim = acts[nl].shape[1]
        dz = _rows(dy.reshape(n, out_dim) if dy.dim() != 2 else dy, out_dim)
        # activation derivative of the output layer (unless the consumers of the output have folded it into their dgrads)
        if out_act != 0 and not (ctx.out_token is not None and ctx.out_token.applied):
            dz_new = _padded_rows(n, out_dim, dz.device)
            call("mmsb_act_bwd", ptr(dz), _i64(dz.stride(0)), ptr(acts[nl]), _i64(acts[nl].stride(0)), ptr(dz_new),
                 _i64(dz_new.stride(0)), _i64(n), _i32(out_dim), _i32(out_act), _f32(act_param), stream_ptr())
            dz = dz_new
        grads = [None] * (2 * nl)
        dx_skip = None
        need_dx = ctx.needs_input_grad[0]
        want = [ctx.needs_input_grad[7 + 2 * i] or (has[2 * i + 1] and ctx.needs_input_grad[8 + 2 * i]) for i in range(nl)]
        shapes = []
        for i in range(nl):
            if want[i]:
                shapes += [tuple(ws[i].shape)] + ([(ws[i].shape[0],)] if has[2 * i + 1] else [])
        bufs = iter(_zeros_many(shapes, ws[0].device)) if shapes else iter(())
        acc = {}
        for i in range(nl):
            if want[i]:
                acc[i] = (next(bufs), next(bufs) if has[2 * i + 1] else None)
        for i in range(nl - 1, -1, -1):
            w = ws[i]
            o, k = w.shape
            xin = acts[i]
            tc = prec != 0 and _use_tc(w)
            if want[i]:
                dw, db = acc[i]
                if tc:
                    linear_bwd_weight_tc(dz, xin, dw, db, prec)
                else:
                    call("mmsb_linear_bwd_weight", ptr(dz), _i64(dz.stride(0)), ptr(xin), _i64(xin.stride(0)), ptr(dw),
                         ptr(db), _i64(n), _i32(k), _i32(o), stream_ptr())
                if i == nl - 1 and n_out_used is not None:
                    full_w = torch.zeros(wshapes[i], device=w.device, dtype=torch.float32)
                    full_w[:n_out_used] = dw
                    dw = full_w
                    if db is not None:
                        full_b = torch.zeros((wshapes[i][0],), device=w.device, dtype=torch.float32)
                        full_b[:n_out_used] = db
                        db = full_b
                grads[2 * i] = dw
                grads[2 * i + 1] = db
            if i == 0 and not need_dx:
                break
            dxin, fold = None, None
            if i == 0 and tc and ctx.sink is not None and 0 not in skips and k == ctx.sink[0].cols and k % 4 == 0:
                dxin = ctx.sink[0].block(ctx.sink[1], n, w.device)      # straight into the split's shared gradient buffer
                if dxin is not None:
                    fold = ctx.sink[0].token     # x = act(z) of the producing MLP: fold act'(x) into this dgrad's epilogue
            if dxin is None:
                dxin = _padded_rows(n, k, w.device)
            # the input of layer i is the hidden activation of layer i-1 (unless a skip concat sits between)
            fuse_prev = i > 0 and i not in skips
            if tc and fold is not None:
                linear_bwd_data_tc(dz, ctx.packed_t[i], k, xin, fold.act, fold.act_param, prec, out=dxin)
            elif tc:
                linear_bwd_data_tc(dz, ctx.packed_t[i], k, xin if fuse_prev else None,
                                   hidden_act if fuse_prev else 0, act_param, prec, out=dxin)
            else:
                call("mmsb_linear_bwd_data", ptr(dz), _i64(dz.stride(0)), ptr(w), ptr(dxin), _i64(dxin.stride(0)),
                     ptr(xin) if fuse_prev else None, _i64(xin.stride(0)), _i32(hidden_act if fuse_prev else 0),
                     _f32(act_param), _i64(n), _i32(k), _i32(o), stream_ptr())
            if i in skips:
                dxin = dxin / math.sqrt(2)
                hk = k - in_dim
                d_skip = dxin[:, hk:]
                dx_skip = d_skip if dx_skip is None else dx_skip + d_skip
                dh = dxin[:, :hk].contiguous()
                if i > 0:
                    hprev = acts[i][:, :hk] * math.sqrt(2)   # undo the /sqrt(2) to recover y_{i-1}
                    dz = torch.empty_like(dh)
                    call("mmsb_act_bwd", ptr(dh), _i64(hk), ptr(hprev.contiguous()), _i64(hk), ptr(dz), _i64(hk), _i64(n),
                         _i32(hk), _i32(hidden_act), _f32(act_param), stream_ptr())
                else:
                    dz = dh
            else:
                dz = dxin
        dx = None
        if need_dx:
            dx = dz if dx_skip is None else dz + dx_skip
            dx = dx.reshape(x_shape)
        return (dx, None, None, None, None, None, None, *grads)


class SdfNetFn(torch.autograd.Function):
    """The SDF network (surface_field.py:99-116, mlp.py:152-171) with its last layer split into the sdf head (output 0)
    and the geometry features (outputs 1..G):  x [n, in] -> sdf [n, 1] for every row, geo [n_full, G] for the "full"
    rows (the centre evaluations; the tap and sampler evaluations keep only the sdf, surface_model.py:143-146).
    Row layout: `group` = 1: the first n_full rows are the full ones; `group` = g > 1: the rows come in groups of g
    (one sample's centre evaluation followed by its g - 1 finite-difference taps) and row 0 of every group is full
    (n_full = n / g).  The grouped layout is what SurfaceModel uses: the + / - contributions of a sample's taps to every
    weight gradient are ~1/(4 delta') larger than their sum, and the tensor core's accumulators TRUNCATE (round toward
    zero) on every MMA — a systematic error proportional to the running sum.  With a sample's rows adjacent the
    cancellation happens inside a few MMAs and the running sums stay at the size of the net gradient; with the taps in
    blocks of their own (round 1) every CTA of the weight-gradient kernel accumulated one sign only and the error was
    ~200 x the fp32 reference's at a BASELINE batch size (tests/test_gpu_model.py::test_whole_step_at_baseline_size_vs_oracle).
    The sdf head never runs as a layer: its dot product is fused into the epilogue of layer 1 and its backward into the
    operand producers of layer 1's dgrad / wgrad (mmsb_linear_*_head_tc) over ALL rows — one arithmetic for centre, taps
    and sampler; the geometry-feature path of the full rows is a second, additive gradient path (dgrad with accumulate).
    args: (x, n_full, group, act, act_param, W0, b0, W1, b1, W2, b2), two hidden layers."""

    @staticmethod
    def forward(ctx, x, n_full, group, act, act_param, w0, b0, w1, b1, w2, b2):
        prec = MLP_PRECISION
        if prec == 0:
            raise RuntimeError("SdfNetFn needs the tcgen05 layer path (MLP precision 1, 2 or 3)")
        in_dim = x.shape[-1]
        x2 = _rows(x, in_dim)
        n = x2.shape[0]
        if n_full > 0 and group > 1 and n_full * group != n:
            raise ValueError(f"SdfNetFn: {n} rows are not {n_full} groups of {group}")
        ws = [_f(w0), _f(w1), _f(w2)]
        bs = [_f(b0), _f(b1), _f(b2)]
        hid = ws[1].shape[0]
        g_dim = ws[2].shape[0] - 1
        need_grad = any(ctx.needs_input_grad)
        dev = x2.device
        # mode 2: fp16 split for the forward products it covers (every row of one call takes the same path: centre, taps
        # and sampler evaluations keep sharing one arithmetic), 3xTF32 for the backward products
        mode2 = prec == 2
        if sdf_fused_eligible(x2, ws[0], ws[1], act):
            # ONE kernel for layer 0 -> layer 1 -> sdf head (h0 stays on chip between the layers); fp16-split products
            # (fp32-accurate) in modes 2 / 3, a single fp16 pass in the fast mode 1.  h0 / h1 are written out only for the
            # backward kernels; without a backward only the centre rows' h1 is (the geometry features' input).
            h0 = torch.empty((n, hid), device=dev, dtype=torch.float32) if need_grad else None
            h1, h1_group = None, 1
            if need_grad or (n_full > 0 and group == 1):
                h1 = torch.empty((n, hid), device=dev, dtype=torch.float32)
            elif n_full > 0:
                h1, h1_group = torch.empty((n_full, hid), device=dev, dtype=torch.float32), group
            sdf = sdf_net_fwd_fused(x2, w0, bs[0], w1, bs[1], ws[2], bs[2], act, act_param, 1 if prec == 1 else 3, h0=h0, h1=h1,
                                    h1_group=h1_group)
            geo = None
            if n_full > 0:
                h1c = h1 if h1_group > 1 else (h1[0::group] if group > 1 else h1[:n_full])
                p2 = 1 if prec == 1 else 3
                geo = linear_fwd_tc(h1c, packed_weight(w2, False, p2, rows=(1, g_dim + 1)), bs[2][1:], g_dim, 0, 1.0, p2,
                                    out=row_slot(n_full, g_dim, dev))
        else:
            amaxes = torch.zeros((2,), device=dev, dtype=torch.float32) if mode2 else None
            p0 = layer_precision(x2, ws[0].shape[0]) if mode2 else prec
            h0 = linear_fwd_tc(x2, packed_weight(w0, False, p0), bs[0], ws[0].shape[0], act, act_param, p0,
                               y_amax=amaxes[0:1] if mode2 else None)
            sdf = torch.zeros((n,), device=dev, dtype=torch.float32)
            keep_h1 = need_grad or n_full > 0
            h1 = torch.empty((n, hid), device=dev, dtype=torch.float32) if keep_h1 else None
            p1 = layer_precision(h0, hid) if mode2 else prec
            call("mmsb_linear_fwd_head_tc", ptr(h0), _i64(h0.stride(0)), ptr(packed_weight(w1, False, p1)), ptr(bs[1]), ptr(h1),
                 _i64(hid), _i64(n), _i32(ws[1].shape[1]), _i32(hid), _i32(act), _f32(act_param), _i32(p1), ptr(ws[2]),
                 ptr(bs[2]), ptr(sdf), ptr(amaxes[0:1]) if p1 == 2 else None, ptr(amaxes[1:2]) if mode2 and keep_h1 else None,
                 stream_ptr())
            geo = None
            if n_full > 0:
                h1c = h1[0::group] if group > 1 else h1[:n_full]
                p2 = layer_precision(h1c, g_dim) if mode2 else prec
                geo = linear_fwd_tc(h1c, packed_weight(w2, False, p2, rows=(1, g_dim + 1)), bs[2][1:], g_dim, 0, 1.0, p2,
                                    out=row_slot(n_full, g_dim, dev), x_amax=amaxes[1:2] if p2 == 2 else None)
        if mode2:
            prec = 3
        ctx.cfg = (n, n_full, group, act, act_param, prec, in_dim, x.shape)
        if need_grad:
            ctx.save_for_backward(x2, h0, h1, *ws)
            ctx.packed_t = (packed_weight(w0, True, prec) if ctx.needs_input_grad[0] else None, packed_weight(w1, True, prec),
                            packed_weight(w2, True, prec, rows=(1, g_dim + 1)) if n_full > 0 else None)
        if geo is None:
            geo = torch.empty((0, g_dim), device=dev, dtype=torch.float32)
            ctx.mark_non_differentiable(geo)
        return sdf[:, None], geo

    @staticmethod
    def backward(ctx, dsdf, dgeo):
        n, n_full, group, act, act_param, prec, in_dim, x_shape = ctx.cfg
        x2, h0, h1, w0, w1, w2 = ctx.saved_tensors
        p0t, p1t, p2t = ctx.packed_t
        dev = x2.device
        hid, g_dim = w1.shape[0], w2.shape[0] - 1
        d = torch.zeros((n,), device=dev) if dsdf is None else _f(dsdf).reshape(n)
        dw0, db0, dw1, db1, dw2, db2 = _zeros_many([tuple(w0.shape), (w0.shape[0],), tuple(w1.shape), (hid,), tuple(w2.shape),
                                                    (w2.shape[0],)], dev)
        dz0 = torch.empty((n, hid), device=dev, dtype=torch.float32)
        # sdf-head path, every row: dz1 = d * w2[0] * act'(h1) is generated in the operand producers, never stored
        call("mmsb_linear_bwd_weight_head_tc", ptr(h1), _i64(hid), _i32(act), _f32(act_param), ptr(d), ptr(w2), ptr(h0),
             _i64(hid), ptr(dw1), ptr(db1), ptr(dw2), _i64(n), _i32(hid), _i32(hid), _i32(prec), stream_ptr())
        db2[0:1] += d.sum()
        call("mmsb_linear_bwd_data_head_tc", ptr(h1), _i64(hid), _i32(act), _f32(act_param), ptr(d), ptr(w2), ptr(p1t),
             ptr(dz0), _i64(hid), ptr(h0), _i64(hid), _i32(act), _f32(act_param), _i64(n), _i32(hid), _i32(hid),
             _i32(prec), stream_ptr())
        if n_full > 0:
            # geometry-feature path of the full rows (strided views when the rows are grouped), added on top
            dg = torch.zeros((n_full, g_dim), device=dev) if dgeo is None else _rows(dgeo.reshape(n_full, g_dim), g_dim)
            if group > 1:
                h1c, h0c, dz0c = h1[0::group], h0[0::group], dz0[0::group]
            else:
                h1c, h0c, dz0c = h1[:n_full], h0[:n_full], dz0[:n_full]
            linear_bwd_weight_tc(dg, h1c, dw2[1:], db2[1:], prec)
            dz1g = linear_bwd_data_tc(dg, p2t, hid, h1c, act, act_param, prec)
            linear_bwd_weight_tc(dz1g, h0c, dw1, db1, prec)
            linear_bwd_data_tc(dz1g, p1t, hid, h0c, act, act_param, prec, out=dz0c, accumulate=True)
        linear_bwd_weight_tc(dz0, x2, dw0, db0, prec)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = _padded_rows(n, in_dim, dev)
            linear_bwd_data_tc(dz0, p0t, in_dim, None, 0, 1.0, prec, out=dx)
            if tuple(dx.shape) != tuple(x_shape):
                dx = dx.reshape(x_shape)
        return dx, None, None, None, None, dw0, db0, dw1, db1, dw2, db2


SDF_FUSED = int(os.environ.get("MMSB_SDF_FUSED", "1"))


def sdf_fused_eligible(x2, w0, w1, act) -> bool:
    """Shapes mmsb_sdf_net_fwd_fused covers (include/mms_b200.h): 64 < in_dim <= 80, hidden width 256, ReLU / Softplus,
    16-byte aligned rows."""
    return (SDF_FUSED != 0 and MLP_PRECISION in (1, 2, 3) and 64 < x2.shape[1] <= 80 and tuple(w1.shape) == (256, 256)
            and w0.shape[0] == 256 and act in (ACT["ReLU"], ACT["Softplus"]) and x2.stride(1) == 1 and x2.stride(0) % 4 == 0
            and x2.data_ptr() % 16 == 0)


def sdf_net_fwd_fused(x2, w0, b0, w1, b1, head_w, head_b, act, act_param, products, h0=None, h1=None, h1_group=1, sdf=None):
    """One launch: sdf = head(act(W1 act(W0 x + b0) + b1)); h0 / h1 are stored when given (see include/mms_b200.h)."""
    n = x2.shape[0]
    if sdf is None:
        sdf = torch.empty((n,), device=x2.device, dtype=torch.float32)
    w0c = _f(w0.detach())
    call("mmsb_sdf_net_fwd_fused", ptr(x2), _i64(x2.stride(0)), _i64(n), _i32(x2.shape[1]), _i32(w1.shape[0]), ptr(w0c),
         ptr(packed_weight(w0, False, 2)), ptr(b0), ptr(packed_weight(w1, False, 2)), ptr(b1), ptr(head_w), ptr(head_b),
         _i32(act), _f32(act_param), _i32(products), ptr(h0), _i64(h0.stride(0) if h0 is not None else 0), ptr(h1),
         _i64(h1.stride(0) if h1 is not None else 0), _i32(h1_group), ptr(sdf), stream_ptr())
    return sdf


def sdf_net_forward(x, n_full, weights, biases, act: str, act_param: float, group: int = 1):
    return SdfNetFn.apply(x, int(n_full), int(group), ACT[act], float(act_param), weights[0], biases[0], weights[1], biases[1],
                          weights[2], biases[2])


def mlp_forward(x, weights: Sequence[torch.Tensor], biases: Sequence[Optional[torch.Tensor]], hidden_act: str,
                out_act: Optional[str], act_param: float = 1.0, skips=(), n_out_used=None):
    params = []
    for w, b in zip(weights, biases):
        params += [w, b]
    # an output activation whose derivative depends on the output only (ReLU) can be folded into the consumers' dgrads:
    # the token travels on the output tensor to ops.split_rows
    token = _OutActToken(ACT[out_act], float(act_param)) if out_act == "ReLU" else None
    out = MLPFn.apply(x, ACT[hidden_act], float(act_param), ACT[out_act], tuple(skips), n_out_used, token, *params)
    if token is not None and out.dim() == 2:
        out._mmsb_out_act = token
    return out


# ------------------------------------------------------------------------------------------------
# polarization head post-processing (A17)
# ------------------------------------------------------------------------------------------------
class PolarizationFn(torch.autograd.Function):
    """stokes [n,3] (raw head output), directions [n,3], up_directions [n,3] -> intensities [n,4]
    (field_heads.py:101-105: leaky_relu(S0), align_polarization_filters, stokes_to_intensity)."""

    @staticmethod
    def forward(ctx, stokes, directions, up_directions):
        s2 = _rows(stokes, 3)
        d2, u2 = _f(directions).reshape(-1, 3), _f(up_directions).reshape(-1, 3)
        n = s2.shape[0]
        out = torch.empty((n, 4), device=s2.device, dtype=torch.float32)
        call("mmsb_polarization_fwd", ptr(s2), _i64(s2.stride(0)), ptr(d2), ptr(u2), ptr(out), _i64(n), stream_ptr())
        ctx.save_for_backward(s2, d2, u2)
        ctx.shapes = (stokes.shape, directions.shape, up_directions.shape)
        return out

    @staticmethod
    def backward(ctx, d_out):
        s2, d2, u2 = ctx.saved_tensors
        n = s2.shape[0]
        g = _f(d_out).reshape(n, 4)
        ds = torch.empty((n, 3), device=s2.device, dtype=torch.float32)
        dd = torch.empty((n, 3), device=s2.device, dtype=torch.float32) if ctx.needs_input_grad[1] else None
        du = torch.empty((n, 3), device=s2.device, dtype=torch.float32) if ctx.needs_input_grad[2] else None
        call("mmsb_polarization_bwd", ptr(s2), _i64(s2.stride(0)), ptr(d2), ptr(u2), ptr(g), ptr(ds), ptr(dd), ptr(du), _i64(n),
             stream_ptr())
        sh = ctx.shapes
        return ds.reshape(sh[0]), (dd.reshape(sh[1]) if dd is not None else None), (du.reshape(sh[2]) if du is not None else None)


# ------------------------------------------------------------------------------------------------
# ray generation / collider (A1-A3)
# ------------------------------------------------------------------------------------------------
class RayGenFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, coords, c2w, intr, dist, pose_adjust, pixel_offset):
        n = coords.shape[0]
        dev = coords.device
        coords = coords.to(torch.int32).contiguous()
        c2w, intr = _f(c2w), _f(intr)
        dist = _f(dist) if dist is not None else None
        pa = _f(pose_adjust) if pose_adjust is not None else None
        n_cam = c2w.shape[0]
        n_pose = pa.shape[0] if pa is not None else 0
        o = torch.empty((n, 3), device=dev)
        d = torch.empty((n, 3), device=dev)
        up = torch.empty((n, 3), device=dev)
        area = torch.empty((n, 1), device=dev)
        dn = torch.empty((n, 1), device=dev)
        call("mmsb_raygen_fwd", ptr(coords), ptr(c2w), ptr(intr), ptr(dist), ptr(pa), _i32(n_pose), _i32(n_cam),
             _f32(pixel_offset), ptr(o), ptr(d), ptr(up), ptr(area), ptr(dn), _i64(n), stream_ptr())
        ctx.save_for_backward(coords, c2w, intr, dist, pa)
        ctx.cfg = (n_pose, n_cam, float(pixel_offset))
        ctx.mark_non_differentiable(area, dn)
        return o, d, up, area, dn

    @staticmethod
    def backward(ctx, do, dd, dup, _da, _dn):
        coords, c2w, intr, dist, pa = ctx.saved_tensors
        n_pose, n_cam, off = ctx.cfg
        if pa is None or not ctx.needs_input_grad[4]:
            return None, None, None, None, None, None
        dp = torch.zeros_like(pa)
        do = _f(do) if do is not None else None
        dd = _f(dd) if dd is not None else None
        dup = _f(dup) if dup is not None else None
        call("mmsb_raygen_bwd", ptr(coords), ptr(c2w), ptr(intr), ptr(dist), ptr(pa), _i32(n_pose), _i32(n_cam),
             _f32(off), ptr(do), ptr(dd), ptr(dup), ptr(dp), _i64(coords.shape[0]), stream_ptr())
        return None, None, None, None, dp, None


def sphere_collide(origins, directions, radius=1.0, background=False):
    """-> nears [n,1], fars [n,1], mask bool [n] (+ bg_nears, bg_fars when background=True). No grad."""
    o, d = _rows(origins.detach(), 3), _rows(directions.detach(), 3)
    n = o.shape[0]
    nears = torch.empty((n, 1), device=o.device)
    fars = torch.empty((n, 1), device=o.device)
    mask = torch.empty((n,), device=o.device, dtype=torch.uint8)
    bgn = torch.empty((n, 1), device=o.device) if background else None
    bgf = torch.empty((n, 1), device=o.device) if background else None
    call("mmsb_sphere_collide", ptr(o), ptr(d), _f32(radius), ptr(nears), ptr(fars), ptr(mask), ptr(bgn), ptr(bgf),
         _i64(n), stream_ptr())
    if background:
        return nears, fars, mask, bgn, bgf
    return nears, fars, mask


# ------------------------------------------------------------------------------------------------
# samplers (A4-A7), all no-grad
# ------------------------------------------------------------------------------------------------
_LIN_CACHE = {}


def _linspace01(n_edges, device):
    key = (n_edges, str(device))
    if key not in _LIN_CACHE:
        _LIN_CACHE[key] = torch.linspace(0.0, 1.0, n_edges).to(device)   # computed by torch on the CPU: same table as the reference
    return _LIN_CACHE[key]


def spaced_bins(nears, fars, num_samples, spacing, t_rand=None):
    """-> spacing bins [n, S+1], euclidean bins [n, S+1]"""
    nears, fars = _f(nears.detach()).reshape(-1), _f(fars.detach()).reshape(-1)
    n = nears.shape[0]
    lin = _linspace01(num_samples + 1, nears.device)
    sb = torch.empty((n, num_samples + 1), device=nears.device)
    eb = torch.empty_like(sb)
    rpr = 0
    if t_rand is not None:
        t_rand = _f(t_rand)
        rpr = t_rand.shape[-1]
    call("mmsb_spaced_bins", ptr(nears), ptr(fars), ptr(lin), ptr(t_rand), _i32(rpr), _i32(num_samples), _i32(spacing),
         ptr(sb), ptr(eb), _i64(n), stream_ptr())
    return sb, eb


def neus_upsample(bins, sdf, u, nears, fars, inv_s, histogram_padding=1e-5, eps=1e-5, want_debug=False):
    bins, sdf, u = _f(bins), _f(sdf), _f(u)
    nears, fars = _f(nears).reshape(-1), _f(fars).reshape(-1)
    n, m = sdf.shape
    k = u.shape[1] - 1
    dev = bins.device
    cdf = torch.empty((n, m + 1), device=dev)
    inds = torch.empty((n, k + 1), device=dev, dtype=torch.int64) if want_debug else None
    new_bins = torch.empty((n, k + 1), device=dev)
    merged = torch.empty((n, m + k + 1), device=dev)
    index = torch.empty((n, m + k), device=dev, dtype=torch.int64)
    call("mmsb_neus_upsample", ptr(bins), ptr(sdf), ptr(u), ptr(nears), ptr(fars), _f32(inv_s), _f32(histogram_padding),
         _f32(eps), _i32(m), _i32(k), ptr(cdf), ptr(inds), ptr(new_bins), ptr(merged), ptr(index), _i64(n), stream_ptr())
    if want_debug:
        return new_bins, merged, index, cdf, inds
    return new_bins, merged, index


def merge_rows(a, b, index):
    a, b = _f(a), _f(b)
    n, m = a.shape
    k = b.shape[1]
    out = torch.empty((n, m + k), device=a.device)
    call("mmsb_merge_rows", ptr(a), _i32(m), ptr(b), _i32(k), ptr(index.contiguous()), ptr(out), _i64(n), stream_ptr())
    return out


def searchsorted_right(cdf, u):
    cdf, u = _f(cdf), _f(u)
    n, m = cdf.shape
    q = u.shape[1]
    inds = torch.empty((n, q), device=cdf.device, dtype=torch.int64)
    call("mmsb_searchsorted_right", ptr(cdf), ptr(u), ptr(inds), _i32(m), _i32(q), _i64(n), stream_ptr())
    return inds


def pdf_inverse(cdf, bins, u):
    cdf, bins, u = _f(cdf), _f(bins), _f(u)
    n, nb = cdf.shape
    q = u.shape[1]
    inds = torch.empty((n, q), device=cdf.device, dtype=torch.int64)
    out = torch.empty((n, q), device=cdf.device)
    call("mmsb_pdf_inverse", ptr(cdf), ptr(bins), ptr(u), _i32(nb), _i32(q), ptr(inds), ptr(out), _i64(n), stream_ptr())
    return inds, out


# ------------------------------------------------------------------------------------------------
# weights / compositing (A13, A14, A18, A19)
# ------------------------------------------------------------------------------------------------
class SdfTapsFn(torch.autograd.Function):
    """(sdf_c [n], sdf_t [4,n]) -> gradients [n,3], hessians [n,3] or None, normals [n,3]"""

    @staticmethod
    def forward(ctx, sdf_c, sdf_t, delta, want_hessian):
        sdf_c, sdf_t = _f(sdf_c).reshape(-1), _f(sdf_t).reshape(4, -1)
        n = sdf_c.shape[0]
        four_delta = float(np.float32(4.0 * delta))
        delta_sq = float(np.float32(delta ** 2))
        g = torch.empty((n, 3), device=sdf_c.device)
        h = torch.empty((n, 3), device=sdf_c.device) if want_hessian else None
        nrm = torch.empty((n, 3), device=sdf_c.device)
        call("mmsb_sdf_taps_fwd", ptr(sdf_c), ptr(sdf_t), _f32(four_delta), _f32(delta_sq), ptr(g), ptr(h), ptr(nrm),
             _i64(n), stream_ptr())
        ctx.save_for_backward(sdf_t)
        ctx.cfg = (four_delta, delta_sq, want_hessian)
        if not want_hessian:
            h = torch.zeros((0,), device=sdf_c.device)
            ctx.mark_non_differentiable(h)
        return g, h, nrm

    @staticmethod
    def backward(ctx, dg, dh, dn):
        (sdf_t,) = ctx.saved_tensors
        four_delta, delta_sq, want_hessian = ctx.cfg
        n = sdf_t.shape[1]
        d_c = torch.empty((n,), device=sdf_t.device)
        d_t = torch.empty((4, n), device=sdf_t.device)
        dg = _f(dg) if dg is not None else None
        dh = _f(dh) if (dh is not None and want_hessian) else None
        dn = _f(dn) if dn is not None else None
        call("mmsb_sdf_taps_bwd", ptr(sdf_t), _f32(four_delta), _f32(delta_sq), ptr(dg), ptr(dh), ptr(dn), ptr(d_c),
             ptr(d_t), _i64(n), stream_ptr())
        return d_c, d_t, None, None


class NeusWeightsFn(torch.autograd.Function):
    """weights [n,s] from sdf [n,s], gradients [n,s,3], dirs [n,3], deltas [n,s], inv_s [1]."""

    @staticmethod
    def forward(ctx, sdf, grad, dirs, deltas, inv_s, mask, anneal):
        n, s = sdf.shape[0], sdf.shape[1]
        sdf, grad, dirs, deltas = _f(sdf).reshape(n, s), _f(grad).reshape(n, s, 3), _f(dirs).reshape(n, 3), _f(deltas).reshape(n, s)
        inv_s = _f(inv_s).reshape(1)
        w = torch.empty((n, s), device=sdf.device)
        call("mmsb_neus_weights_fwd", ptr(sdf), ptr(grad), ptr(dirs), ptr(deltas), ptr(inv_s), ptr(mask), _f32(anneal),
             ptr(w), _i32(s), _i64(n), stream_ptr())
        ctx.save_for_backward(sdf, grad, dirs, deltas, inv_s, mask)
        ctx.anneal = float(anneal)
        return w

    @staticmethod
    def backward(ctx, dw):
        sdf, grad, dirs, deltas, inv_s, mask = ctx.saved_tensors
        n, s = sdf.shape
        dw = _f(dw).reshape(n, s)
        d_sdf = torch.empty_like(sdf)
        d_grad = torch.empty_like(grad)
        d_dirs = torch.empty_like(dirs) if ctx.needs_input_grad[2] else None
        d_deltas = torch.empty_like(deltas) if ctx.needs_input_grad[3] else None
        d_inv_s = torch.zeros_like(inv_s)
        call("mmsb_neus_weights_bwd", ptr(sdf), ptr(grad), ptr(dirs), ptr(deltas), ptr(inv_s), ptr(mask),
             _f32(ctx.anneal), ptr(dw), ptr(d_sdf), ptr(d_grad), ptr(d_dirs), ptr(d_deltas), ptr(d_inv_s), _i32(s),
             _i64(n), stream_ptr())
        return d_sdf, d_grad, d_dirs, d_deltas, d_inv_s, None, None


class DensityWeightsFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, density, deltas):
        n, s = density.shape[0], density.shape[1]
        density, deltas = _f(density).reshape(n, s), _f(deltas).reshape(n, s)
        w = torch.empty((n, s), device=density.device)
        call("mmsb_density_weights_fwd", ptr(density), ptr(deltas), ptr(w), _i32(s), _i64(n), stream_ptr())
        ctx.save_for_backward(density, deltas)
        return w

    @staticmethod
    def backward(ctx, dw):
        density, deltas = ctx.saved_tensors
        n, s = density.shape
        dd = torch.empty_like(density)
        ddel = torch.empty_like(deltas) if ctx.needs_input_grad[1] else None
        call("mmsb_density_weights_bwd", ptr(density), ptr(deltas), ptr(_f(dw).reshape(n, s)), ptr(dd), ptr(ddel),
             _i32(s), _i64(n), stream_ptr())
        return dd, ddel


class CompositeFn(torch.autograd.Function):
    """color [n,c] = sum_s w v + bg (1 - sum_s w)"""

    @staticmethod
    def forward(ctx, weights, values, background):
        n, s = weights.shape[0], weights.shape[1]
        c = values.shape[-1]
        weights, values = _f(weights).reshape(n, s), _f(values).reshape(n, s, c)
        background = _f(background).reshape(n, c) if background is not None else None
        out = torch.empty((n, c), device=weights.device)
        call("mmsb_composite_fwd", ptr(weights), ptr(values), ptr(background), _i32(c), None, None, None, ptr(out),
             None, None, None, _i32(s), _i64(n), stream_ptr())
        ctx.save_for_backward(weights, values, background)
        return out

    @staticmethod
    def backward(ctx, dc):
        weights, values, background = ctx.saved_tensors
        n, s = weights.shape
        c = values.shape[-1]
        dw = torch.empty_like(weights)
        dv = torch.empty_like(values) if ctx.needs_input_grad[1] else None
        dbg = torch.empty_like(background) if (background is not None and ctx.needs_input_grad[2]) else None
        call("mmsb_composite_bwd", ptr(weights), ptr(values), ptr(background), _i32(c), ptr(_f(dc).reshape(n, c)), None,
             None, None, None, None, None, ptr(dw), ptr(dv), ptr(dbg), None, _i32(s), _i64(n), stream_ptr())
        return dw, dv, dbg


def composite_aux(weights, normals, starts, ends):
    """No-grad diagnostics: rendered normals [n,3], unclipped depth [n,1], accumulation [n,1]."""
    n, s = weights.shape[0], weights.shape[1]
    weights = _f(weights.detach()).reshape(n, s)
    normals = _f(normals.detach()).reshape(n, s, 3)
    starts, ends = _f(starts.detach()).reshape(n, s), _f(ends.detach()).reshape(n, s)
    on = torch.empty((n, 3), device=weights.device)
    od = torch.empty((n, 1), device=weights.device)
    oa = torch.empty((n, 1), device=weights.device)
    call("mmsb_composite_fwd", ptr(weights), None, None, _i32(0), ptr(normals), ptr(starts), ptr(ends), None, ptr(on),
         ptr(od), ptr(oa), _i32(s), _i64(n), stream_ptr())
    return on, od, oa


# ------------------------------------------------------------------------------------------------
# losses (A21, A22)
# ------------------------------------------------------------------------------------------------
def first_saturated(target, threshold):
    t = _f(target.detach()).reshape(-1)
    idx = torch.empty((1,), device=t.device, dtype=torch.int64)
    call("mmsb_first_saturated", ptr(t), _f32(threshold), ptr(idx), _i64(t.numel()), stream_ptr())
    return idx


class MosaickL1Fn(torch.autograd.Function):
    """mean |select(rendered) - target| with the mosaick band select fused (pattern may be None)."""

    @staticmethod
    def forward(ctx, rendered, target, coords, pattern, ph, pw, sat_threshold, sat_index):
        n, c = rendered.shape
        rendered = _f(rendered)
        target = _f(target.detach())
        coords = coords.to(torch.int32).contiguous() if coords is not None else None
        count = n if pattern is not None else n * c
        loss = torch.zeros((1,), device=rendered.device)
        sel = torch.empty((n,), device=rendered.device) if pattern is not None else None
        call("mmsb_mosaick_l1_fwd", ptr(coords), ptr(pattern), _i32(ph), _i32(pw), ptr(rendered), _i32(c), ptr(target),
             _f32(sat_threshold), ptr(sat_index), None, ptr(sel), ptr(loss), _i64(n), stream_ptr())
        ctx.save_for_backward(rendered, target, coords, pattern, sat_index)
        ctx.cfg = (ph, pw, float(sat_threshold), 1.0 / max(count, 1))
        out = (loss / max(count, 1)).reshape(())
        if sel is None:
            sel = torch.zeros((0,), device=rendered.device)
        ctx.mark_non_differentiable(sel)
        return out, sel

    @staticmethod
    def backward(ctx, dloss, _dsel):
        rendered, target, coords, pattern, sat_index = ctx.saved_tensors
        ph, pw, thr, inv_count = ctx.cfg
        n, c = rendered.shape
        dr = torch.empty_like(rendered)
        dl = _f(dloss).reshape(1)
        call("mmsb_mosaick_l1_bwd", ptr(coords), ptr(pattern), _i32(ph), _i32(pw), ptr(rendered), _i32(c), ptr(target),
             _f32(thr), ptr(sat_index), ptr(dl), _f32(inv_count), ptr(dr), _i64(n), stream_ptr())
        return dr, None, None, None, None, None, None, None


def mosaick_bands(coords, pattern, ph, pw, rendered):
    """band int64 [n] and the gathered channel [n] (index-parity surface; A21)."""
    n, c = rendered.shape
    rendered = _f(rendered.detach())
    coords = coords.to(torch.int32).contiguous()
    band = torch.empty((n,), device=rendered.device, dtype=torch.int64)
    sel = torch.empty((n,), device=rendered.device)
    loss = torch.zeros((1,), device=rendered.device)
    tgt = torch.zeros((n,), device=rendered.device)
    call("mmsb_mosaick_l1_fwd", ptr(coords), ptr(pattern), _i32(ph), _i32(pw), ptr(rendered), _i32(c), ptr(tgt),
         _f32(float("inf")), None, ptr(band), ptr(sel), ptr(loss), _i64(n), stream_ptr())
    return band, sel


class GeometryLossFn(torch.autograd.Function):
    """(eikonal, curvature) means over the unmasked samples of gradients/hessians [n,s,3].  `count` (device fp32 [1],
    optional): divide the sums by this number instead of the batch's own unmasked-sample count — the GLOBAL count when
    the batch is one shard of a step (pipelines.ShardPlan), so that the shards' losses and gradients add up to the
    unsharded ones."""

    @staticmethod
    def forward(ctx, gradients, hessians, ray_mask, count=None):
        nr, s = gradients.shape[0], gradients.shape[1]
        g = _f(gradients).reshape(nr * s, 3)
        h = _f(hessians).reshape(nr * s, 3) if hessians is not None else None
        sums = torch.zeros((3,), device=g.device)
        call("mmsb_geometry_loss_fwd", ptr(g), ptr(h), ptr(ray_mask), _i32(s), ptr(sums), _i64(nr * s), stream_ptr())
        if count is not None:
            sums = torch.cat([sums[:2], _f(count.detach()).reshape(1)])
        ctx.save_for_backward(g, h, ray_mask, sums)
        ctx.cfg = (s, gradients.shape)
        cnt = sums[2].clamp_min(1.0)
        return sums[0] / cnt, sums[1] / cnt

    @staticmethod
    def backward(ctx, d_eik, d_curv):
        g, h, ray_mask, sums = ctx.saved_tensors
        s, shape = ctx.cfg
        dg = torch.empty_like(g)
        dh = torch.empty_like(h) if h is not None else None
        de = _f(d_eik).reshape(1) if d_eik is not None else None
        dc = _f(d_curv).reshape(1) if d_curv is not None else None
        call("mmsb_geometry_loss_bwd", ptr(g), ptr(h), ptr(ray_mask), _i32(s), ptr(sums), ptr(de), ptr(dc), ptr(dg),
             ptr(dh), _i64(g.shape[0]), stream_ptr())
        return dg.reshape(shape), (dh.reshape(shape) if dh is not None else None), None, None


# ------------------------------------------------------------------------------------------------
# optimiser (A23, "next" row 1)
# ------------------------------------------------------------------------------------------------
def sumsq(x, out):
    call("mmsb_sumsq", ptr(x), ptr(out), _i64(x.numel()), stream_ptr())


def adamw_step(param, grad, exp_avg, exp_avg_sq, grad_scale, lr, beta1, beta2, eps, weight_decay, step):
    call("mmsb_adamw_step", ptr(param), ptr(grad), ptr(exp_avg), ptr(exp_avg_sq), ptr(grad_scale), _f32(lr), _f32(beta1),
         _f32(beta2), _f32(eps), _f32(weight_decay), _i32(step), _i64(param.numel()), stream_ptr())


def adamw_step_dev(param, grad, exp_avg, exp_avg_sq, grad_sumsq, max_norm, hyper, beta1, beta2, eps, weight_decay):
    """hyper: device tensor [4] = {lr, 1 - beta1^t, sqrt(1 - beta2^t), gradient prescale} (graph-replayable AdamW)."""
    call("mmsb_adamw_step_dev", ptr(param), ptr(grad), ptr(exp_avg), ptr(exp_avg_sq), ptr(grad_sumsq), _f32(max_norm or 0.0),
         ptr(hyper), _f32(beta1), _f32(beta2), _f32(eps), _f32(weight_decay), _i64(param.numel()), stream_ptr())


def sample_pixels(seed: int, step: int, stream_id: int, n_cam: int, height: int, width: int, n: int, device, frames=None):
    """(camera, y, x) int32 [n,3] drawn on the device (+ targets [n,C] gathered from frames [n_cam,H,W,C] when given)."""
    coords = torch.empty((n, 3), device=device, dtype=torch.int32)
    targets, channels = None, 0
    if frames is not None:
        if not frames.is_cuda or frames.dtype != torch.float32 or not frames.is_contiguous() or frames.dim() != 4:
            raise ValueError("frames must be a contiguous fp32 CUDA tensor [n_cam, H, W, C]")
        if tuple(frames.shape[:3]) != (n_cam, height, width):
            raise ValueError(f"frames {tuple(frames.shape)} do not match n_cam={n_cam} H={height} W={width}")
        channels = frames.shape[3]
        targets = torch.empty((n, channels), device=device, dtype=torch.float32)
    call("mmsb_sample_pixels", ctypes.c_uint64(seed & 0xFFFFFFFFFFFFFFFF), _i32(step), _i32(stream_id), _i32(n_cam), _i32(height),
         _i32(width), ptr(frames), _i32(channels), ptr(coords), ptr(targets), _i64(n), stream_ptr())
    return coords, targets
