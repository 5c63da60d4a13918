"""Host-side mirror of the reference's codes/style_transformer.py.

Same class names, constructor arguments, forward signatures and state_dict layout as the
reference, so checkpoints and the callers train.py / test_model.py work
unchanged; the arithmetic runs in the sm_100a kernels behind the C ABI (engine.py).  The
parameters stay ordinary nn.Linear / nn.LayerNorm sub-modules so deepcopy, .apply(init_fn),
optimizers and load_state_dict behave as in the reference.

The reference's default configuration is built in CUDA for inference and training; of the alternates SURVEY.md section 8f
lists, the three that are re-orderings of the same ops (unprocessed Key for Scale/Shift, Key InstanceNorm before Wk, no MLP
after the decoder's self-attention) run in inference; anything else raises NotImplementedError at call time -- never a
silent fallback.
"""
from __future__ import annotations

import weakref
from typing import Callable, List, Optional

import torch
from torch import Tensor, nn
from torchvision.ops.misc import MLP
from torchvision.ops.stochastic_depth import StochasticDepth

from . import engine

_cache: "weakref.WeakKeyDictionary[nn.Module, tuple]" = weakref.WeakKeyDictionary()
_workspaces: "weakref.WeakKeyDictionary[nn.Module, engine.Workspace]" = weakref.WeakKeyDictionary()


def params_key(module: nn.Module):
    return tuple((p.data_ptr(), p._version) for p in module.parameters())


def packed_weights(module: nn.Module, builder):
    """Pack the module's parameters for the kernels once per parameter version (one cache slot per builder:
    the inference and the training engines keep different packings of the same module)."""
    key = params_key(module)
    slots = _cache.get(module)
    if slots is None:
        slots = {}
        _cache[module] = slots
    hit = slots.get(builder)
    if hit is None or hit[0] != key:
        with torch.no_grad():
            hit = (key, builder({k: v for k, v in module.state_dict().items()}))
        slots[builder] = hit
    return hit[1]


def pin_state(modules) -> list:
    """References to everything a CUDA graph captured over `modules` has baked in by address and that this file's caches
    could otherwise drop later: the current packed weights (replaced when a parameter's version changes, e.g. a reload of
    the frozen Swin / VGG weights) and every workspace buffer (replaced when a larger shape arrives; engine.Workspace keeps
    superseded buffers itself, this also covers a workspace replaced as a whole).  The Graphed* objects hold the returned
    list for their lifetime, so a replay never writes to memory the caching allocator has handed to someone else."""
    keep = []
    for root in modules:
        for m in root.modules():
            if m in _cache:
                keep.append(dict(_cache[m]))
            ws = _workspaces.get(m)
            if ws is not None:
                keep.append(ws)
                keep.append(list(ws.bufs.values()))
    return keep


def workspace_of(module: nn.Module, device) -> engine.Workspace:
    ws = _workspaces.get(module)
    if ws is None or ws.device != device:
        ws = engine.Workspace(device)
        _workspaces[module] = ws
    return ws


def require_cuda(*tensors: Tensor) -> None:
    for t in tensors:
        if not t.is_cuda:
            raise RuntimeError("mastermetastyletransfer_b200 runs on sm_100a only: inputs must be CUDA tensors (no CPU fallback)")


def wants_grad(module: nn.Module, *tensors: Tensor) -> bool:
    """True when this call must go through the training engine (autograd is recording and something upstream or a
    parameter of this module requires a gradient)."""
    if not torch.is_grad_enabled():
        return False
    return any(p.requires_grad for p in module.parameters()) or any(t.requires_grad for t in tensors)


def require_inference(module: nn.Module, *tensors: Tensor) -> None:
    require_cuda(*tensors)
    if wants_grad(module, *tensors):
        raise NotImplementedError(
            "this module has no backward kernels (the Swin encoder is frozen in the reference's training setup, train.py:216-218): "
            "freeze its parameters or call under torch.no_grad()")


def _relative_position_index(ws: List[int]) -> Tensor:
    """idx[i,j] = (yi-yj+Wh-1)*(2Ww-1) + (xi-xj+Ww-1), flattened (reference :227-239)."""
    ys, xs = torch.meshgrid(torch.arange(ws[0]), torch.arange(ws[1]), indexing="ij")
    ys, xs = ys.reshape(-1), xs.reshape(-1)
    dy = ys[:, None] - ys[None, :] + ws[0] - 1
    dx = xs[:, None] - xs[None, :] + ws[1] - 1
    return (dy * (2 * ws[1] - 1) + dx).reshape(-1)


class _WindowAttentionBase(nn.Module):
    def _common(self, dim, num_heads, window_size, shift_size, dropout, attention_dropout):
        if len(window_size) != 2 or len(shift_size) != 2:
            raise ValueError("window_size and shift_size must be of length 2")
        self.window_size, self.shift_size, self.num_heads = window_size, shift_size, num_heads
        self.attention_dropout, self.dropout = attention_dropout, dropout

    def _tables(self):
        n = (2 * self.window_size[0] - 1) * (2 * self.window_size[1] - 1)
        self.relative_position_bias_table = nn.Parameter(torch.zeros(n, self.num_heads))
        nn.init.trunc_normal_(self.relative_position_bias_table, std=0.02)
        self.register_buffer("relative_position_index", _relative_position_index(self.window_size))

    def get_relative_position_bias(self) -> Tensor:
        n = self.window_size[0] * self.window_size[1]
        return self.relative_position_bias_table[self.relative_position_index].view(n, n, -1).permute(2, 0, 1).contiguous().unsqueeze(0)


class ShiftedWindowAttention(_WindowAttentionBase):
    """Mirror of reference :175-292 (split Wq/Wk/Wv so q, k, v may come from different tensors)."""

    def __init__(self, dim: int, num_heads: int, window_size: List[int], shift_size: List[int], dropout: float = 0.0,
                 attention_dropout: float = 0.0, qkv_bias: bool = True, proj_bias: bool = True):
        super().__init__()
        self._common(dim, num_heads, window_size, shift_size, dropout, attention_dropout)
        self.Wq = nn.Linear(dim, dim, bias=qkv_bias)
        self.Wk = nn.Linear(dim, dim, bias=qkv_bias)
        self.Wv = nn.Linear(dim, dim, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim, bias=proj_bias)
        self._tables()

    def forward(self, input_q: Tensor, input_k: Tensor, input_v: Tensor):
        raise NotImplementedError("call through StyleTransformer.forward: the B200 path fuses whole layers (DESIGN.md)")


class ShiftedWindowAttention_for_decoder_last_MHA(_WindowAttentionBase):
    """Mirror of reference :616-764 (no Wq by default; Wk, Wv_scale, Wv_shift and one shared proj)."""

    def __init__(self, dim: int, num_heads: int, window_size: List[int], shift_size: List[int], instance_norm_q: nn.Module,
                 instance_norm_k: nn.Module, dropout: float = 0.0, attention_dropout: float = 0.0, qkv_bias: bool = True,
                 proj_bias: bool = True, use_q_proj: bool = False,
                 use_Key_instance_norm_after_linear_transformation: bool = True):
        super().__init__()
        self._common(dim, num_heads, window_size, shift_size, dropout, attention_dropout)
        self.instance_norm_q, self.instance_norm_k = instance_norm_q, instance_norm_k
        self.use_q_proj = use_q_proj
        self.use_Key_instance_norm_after_linear_transformation = use_Key_instance_norm_after_linear_transformation
        if use_q_proj:
            self.Wq = nn.Linear(dim, dim, bias=qkv_bias)
        self.Wk = nn.Linear(dim, dim, bias=qkv_bias)
        self.Wv_scale = nn.Linear(dim, dim, bias=qkv_bias)
        self.Wv_shift = nn.Linear(dim, dim, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim, bias=proj_bias)
        self._tables()

    def forward(self, input_q, input_k, input_v_scale, input_v_shift):
        raise NotImplementedError("call through StyleTransformer.forward: the B200 path fuses whole layers (DESIGN.md)")


def _xavier_mlp(mlp: nn.Module) -> None:
    for m in mlp.modules():
        if isinstance(m, nn.Linear):
            nn.init.xavier_uniform_(m.weight)
            if m.bias is not None:
                nn.init.normal_(m.bias, std=1e-6)


class StyleSwinTransformerBlock(nn.Module):
    """Mirror of reference :303-398."""

    def __init__(self, dim: int, num_heads: int, window_size: List[int] = [8, 8], shift_size: List[int] = [4, 4],
                 dropout: float = 0.0, attention_dropout: float = 0.0, qkv_bias: bool = True, proj_bias: bool = True,
                 mlp_ratio: float = 4.0, stochastic_depth_prob: float = 0.0,
                 norm_layer: Callable[..., nn.Module] = nn.LayerNorm, MLP_activation_layer: Optional[nn.Module] = nn.GELU,
                 exclude_MLP_after: bool = False):
        super().__init__()
        self.exclude_MLP_after = exclude_MLP_after
        self.use_norm = norm_layer is not None
        if self.use_norm:
            self.norm1 = norm_layer(dim)
            if not exclude_MLP_after:
                self.norm2 = norm_layer(dim)
        self.attn = ShiftedWindowAttention(dim=dim, num_heads=num_heads, window_size=window_size, shift_size=shift_size,
                                           dropout=dropout, attention_dropout=attention_dropout, qkv_bias=qkv_bias,
                                           proj_bias=proj_bias)
        self.stochastic_depth = StochasticDepth(stochastic_depth_prob, "row")
        if not exclude_MLP_after:
            self.mlp = MLP(dim, [int(dim * mlp_ratio), dim], activation_layer=MLP_activation_layer, inplace=None, dropout=dropout)
            _xavier_mlp(self.mlp)

    def forward(self, input_q, input_k, input_v, calculating_Key_in_encoder: bool = None):
        raise NotImplementedError("call through StyleTransformer.forward: the B200 path fuses whole layers (DESIGN.md)")


class StyleEncoder(nn.Module):
    """Mirror of reference :777-912: one shared MHA block without MLP, three private MLPs."""

    def __init__(self, encoder_dim: int, encoder_num_heads: int, encoder_window_size: List[int], encoder_shift_size: List[int],
                 encoder_mlp_ratio: float = 4.0, encoder_dropout: float = 0.0, encoder_attention_dropout: float = 0.0,
                 encoder_qkv_bias: bool = True, encoder_proj_bias: bool = True, encoder_stochastic_depth_prob: float = 0.1,
                 encoder_norm_layer: Callable[..., nn.Module] = None, encoder_MLP_activation_layer: Optional[nn.Module] = nn.GELU,
                 encoder_if_use_processed_Key_in_Scale_and_Shift_calculation: bool = True):
        super().__init__()
        self.if_use_processed_Key_in_Scale_and_Shift_calculation = encoder_if_use_processed_Key_in_Scale_and_Shift_calculation
        self.encoder_stochastic_depth_prob = encoder_stochastic_depth_prob
        self.stochastic_depth = StochasticDepth(encoder_stochastic_depth_prob, "row")
        self.shared_MHA_without_MLP = StyleSwinTransformerBlock(
            dim=encoder_dim, num_heads=encoder_num_heads, window_size=encoder_window_size, shift_size=encoder_shift_size,
            dropout=encoder_dropout, attention_dropout=encoder_attention_dropout, qkv_bias=encoder_qkv_bias,
            proj_bias=encoder_proj_bias, mlp_ratio=encoder_mlp_ratio, stochastic_depth_prob=encoder_stochastic_depth_prob,
            norm_layer=encoder_norm_layer, MLP_activation_layer=encoder_MLP_activation_layer, exclude_MLP_after=True)
        hidden = [int(encoder_dim * encoder_mlp_ratio), encoder_dim]
        mk = lambda: MLP(encoder_dim, hidden, activation_layer=encoder_MLP_activation_layer, inplace=None, dropout=encoder_dropout)
        self.encoder_MLP_Key, self.encoder_MLP_Scale, self.encoder_MLP_Shift = mk(), mk(), mk()
        # (the reference's xavier loop over these three MLPs never matches a Linear -- SURVEY 0.2-6 -- so: default init)

    def forward(self, Key: Tensor, Scale: Tensor, Shift: Tensor):
        raise NotImplementedError("call through StyleTransformer.forward: the B200 path fuses whole layers (DESIGN.md)")


class StyleDecoder(nn.Module):
    """Mirror of reference :918-1128."""

    def __init__(self, decoder_dim: int, decoder_num_heads: int, decoder_window_size: List[int], decoder_shift_size: List[int],
                 decoder_mlp_ratio: float = 4.0, decoder_dropout: float = 0.0, decoder_attention_dropout: float = 0.0,
                 decoder_qkv_bias: bool = True, decoder_proj_bias: bool = True, decoder_stochastic_depth_prob: float = 0.1,
                 decoder_norm_layer: Callable[..., nn.Module] = nn.LayerNorm,
                 decoder_MLP_activation_layer: Optional[nn.Module] = nn.GELU, decoder_use_instance_norm_with_affine: bool = False,
                 decoder_use_regular_MHA_instead_of_Swin_at_the_end: bool = False,
                 decoder_use_Key_instance_norm_after_linear_transformation: bool = True,
                 decoder_exclude_MLP_after_Fcs_self_MHA: bool = False):
        super().__init__()
        self.decoder_dim, self.decoder_num_heads, self.decoder_mlp_ratio = decoder_dim, decoder_num_heads, decoder_mlp_ratio
        self.decoder_MLP_activation_layer = decoder_MLP_activation_layer
        self.decoder_use_instance_norm_with_affine = decoder_use_instance_norm_with_affine
        self.decoder_use_regular_MHA_instead_of_Swin_at_the_end = decoder_use_regular_MHA_instead_of_Swin_at_the_end
        self.decoder_use_Key_instance_norm_after_linear_transformation = decoder_use_Key_instance_norm_after_linear_transformation
        self.MHA_self_attn = StyleSwinTransformerBlock(
            dim=decoder_dim, num_heads=decoder_num_heads, window_size=decoder_window_size, shift_size=decoder_shift_size,
            dropout=decoder_dropout, attention_dropout=decoder_attention_dropout, qkv_bias=decoder_qkv_bias,
            proj_bias=decoder_proj_bias, mlp_ratio=decoder_mlp_ratio, stochastic_depth_prob=decoder_stochastic_depth_prob,
            norm_layer=decoder_norm_layer, MLP_activation_layer=decoder_MLP_activation_layer,
            exclude_MLP_after=decoder_exclude_MLP_after_Fcs_self_MHA)
        if decoder_use_instance_norm_with_affine:
            self.instance_norm_Query = nn.InstanceNorm2d(decoder_dim, affine=True)
            self.instance_norm_Key = nn.InstanceNorm2d(decoder_dim, affine=True)
            in_q, in_k = self.instance_norm_Query, self.instance_norm_Key
        else:
            self.instance_norm = nn.InstanceNorm2d(decoder_dim, affine=False)
            in_q = in_k = self.instance_norm
        self.stochastic_depth = StochasticDepth(decoder_stochastic_depth_prob, "row")
        self.last_MLP = MLP(decoder_dim, [int(decoder_dim * decoder_mlp_ratio), decoder_dim],
                            activation_layer=decoder_MLP_activation_layer, inplace=None, dropout=decoder_dropout)
        if not decoder_use_regular_MHA_instead_of_Swin_at_the_end:
            self.decoder_MHA_for_sigma_and_mu = ShiftedWindowAttention_for_decoder_last_MHA(
                dim=decoder_dim, num_heads=decoder_num_heads, window_size=decoder_window_size, shift_size=decoder_shift_size,
                instance_norm_q=in_q, instance_norm_k=in_k, dropout=decoder_dropout, attention_dropout=decoder_attention_dropout,
                qkv_bias=decoder_qkv_bias, proj_bias=decoder_proj_bias, use_q_proj=False,
                use_Key_instance_norm_after_linear_transformation=decoder_use_Key_instance_norm_after_linear_transformation)
        else:
            for name in ("linear_transformation_Key", "linear_transformation_Scale", "linear_transformation_Shift", "proj_sigma", "proj_mu"):
                setattr(self, name, nn.Linear(decoder_dim, decoder_dim))
            _xavier_mlp(self.last_MLP)

    def forward(self, Fcs: Tensor, Key: Tensor, Scale: Tensor, Shift: Tensor):
        raise NotImplementedError("call through StyleTransformer.forward: the B200 path fuses whole layers (DESIGN.md)")


class StyleTransformer(nn.Module):
    """Mirror of reference :1133-1245.  forward(Fc, Fs, k) BHWC -> BHWC runs the sm_100a path."""

    def __init__(self, encoder_dim: int, decoder_dim: int, encoder_num_heads: int, decoder_num_heads: int,
                 encoder_window_size: List[int], decoder_window_size: List[int], encoder_shift_size: List[int],
                 decoder_shift_size: List[int], encoder_mlp_ratio: float = 4.0, decoder_mlp_ratio: float = 4.0,
                 encoder_dropout: float = 0.0, decoder_dropout: float = 0.0, encoder_attention_dropout: float = 0.0,
                 decoder_attention_dropout: float = 0.0, encoder_qkv_bias: bool = True, decoder_qkv_bias: bool = True,
                 encoder_proj_bias: bool = True, decoder_proj_bias: bool = True, encoder_stochastic_depth_prob: float = 0.1,
                 decoder_stochastic_depth_prob: float = 0.1, encoder_norm_layer: Callable[..., nn.Module] = None,
                 decoder_norm_layer: Callable[..., nn.Module] = nn.LayerNorm,
                 encoder_MLP_activation_layer: Optional[nn.Module] = nn.GELU,
                 decoder_MLP_activation_layer: Optional[nn.Module] = nn.GELU,
                 encoder_if_use_processed_Key_in_Scale_and_Shift_calculation: bool = True,
                 decoder_use_instance_norm_with_affine: bool = False,
                 decoder_use_regular_MHA_instead_of_Swin_at_the_end: bool = False,
                 decoder_use_Key_instance_norm_after_linear_transformation: bool = True,
                 decoder_exclude_MLP_after_Fcs_self_MHA: bool = False):
        super().__init__()
        self.encoder = StyleEncoder(
            encoder_dim=encoder_dim, encoder_num_heads=encoder_num_heads, encoder_window_size=encoder_window_size,
            encoder_shift_size=encoder_shift_size, encoder_mlp_ratio=encoder_mlp_ratio, encoder_dropout=encoder_dropout,
            encoder_attention_dropout=encoder_attention_dropout, encoder_qkv_bias=encoder_qkv_bias,
            encoder_proj_bias=encoder_proj_bias, encoder_stochastic_depth_prob=encoder_stochastic_depth_prob,
            encoder_norm_layer=encoder_norm_layer, encoder_MLP_activation_layer=encoder_MLP_activation_layer,
            encoder_if_use_processed_Key_in_Scale_and_Shift_calculation=encoder_if_use_processed_Key_in_Scale_and_Shift_calculation)
        self.decoder = StyleDecoder(
            decoder_dim=decoder_dim, decoder_num_heads=decoder_num_heads, decoder_window_size=decoder_window_size,
            decoder_shift_size=decoder_shift_size, decoder_mlp_ratio=decoder_mlp_ratio, decoder_dropout=decoder_dropout,
            decoder_attention_dropout=decoder_attention_dropout, decoder_qkv_bias=decoder_qkv_bias,
            decoder_proj_bias=decoder_proj_bias, decoder_stochastic_depth_prob=decoder_stochastic_depth_prob,
            decoder_norm_layer=decoder_norm_layer, decoder_MLP_activation_layer=decoder_MLP_activation_layer,
            decoder_use_instance_norm_with_affine=decoder_use_instance_norm_with_affine,
            decoder_use_regular_MHA_instead_of_Swin_at_the_end=decoder_use_regular_MHA_instead_of_Swin_at_the_end,
            decoder_use_Key_instance_norm_after_linear_transformation=decoder_use_Key_instance_norm_after_linear_transformation,
            decoder_exclude_MLP_after_Fcs_self_MHA=decoder_exclude_MLP_after_Fcs_self_MHA)
        self._cfg = dict(
            dim=encoder_dim, heads=encoder_num_heads, window=list(encoder_window_size), shift=list(encoder_shift_size),
            default=(encoder_dim == decoder_dim and encoder_num_heads == decoder_num_heads
                     and list(encoder_window_size) == list(decoder_window_size)
                     and list(encoder_shift_size) == list(decoder_shift_size)
                     and encoder_mlp_ratio == decoder_mlp_ratio == 4.0
                     and encoder_norm_layer is None and decoder_norm_layer is nn.LayerNorm
                     and encoder_MLP_activation_layer is nn.GELU and decoder_MLP_activation_layer is nn.GELU
                     and encoder_qkv_bias and decoder_qkv_bias and encoder_proj_bias and decoder_proj_bias
                     and encoder_dropout == decoder_dropout == encoder_attention_dropout == decoder_attention_dropout == 0.0),
            # alternates the inference engine sequences from the same kernels (reference :883-909, :470-472, :389-392)
            flags=dict(processed_key=bool(encoder_if_use_processed_Key_in_Scale_and_Shift_calculation),
                       key_in_after_linear=bool(decoder_use_Key_instance_norm_after_linear_transformation),
                       exclude_mlp=bool(decoder_exclude_MLP_after_Fcs_self_MHA)),
            # variants with kernels of their own in the inference engine (packed from the state_dict: engine.StyleTransformerWeights)
            affine_in=bool(decoder_use_instance_norm_with_affine), regular_mha=bool(decoder_use_regular_MHA_instead_of_Swin_at_the_end))

    def engine_flags(self) -> dict:
        """Keyword arguments of engine.style_transformer_forward that select the reference's alternate orderings."""
        return dict(self._cfg["flags"])

    def _check_config(self, training: bool = False):
        c = self._cfg
        if not c["default"]:
            raise NotImplementedError("this StyleTransformer configuration has no sm_100a kernels (dropout, non-GELU / non-LayerNorm "
                                      "variants, different encoder / decoder geometry: SURVEY.md 8f-4)")
        if training and (c["affine_in"] or c["regular_mha"]):
            raise NotImplementedError("the training step (taped forward + backward kernels) covers the default StyleTransformer configuration and "
                                      "its three re-orderings; affine InstanceNorm and the regular-MHA tail run in inference only (SURVEY.md 8f-4)")
        if c["window"][0] != c["window"][1] or c["shift"][0] != c["shift"][1] or c["dim"] // c["heads"] != 32:
            raise NotImplementedError("square windows and head_dim 32 only")

    def forward(self, Fc: Tensor, Fs: Tensor, k: int = 1) -> Tensor:
        self._check_config()
        require_cuda(Fc, Fs)
        if Fc.shape != Fs.shape or Fc.dim() != 4 or Fc.shape[-1] != self._cfg["dim"]:
            raise ValueError("Fc and Fs must both be [B,H,W,C] with identical shapes")
        B, H, W, C = Fc.shape
        sd_active = self.training and (self.encoder.encoder_stochastic_depth_prob > 0 or self.decoder.stochastic_depth.p > 0)
        if wants_grad(self, Fc, Fs) or sd_active:
            self._check_config(training=True)
            from .autograd_fns import style_transformer_apply
            return style_transformer_apply(self, Fc, Fs, int(k))
        with torch.no_grad():
            w = packed_weights(self, engine.StyleTransformerWeights)
            ws = workspace_of(self, Fc.device)
            out = torch.empty(B, H, W, C, dtype=torch.float32, device=Fc.device)
            engine.style_transformer_forward(w, Fc.float().contiguous(), Fs.float().contiguous(), int(k), ws, B, H, W,
                                             self._cfg["window"][0], self._cfg["shift"][0], self._cfg["heads"], out,
                                             **self.engine_flags())
        return out
