"""Stand-in for the parts of `diffusers==0.27` the reference hooks touch (diffusers is not installable offline).

NOT product code: test / bench infrastructure. Semantics restated from memory of diffusers 0.27
(evidence class [MEMORY]); the reference's closures only rely on the attribute
and child-module NAMES used here (`Attention.to_q/to_k/to_v/to_out/heads/scale/head_to_batch_dim/
get_attention_scores/...`, `unet.down_blocks[i].attentions[j].transformer_blocks[k].attn1`, `ResnetBlock2D.norm1/conv1/...`),
which is what is reproduced. The class literally named `Attention` matters: the reference discovers modules by
`net.__class__.__name__ == 'Attention'` (p2p/model/register.py:78, masactrl/model/register.py:56,
pix2pix-zero/model/attention_control.py:85-89).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F


class _Output(dict):
    """Supports both `.sample` and `["sample"]` like diffusers' BaseOutput."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


class AttnProcessor:
    """Plain attention (what an un-hooked diffusers Attention computes). `use_sdpa = True` switches GPU calls to
    F.scaled_dot_product_attention, diffusers 0.27's default AttnProcessor2_0 (full-size benchmarks); the default is the
    spelled-out arithmetic the golden fixtures were generated with."""
    use_sdpa = False

    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None):
        residual = hidden_states
        input_ndim = hidden_states.ndim
        if input_ndim == 4:
            b, c, hh, ww = hidden_states.shape
            hidden_states = hidden_states.view(b, c, hh * ww).transpose(1, 2)
        q = attn.to_q(hidden_states)
        ctx = hidden_states if encoder_hidden_states is None else encoder_hidden_states
        k, v = attn.to_k(ctx), attn.to_v(ctx)
        if self.use_sdpa and q.is_cuda and attention_mask is None and not (attn.upcast_attention or attn.upcast_softmax):
            h = attn.heads
            q4, k4, v4 = (t.view(t.shape[0], t.shape[1], h, t.shape[2] // h).transpose(1, 2) for t in (q, k, v))
            out = torch.nn.functional.scaled_dot_product_attention(q4, k4, v4, scale=attn.scale)
            out = out.transpose(1, 2).reshape(q.shape[0], q.shape[1], -1)
        else:
            q, k, v = attn.head_to_batch_dim(q), attn.head_to_batch_dim(k), attn.head_to_batch_dim(v)
            probs = attn.get_attention_scores(q, k, attention_mask)
            out = attn.batch_to_head_dim(torch.bmm(probs, v))
        out = attn.to_out[1](attn.to_out[0](out))
        if input_ndim == 4:
            out = out.transpose(-1, -2).reshape(b, c, hh, ww)
        if attn.residual_connection:
            out = out + residual
        return out / attn.rescale_output_factor


class Attention(nn.Module):
    def __init__(self, query_dim: int, cross_attention_dim: Optional[int] = None, heads: int = 8, dim_head: int = 64,
                 dropout: float = 0.0, bias: bool = False, upcast_attention: bool = False, upcast_softmax: bool = False):
        super().__init__()
        self.inner_dim = dim_head * heads
        self.cross_attention_dim = cross_attention_dim if cross_attention_dim is not None else query_dim
        self.upcast_attention, self.upcast_softmax = upcast_attention, upcast_softmax
        self.rescale_output_factor = 1.0
        self.residual_connection = False
        self.heads = heads
        self.scale = dim_head ** -0.5
        self.spatial_norm = None
        self.group_norm = None
        self.norm_cross = None
        self.to_q = nn.Linear(query_dim, self.inner_dim, bias=bias)
        self.to_k = nn.Linear(self.cross_attention_dim, self.inner_dim, bias=bias)
        self.to_v = nn.Linear(self.cross_attention_dim, self.inner_dim, bias=bias)
        self.to_out = nn.ModuleList([nn.Linear(self.inner_dim, query_dim), nn.Dropout(dropout)])
        self.processor = AttnProcessor()

    def set_processor(self, processor):
        self.processor = processor

    def get_processor(self):
        return self.processor

    def head_to_batch_dim(self, t: torch.Tensor) -> torch.Tensor:
        b, n, c = t.shape
        h = self.heads
        return t.reshape(b, n, h, c // h).permute(0, 2, 1, 3).reshape(b * h, n, c // h)

    def batch_to_head_dim(self, t: torch.Tensor) -> torch.Tensor:
        bh, n, d = t.shape
        h = self.heads
        return t.reshape(bh // h, h, n, d).permute(0, 2, 1, 3).reshape(bh // h, n, d * h)

    def prepare_attention_mask(self, attention_mask, target_length, batch_size, out_dim=3):
        if attention_mask is None:
            return None
        raise NotImplementedError("stand-in Attention: attention masks are not used by any reference script")

    def get_attention_scores(self, query, key, attention_mask=None):
        dtype = query.dtype
        if self.upcast_attention:
            query, key = query.float(), key.float()
        if attention_mask is None:
            base = torch.empty(query.shape[0], query.shape[1], key.shape[1], dtype=query.dtype, device=query.device)
            beta = 0
        else:
            base, beta = attention_mask, 1
        scores = torch.baddbmm(base, query, key.transpose(-1, -2), beta=beta, alpha=self.scale)
        if self.upcast_softmax:
            scores = scores.float()
        return scores.softmax(dim=-1).to(dtype)

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None, **cross_attention_kwargs):
        return self.processor(self, hidden_states, encoder_hidden_states=encoder_hidden_states, attention_mask=attention_mask,
                              **cross_attention_kwargs)


class GEGLU(nn.Module):
    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)

    def forward(self, x):
        x, gate = self.proj(x).chunk(2, dim=-1)
        return x * F.gelu(gate)


class FeedForward(nn.Module):
    def __init__(self, dim, mult=4):
        super().__init__()
        self.net = nn.ModuleList([GEGLU(dim, dim * mult), nn.Dropout(0.0), nn.Linear(dim * mult, dim)])

    def forward(self, x):
        for m in self.net:
            x = m(x)
        return x


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim, heads, dim_head, cross_attention_dim):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.attn1 = Attention(dim, None, heads, dim_head)
        self.norm2 = nn.LayerNorm(dim)
        self.attn2 = Attention(dim, cross_attention_dim, heads, dim_head)
        self.norm3 = nn.LayerNorm(dim)
        self.ff = FeedForward(dim)

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None, encoder_attention_mask=None):
        hidden_states = self.attn1(self.norm1(hidden_states), encoder_hidden_states=None, attention_mask=attention_mask) + hidden_states
        hidden_states = self.attn2(self.norm2(hidden_states), encoder_hidden_states=encoder_hidden_states,
                                   attention_mask=encoder_attention_mask) + hidden_states
        return self.ff(self.norm3(hidden_states)) + hidden_states


class Transformer2DModel(nn.Module):
    def __init__(self, heads, dim_head, in_channels, num_layers, cross_attention_dim, groups, use_linear_projection):
        super().__init__()
        inner = heads * dim_head
        self.use_linear_projection = use_linear_projection
        self.norm = nn.GroupNorm(groups, in_channels, eps=1e-6)
        self.proj_in = nn.Linear(in_channels, inner) if use_linear_projection else nn.Conv2d(in_channels, inner, 1)
        self.transformer_blocks = nn.ModuleList(
            [BasicTransformerBlock(inner, heads, dim_head, cross_attention_dim) for _ in range(num_layers)])
        self.proj_out = nn.Linear(inner, in_channels) if use_linear_projection else nn.Conv2d(inner, in_channels, 1)

    def forward(self, hidden_states, encoder_hidden_states=None):
        b, c, h, w = hidden_states.shape
        residual = hidden_states
        x = self.norm(hidden_states)
        if self.use_linear_projection:
            x = self.proj_in(x.permute(0, 2, 3, 1).reshape(b, h * w, c))
        else:
            x = self.proj_in(x).permute(0, 2, 3, 1).reshape(b, h * w, -1)
        for blk in self.transformer_blocks:
            x = blk(x, encoder_hidden_states=encoder_hidden_states)
        if self.use_linear_projection:
            x = self.proj_out(x).reshape(b, h, w, c).permute(0, 3, 1, 2).contiguous()
        else:
            x = self.proj_out(x.reshape(b, h, w, -1).permute(0, 3, 1, 2).contiguous())
        return x + residual


class Upsample2D(nn.Module):
    def __init__(self, channels):
        super().__init__()
        self.conv = nn.Conv2d(channels, channels, 3, padding=1)

    def forward(self, x, output_size=None):
        if output_size is None:
            x = F.interpolate(x, scale_factor=2.0, mode="nearest")
        else:
            x = F.interpolate(x, size=output_size, mode="nearest")
        return self.conv(x)


class Downsample2D(nn.Module):
    def __init__(self, channels):
        super().__init__()
        self.conv = nn.Conv2d(channels, channels, 3, stride=2, padding=1)

    def forward(self, x):
        return self.conv(x)


class ResnetBlock2D(nn.Module):
    def __init__(self, in_channels, out_channels, temb_channels, groups):
        super().__init__()
        self.time_embedding_norm = "default"
        self.skip_time_act = False
        self.output_scale_factor = 1.0
        self.upsample = self.downsample = None
        self.norm1 = nn.GroupNorm(groups, in_channels, eps=1e-5)
        self.conv1 = nn.Conv2d(in_channels, out_channels, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb_channels, out_channels)
        self.norm2 = nn.GroupNorm(groups, out_channels, eps=1e-5)
        self.dropout = nn.Dropout(0.0)
        self.conv2 = nn.Conv2d(out_channels, out_channels, 3, padding=1)
        self.nonlinearity = nn.SiLU()
        self.conv_shortcut = nn.Conv2d(in_channels, out_channels, 1) if in_channels != out_channels else None

    def forward(self, input_tensor, temb, scale=1.0):
        h = self.conv1(self.nonlinearity(self.norm1(input_tensor)))
        if self.time_emb_proj is not None:
            h = h + self.time_emb_proj(self.nonlinearity(temb))[:, :, None, None]
        h = self.conv2(self.dropout(self.nonlinearity(self.norm2(h))))
        if self.conv_shortcut is not None:
            input_tensor = self.conv_shortcut(input_tensor)
        return (input_tensor + h) / self.output_scale_factor


class _DownBlock(nn.Module):
    def __init__(self, cin, cout, temb, groups, layers, tf_layers, heads, dim_head, ctx_dim, add_down, linear_proj):
        super().__init__()
        self.has_cross_attention = tf_layers > 0
        self.resnets = nn.ModuleList([ResnetBlock2D(cin if i == 0 else cout, cout, temb, groups) for i in range(layers)])
        if self.has_cross_attention:
            self.attentions = nn.ModuleList(
                [Transformer2DModel(heads, dim_head, cout, tf_layers, ctx_dim, groups, linear_proj) for _ in range(layers)])
        self.downsamplers = nn.ModuleList([Downsample2D(cout)]) if add_down else None

    def forward(self, h, temb, ctx):
        outs = []
        for i, res in enumerate(self.resnets):
            h = res(h, temb)
            if self.has_cross_attention:
                h = self.attentions[i](h, encoder_hidden_states=ctx)
            outs.append(h)
        if self.downsamplers is not None:
            h = self.downsamplers[0](h)
            outs.append(h)
        return h, outs


class _MidBlock(nn.Module):
    def __init__(self, c, temb, groups, tf_layers, heads, dim_head, ctx_dim, linear_proj):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(c, c, temb, groups), ResnetBlock2D(c, c, temb, groups)])
        self.attentions = nn.ModuleList([Transformer2DModel(heads, dim_head, c, tf_layers, ctx_dim, groups, linear_proj)])

    def forward(self, h, temb, ctx):
        h = self.resnets[0](h, temb)
        h = self.attentions[0](h, encoder_hidden_states=ctx)
        return self.resnets[1](h, temb)


class _UpBlock(nn.Module):
    def __init__(self, cin, cout, cprev, temb, groups, layers, tf_layers, heads, dim_head, ctx_dim, add_up, linear_proj):
        super().__init__()
        self.has_cross_attention = tf_layers > 0
        res = []
        for i in range(layers):
            skip = cin if i == layers - 1 else cout
            res.append(ResnetBlock2D((cprev if i == 0 else cout) + skip, cout, temb, groups))
        self.resnets = nn.ModuleList(res)
        if self.has_cross_attention:
            self.attentions = nn.ModuleList(
                [Transformer2DModel(heads, dim_head, cout, tf_layers, ctx_dim, groups, linear_proj) for _ in range(layers)])
        self.upsamplers = nn.ModuleList([Upsample2D(cout)]) if add_up else None

    def forward(self, h, skips, temb, ctx):
        for i, res in enumerate(self.resnets):
            h = res(torch.cat([h, skips.pop()], dim=1), temb)
            if self.has_cross_attention:
                h = self.attentions[i](h, encoder_hidden_states=ctx)
        if self.upsamplers is not None:
            h = self.upsamplers[0](h)
        return h


class Timesteps(nn.Module):
    def __init__(self, num_channels):
        super().__init__()
        self.num_channels = num_channels

    def forward(self, timesteps):
        half = self.num_channels // 2
        exponent = -math.log(10000) * torch.arange(half, dtype=torch.float32, device=timesteps.device) / half
        emb = timesteps[:, None].float() * torch.exp(exponent)[None, :]
        return torch.cat([torch.cos(emb), torch.sin(emb)], dim=-1)  # flip_sin_to_cos=True


class TimestepEmbedding(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.linear_1, self.act, self.linear_2 = nn.Linear(cin, cout), nn.SiLU(), nn.Linear(cout, cout)

    def forward(self, x):
        return self.linear_2(self.act(self.linear_1(x)))


@dataclass
class UNetConfig:
    in_channels: int = 4
    out_channels: int = 4
    sample_size: int = 64
    block_out_channels: Tuple[int, ...] = (320, 640, 1280, 1280)
    layers_per_block: int = 2
    # transformer layers per down block (0 = block without attention); SD: (1,1,1,0), SDXL: (0,2,10)
    transformer_layers: Tuple[int, ...] = (1, 1, 1, 0)
    # number of heads per block (SD-1.5: 8 everywhere; SD-2.1: 5,10,20,20; SDXL: 5,10,20)
    num_heads: Tuple[int, ...] = (8, 8, 8, 8)
    cross_attention_dim: int = 768
    norm_num_groups: int = 32
    use_linear_projection: bool = False
    name: str = "sd15"

    def __getitem__(self, k):
        return getattr(self, k)


def sd15_config() -> UNetConfig:
    return UNetConfig()


def sd21_config() -> UNetConfig:
    return UNetConfig(sample_size=96, num_heads=(5, 10, 20, 20), cross_attention_dim=1024, use_linear_projection=True, name="sd21")


def sdxl_config() -> UNetConfig:
    return UNetConfig(sample_size=128, block_out_channels=(320, 640, 1280), transformer_layers=(0, 2, 10), num_heads=(5, 10, 20),
                      cross_attention_dim=2048, use_linear_projection=True, name="sdxl")


def tiny_config(heads: int = 2, ctx: int = 32) -> UNetConfig:
    """A 32-attention-layer miniature with SD-1.5's topology, cheap enough for CPU tests."""
    return UNetConfig(sample_size=16, block_out_channels=(16 * heads, 32 * heads, 32 * heads, 32 * heads), num_heads=(heads,) * 4,
                      cross_attention_dim=ctx, norm_num_groups=8, name="tiny")


class UNet2DConditionModel(nn.Module):
    def __init__(self, config: UNetConfig):
        super().__init__()
        self.config = config
        ch = config.block_out_channels
        temb = ch[0] * 4
        g = config.norm_num_groups
        self.conv_in = nn.Conv2d(config.in_channels, ch[0], 3, padding=1)
        self.time_proj = Timesteps(ch[0])
        self.time_embedding = TimestepEmbedding(ch[0], temb)
        self.down_blocks = nn.ModuleList()
        cout = ch[0]
        for i, c in enumerate(ch):
            cin, cout = cout, c
            heads = config.num_heads[i]
            self.down_blocks.append(_DownBlock(cin, cout, temb, g, config.layers_per_block, config.transformer_layers[i], heads,
                                               cout // heads, config.cross_attention_dim, i != len(ch) - 1, config.use_linear_projection))
        mid_tf = max(config.transformer_layers[-1], 1) if config.transformer_layers[-1] == 0 and len(ch) == 4 else config.transformer_layers[-1]
        self.mid_block = _MidBlock(ch[-1], temb, g, mid_tf, config.num_heads[-1], ch[-1] // config.num_heads[-1],
                                   config.cross_attention_dim, config.use_linear_projection)
        self.up_blocks = nn.ModuleList()
        rev = list(reversed(ch))
        rev_tf = list(reversed(config.transformer_layers))
        rev_heads = list(reversed(config.num_heads))
        cout = rev[0]
        for i, c in enumerate(rev):
            cprev, cout = cout, c
            cin = rev[min(i + 1, len(ch) - 1)]
            self.up_blocks.append(_UpBlock(cin, cout, cprev, temb, g, config.layers_per_block + 1, rev_tf[i], rev_heads[i],
                                           cout // rev_heads[i], config.cross_attention_dim, i != len(ch) - 1, config.use_linear_projection))
        self.conv_norm_out = nn.GroupNorm(g, ch[0], eps=1e-5)
        self.conv_act = nn.SiLU()
        self.conv_out = nn.Conv2d(ch[0], config.out_channels, 3, padding=1)

    @property
    def device(self):
        return next(self.parameters()).device

    @property
    def dtype(self):
        return next(self.parameters()).dtype

    def forward(self, sample, timestep, encoder_hidden_states, cross_attention_kwargs=None, added_cond_kwargs=None, **kw):
        if cross_attention_kwargs:
            raise NotImplementedError("stand-in UNet: cross_attention_kwargs are not supported (the reference passes None)")
        t = timestep
        if not torch.is_tensor(t):
            t = torch.tensor([t], dtype=torch.int64, device=sample.device)
        elif t.dim() == 0:
            t = t[None].to(sample.device)
        t = t.expand(sample.shape[0])
        emb = self.time_embedding(self.time_proj(t).to(dtype=sample.dtype))
        h = self.conv_in(sample)
        skips = [h]
        for blk in self.down_blocks:
            h, outs = blk(h, emb, encoder_hidden_states)
            skips.extend(outs)
        h = self.mid_block(h, emb, encoder_hidden_states)
        for blk in self.up_blocks:
            h = blk(h, skips, emb, encoder_hidden_states)
        h = self.conv_out(self.conv_act(self.conv_norm_out(h)))
        return _Output(sample=h)


def attention_geometry(config: UNetConfig, latent_hw: Optional[int] = None) -> List[Tuple[str, int, int, int]]:
    """(place, tokens, heads, head_dim) of every SELF-attention layer in forward order."""
    hw = latent_hw or config.sample_size
    out = []
    ch = config.block_out_channels
    for i, c in enumerate(ch):
        res = hw >> i
        for _ in range(config.layers_per_block):
            for _ in range(config.transformer_layers[i]):
                out.append(("down", res * res, config.num_heads[i], c // config.num_heads[i]))
    res = hw >> (len(ch) - 1)
    mid_tf = max(config.transformer_layers[-1], 1) if config.transformer_layers[-1] == 0 and len(ch) == 4 else config.transformer_layers[-1]
    for _ in range(mid_tf):
        out.append(("mid", res * res, config.num_heads[-1], ch[-1] // config.num_heads[-1]))
    for i, c in enumerate(reversed(ch)):
        j = len(ch) - 1 - i
        res = hw >> j
        for _ in range(config.layers_per_block + 1):
            for _ in range(config.transformer_layers[j]):
                out.append(("up", res * res, config.num_heads[j], c // config.num_heads[j]))
    return out
