"""fp32 restatement of every function on the hot path, in the reference's own "materialise the probabilities, edit
them, multiply by V" form — deliberately NOT the fused formulation the kernels use, so agreement is evidence.

TEST INFRASTRUCTURE (see oracle/__init__.py). Citations are relative to the reference root.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F


# ---------------------------------------------------------------------------------------------- layout helpers
def head_to_batch(t: torch.Tensor, heads: int) -> torch.Tensor:
    """diffusers Attention.head_to_batch_dim: [B, N, H*d] -> [B*H, N, d], batch-major / head-minor."""
    b, n, c = t.shape
    return t.reshape(b, n, heads, c // heads).permute(0, 2, 1, 3).reshape(b * heads, n, c // heads)


def batch_to_head(t: torch.Tensor, heads: int) -> torch.Tensor:
    bh, n, d = t.shape
    return t.reshape(bh // heads, heads, n, d).permute(0, 2, 1, 3).reshape(bh // heads, n, heads * d)


def attention_probs(q: torch.Tensor, k: torch.Tensor, heads: int, scale: float, key_bias=None) -> torch.Tensor:
    """softmax(scale * Q K^T) as [B*H, N, M] fp32 (diffusers get_attention_scores, called at p2p/model/register.py:47;
    the same einsum->softmax is spelled out at masactrl/model/register.py:35-44 and pnp/model/register.py:65-75).
    key_bias [B, M]: added to the scaled scores before the softmax, the way the masked MasaCtrl variants add their
    fore-/background key masks (masactrl/model/attention_control.py:139-147, 238-246)."""
    qh, kh = head_to_batch(q.float(), heads), head_to_batch(k.float(), heads)
    sim = torch.bmm(qh, kh.transpose(1, 2)) * scale
    if key_bias is not None:
        sim = sim + key_bias.float().repeat_interleave(heads, 0)[:, None, :]
    return sim.softmax(dim=-1)


def apply_probs(p: torch.Tensor, v: torch.Tensor, heads: int) -> torch.Tensor:
    """bmm(probs, V) + batch_to_head_dim (p2p/model/register.py:50-51)."""
    return batch_to_head(torch.bmm(p, head_to_batch(v.float(), heads)), heads)


def plain_attention(q, k, v, heads, scale):
    return apply_probs(attention_probs(q, k, heads, scale), v, heads)


def indexed_attention(q, k, v, heads, scale, q_src=None, k_src=None, v_src=None, k_src2=None, v_src2=None):
    """Definition of ief_attn_fwd: O[b] = softmax(scale Q[q_src[b]] [K[k_src[b]]; K[k_src2[b]]]^T) [V[v_src[b]]; V[v_src2[b]]]."""
    B = q.shape[0]
    ident = list(range(B))
    qq = q[list(q_src or ident)]
    kk, vv = k[list(k_src or ident)], v[list(v_src or ident)]
    if k_src2 is not None:
        kk = torch.cat([kk, k[list(k_src2)]], dim=1)
        vv = torch.cat([vv, v[list(v_src2)]], dim=1)
    return plain_attention(qq, kk, vv, heads, scale)


# ---------------------------------------------------------------------------------------------- Prompt-to-Prompt
def p2p_edit_probs(probs: torch.Tensor, heads: int, n_prompts: int, is_cross: bool, cur_step: int, *, mode: str,
                   alpha_table: torch.Tensor, num_self_replace=(0, 0), mapper: Optional[torch.Tensor] = None,
                   refine_alphas: Optional[torch.Tensor] = None, equalizer: Optional[torch.Tensor] = None,
                   low_resource: bool = False) -> torch.Tensor:
    """AttentionControl.__call__ (p2p/model/attention_base.py:16-28, non-LOW_RESOURCE edits only the cond half
    attn[h//2:]) followed by AttentionControlEdit.forward (:113-125). probs: [B*H, N, M]; returns the edited copy.

    mode: 'replace' (attention_control.py:15-16), 'refine' (:28-31), 'none' (base copied), optionally followed by the
    equaliser of AttentionReweight (:42-46). alpha_table: cross_replace_alpha [steps+1, n_prompts-1, 1, 1, 77].
    """
    out = probs.clone()
    lo = 0 if low_resource else probs.shape[0] // 2
    part = out[lo:]
    if is_cross or (num_self_replace[0] <= cur_step < num_self_replace[1]):
        a = part.reshape(n_prompts, part.shape[0] // n_prompts, *part.shape[1:])  # [prompts, heads, N, M] view
        base, repl = a[0], a[1:]
        if is_cross:
            if mode == "replace":
                new = torch.einsum('hpw,bwn->bhpn', base, mapper.float())
            elif mode == "refine":
                gathered = base[:, :, mapper].permute(2, 0, 1, 3)           # mapper may hold -1 -> last column
                new = gathered * refine_alphas + repl * (1 - refine_alphas)
            else:
                new = base.unsqueeze(0).expand_as(repl)
            if equalizer is not None:
                new = new * equalizer[:, None, None, :]
            alpha_words = alpha_table[cur_step]
            a[1:] = new * alpha_words + (1 - alpha_words) * repl
        elif repl.shape[2] <= 16 ** 2:                                       # replace_self_attention :132-136
            a[1:] = base.unsqueeze(0).expand(repl.shape[0], *base.shape)
    return out


# ---------------------------------------------------------------------------------------------- MasaCtrl
def masactrl_mutual(q, k, v, heads, scale):
    """MutualSelfAttentionControl.forward on a controlled layer (masactrl/model/attention_control.py:52-68): per CFG half,
    the queries of all rows are stacked along the sequence and attend to the keys/values of the half's first row."""
    def half(qh, kh, vh):
        qs = head_to_batch(qh.float(), heads)          # '(b h) n d'
        ks = head_to_batch(kh.float(), heads)[:heads]  # ku[:num_heads]
        vs = head_to_batch(vh.float(), heads)[:heads]
        b = qs.shape[0] // heads
        n, d = qs.shape[1], qs.shape[2]
        qq = qs.reshape(b, heads, n, d).permute(1, 0, 2, 3).reshape(heads, b * n, d)   # 'h (b n) d'
        attn = (torch.einsum("hid,hjd->hij", qq, ks) * scale).softmax(-1)
        o = torch.einsum("hij,hjd->hid", attn, vs)
        return o.reshape(heads, b, n, d).permute(1, 2, 0, 3).reshape(b, n, heads * d)  # 'b n (h d)'
    qu, qc = q.chunk(2)
    ku, kc = k.chunk(2)
    vu, vc = v.chunk(2)
    return torch.cat([half(qu, ku, vu), half(qc, kc, vc)], dim=0)


# ---------------------------------------------------------------------------------------------- Plug-and-Play
def pnp_inject_qk(q: torch.Tensor, k: torch.Tensor):
    """pnp/model/register.py:46-52: the source rows' q and k overwrite the unconditional and conditional target rows."""
    q, k = q.clone(), k.clone()
    s = q.shape[0] // 4
    q[s:2 * s], k[s:2 * s] = q[2 * s:3 * s], k[2 * s:3 * s]
    q[3 * s:4 * s], k[3 * s:4 * s] = q[2 * s:3 * s], k[2 * s:3 * s]
    return q, k


# ---------------------------------------------------------------------------------------------- DDIM
def cfg_combine(eps_uncond, eps_cond, guidance: float):
    return eps_uncond + guidance * (eps_cond - eps_uncond)       # p2p/model/sd_utils.py:75


def ddim_step(eps, sample, alpha_t: torch.Tensor, alpha_prev: torch.Tensor):
    """diffusers DDIMScheduler.step, eta = 0, epsilon prediction, no clipping (called at p2p/model/sd_utils.py:76);
    the same closed form, upwards in t, is ddim_reverse (inversion/ddim.py:9-18)."""
    x0 = (sample - (1 - alpha_t) ** 0.5 * eps) / alpha_t ** 0.5
    return alpha_prev ** 0.5 * x0 + (1 - alpha_prev) ** 0.5 * eps


# ---------------------------------------------------------------------------------------------- LocalBlend
def local_blend(x_t: torch.Tensor, maps: Sequence[torch.Tensor], alpha_layers: torch.Tensor, threshold: float, res: int = 16,
                return_mask: bool = False):
    """LocalBlend.__call__ (p2p/model/ptp_utils.py:20-32). maps: the five [prompts*heads, res*res, 77] stored tensors."""
    n_prompts, words = alpha_layers.shape[0], alpha_layers.shape[-1]
    k = 1
    m = torch.cat([item.reshape(n_prompts, -1, 1, res, res, words) for item in maps], dim=1)
    m = (m * alpha_layers.reshape(n_prompts, 1, 1, 1, 1, words)).sum(-1).mean(1)
    mask = F.max_pool2d(m, (k * 2 + 1, k * 2 + 1), (1, 1), padding=(k, k))
    mask = F.interpolate(mask, size=(x_t.shape[2:]))
    mask = mask / mask.max(2, keepdims=True)[0].max(3, keepdims=True)[0]
    per_prompt = mask.gt(threshold)
    mask = (per_prompt[:1] + per_prompt[1:]).float()
    out = x_t[:1] + mask * (x_t - x_t[:1])
    return (out, per_prompt[:, 0].float()) if return_mask else out
