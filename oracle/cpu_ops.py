"""Oracle-backed CPU stand-ins for image_editing_framework_b200.ops — TEST / BASELINE INFRASTRUCTURE ONLY.

They let `-m "not gpu"` tests drive the real host logic (register closures, controller state machines, row tables, store
bookkeeping, drivers) on a CPU against the golden vectors, and they are what bench.py's `cpu_baseline` / `--impl reference`
legs time: the reference's materialise-edit-multiply arithmetic (oracle/controlled_attention.py) in fp32 on the host cores.
The product never imports this module; it is installed by explicit patching only (install / patched).
"""
import contextlib

import torch

from oracle import controlled_attention as orc
from image_editing_framework_b200 import ops as real_ops


def _as3(t, heads):
    if t.dim() == 4:
        b, n, h, d = t.shape
        return t.reshape(b, n, h * d)
    return t


# bench.py's reference-formulation arms set this (patched(reference_work=True)): the reference closure materialises `sim` and `attn`
# for EVERY layer before it calls the editor (masactrl/model/register.py:35-44), and a controlled layer then recomputes its own
# probabilities (attention_control.py:37-50) — so a controlled layer costs the plain probabilities (discarded) on top of the indexed ones.
REFERENCE_WORK = False


def attention(q, k, v, heads, scale, *, q_src=None, k_src=None, v_src=None, k_src2=None, v_src2=None, impl=0, probs_out=None,
              probs_accum=False, probs_slot=None, rows=None, key_bias=None, bias_sel=None, out=None):
    q, k, v = _as3(q, heads), _as3(k, heads), _as3(v, heads)
    B = q.shape[0]
    ident = list(range(B))
    if REFERENCE_WORK and (q_src is not None or k_src is not None or v_src is not None):
        orc.attention_probs(q, k, heads, scale)      # the closure's own sim + attn, thrown away by the editor
    qq, kk, vv = q[list(q_src or ident)], k[list(k_src or ident)], v[list(v_src or ident)]
    if k_src2 is not None:
        kk, vv = torch.cat([kk, k[list(k_src2)]], 1), torch.cat([vv, v[list(v_src2)]], 1)
    bias = None
    if key_bias is not None:
        bias = torch.stack([key_bias[s] if s >= 0 else torch.zeros_like(key_bias[0]) for s in bias_sel])
    p = orc.attention_probs(qq, kk, heads, scale, key_bias=bias)
    o = orc.apply_probs(p, vv, heads).to(q.dtype)
    if probs_out is not None:
        p4 = p.reshape(B, heads, *p.shape[1:])
        po = probs_out.view(-1, heads, *p.shape[1:])
        for b in range(B):
            s = b if probs_slot is None else probs_slot[b]
            if s >= 0:
                po[s] = po[s] + p4[b] if probs_accum else p4[b]
    if out is None:
        return o
    sel = list(range(B)) if rows is None else list(rows)
    out[sel] = o[sel]
    return out


def cross_attention_edit(q, k, v, heads, scale, *, edit=None, step_alpha=None, base_row=None, edit_slot=None, probs_out=None,
                         probs_accum=False, store_slot=None, out=None):
    q, k, v = _as3(q, heads), _as3(k, heads), _as3(v, heads)
    B, M = q.shape[0], k.shape[1]
    p = orc.attention_probs(q, k, heads, scale)
    p4 = p.reshape(B, heads, *p.shape[1:]).clone()
    orig = p4.clone()
    if base_row is not None and edit is not None:
        for b in range(B):
            if base_row[b] < 0:
                continue
            e = edit_slot[b] if edit_slot is not None else 0
            base, repl = orig[base_row[b]], orig[b]
            if edit.mode == real_ops.IEF_EDIT_REPLACE:
                new = torch.einsum('hpw,wn->hpn', base, edit.mapper[e].float())
            elif edit.mode == real_ops.IEF_EDIT_REFINE:
                new = base[:, :, edit.mapper_idx[e].long()] * edit.refine_alpha[e] + repl * (1 - edit.refine_alpha[e])
            else:
                new = base
            if edit.equalizer is not None:
                new = new * edit.equalizer[e]
            a = step_alpha.reshape(-1, M)[e]
            p4[b] = new * a + (1 - a) * repl
    o = orc.apply_probs(p4.reshape(B * heads, *p.shape[1:]), v, heads).to(q.dtype)
    if probs_out is not None:
        po = probs_out.view(-1, heads, *p.shape[1:])
        for b in range(B):
            s = b if store_slot is None else store_slot[b]
            if s >= 0:
                po[s] = po[s] + p4[b] if probs_accum else p4[b]
    if out is not None:
        out.copy_(o)
        return out
    return o


def mask_blend(fg, bg, w, rows=None):
    sel = list(range(fg.shape[0])) if rows is None else list(rows)
    wf = w.float().reshape(1, -1, 1)
    fg[sel] = (fg[sel].float() * wf + bg[sel].float() * (1 - wf)).to(fg.dtype)
    return fg


def store_accumulate(dst, src):
    for d, s in zip(dst, src):
        d += s


def local_blend(x_t, maps, n_prompts, word_alpha, threshold, res=16, return_mask=False):
    r = orc.local_blend(x_t, maps, word_alpha.reshape(n_prompts, -1), threshold, res=res, return_mask=return_mask)
    if return_mask:
        x_t.copy_(r[0])
        return x_t, r[1]
    x_t.copy_(r)
    return x_t


def cfg_ddim_step(eps_uncond, eps_cond, x, guidance, alpha_t, alpha_prev, out=None):
    eps = eps_uncond if eps_cond is None else orc.cfg_combine(eps_uncond, eps_cond, guidance)
    return orc.ddim_step(eps, x, torch.tensor(alpha_t, dtype=torch.float32), torch.tensor(alpha_prev, dtype=torch.float32)).to(x.dtype)


_NAMES = ("attention", "cross_attention_edit", "mask_blend", "store_accumulate", "local_blend", "cfg_ddim_step")


def install(monkeypatch):
    """pytest flavour: undone automatically at test teardown."""
    from image_editing_framework_b200 import hooks
    monkeypatch.setattr(hooks, "compute_dtype", lambda t: t.dtype)  # keep fp32: isolates host logic from bf16 rounding
    for name in _NAMES:
        monkeypatch.setattr(real_ops, name, globals()[name])


@contextlib.contextmanager
def patched(reference_work: bool = False):
    """bench.py flavour: route the host logic to the oracle (torch fp32 on whatever device the tensors live on) inside the `with`
    block only. reference_work: also spend the work the reference spends and discards (see REFERENCE_WORK)."""
    global REFERENCE_WORK
    from image_editing_framework_b200 import hooks
    saved = {n: getattr(real_ops, n) for n in _NAMES}
    saved_dt, saved_rw = hooks.compute_dtype, REFERENCE_WORK
    try:
        REFERENCE_WORK = reference_work
        hooks.compute_dtype = lambda t: t.dtype
        for n in _NAMES:
            setattr(real_ops, n, globals()[n])
        yield
    finally:
        hooks.compute_dtype = saved_dt
        REFERENCE_WORK = saved_rw
        for n, f in saved.items():
            setattr(real_ops, n, f)
