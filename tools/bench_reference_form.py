"""The reference's own formulation of the hot ops, run as-is (fp32, torch eager, probabilities materialised) on the SAME B200,
next to the fused kernels — the fair GPU baseline SURVEY.md section 8(d) asks for. Self-contained torch restatement of
masactrl/model/attention_control.py:37-68 (attn_batch / forward) and p2p/model/register.py:43-51 +
attention_base.py:113-125 + attention_control.py:15-16 (cross-attention replace edit); no oracle import.
CUDA events, L2 flushed between repetitions, GPU kept busy while the host enqueues."""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from image_editing_framework_b200 import ops

dev = torch.device("cuda:0")
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def time_call(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        torch.cuda._sleep(400000)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        e.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


def head_to_batch(t, H):  # [B, N, H*d] -> [(B H), N, d]   (diffusers head_to_batch_dim)
    B, N, C = t.shape
    return t.reshape(B, N, H, C // H).permute(0, 2, 1, 3).reshape(B * H, N, C // H)


def masactrl_reference(q, k, v, H, scale):
    """forward :52-68 with attn_batch :37-50 on '(b h) n d' tensors (q, k, v already head-major, fp32)."""
    def attn_batch(qh, kh, vh):
        b = qh.shape[0] // H
        n, d = qh.shape[1], qh.shape[2]
        qs = qh.reshape(b, H, n, d).permute(1, 0, 2, 3).reshape(H, b * n, d)
        sim = torch.einsum("h i d, h j d -> h i j".replace(" ", ""), qs, kh) * scale
        attn = sim.softmax(-1)
        out = torch.einsum("hij,hjd->hid", attn, vh)
        return out.reshape(H, b, n, d).permute(1, 2, 0, 3).reshape(b, n, H * d)
    qu, qc = q.chunk(2)
    ku, kc = k.chunk(2)
    vu, vc = v.chunk(2)
    return torch.cat([attn_batch(qu, ku[:H], vu[:H]), attn_batch(qc, kc[:H], vc[:H])], 0)


def p2p_cross_reference(q, k, v, H, scale, mapper, alpha):
    """register.py:47-50 with AttentionControl.__call__ (:21-22 cond half edited in place) and the replace edit."""
    attn = (torch.bmm(q, k.transpose(1, 2)) * scale).softmax(-1)            # get_attention_scores
    h = attn.shape[0]
    cond = attn[h // 2:]
    a4 = cond.reshape(2, H, *cond.shape[1:])
    base, repl = a4[0], a4[1:]
    new = torch.einsum("hpw,bwn->bhpn", base, mapper) * alpha + (1 - alpha) * repl
    a4[1:] = new
    attn[h // 2:] = a4.reshape(cond.shape)
    return torch.bmm(attn, v)


def main():
    B, H = 4, 8
    for N, d in ((4096, 40), (1024, 80)):
        q, k, v = (torch.randn(B, N, H * d, device=dev) for _ in range(3))
        qh, kh, vh = (head_to_batch(t, H).contiguous() for t in (q, k, v))
        scale = d ** -0.5
        ms_ref = time_call(lambda: masactrl_reference(qh, kh, vh, H, scale))
        qb, kb, vb = (t.to(torch.bfloat16) for t in (q, k, v))
        ms_ours = time_call(lambda: ops.attention(qb, kb, vb, H, scale, k_src=[0, 0, 2, 2], v_src=[0, 0, 2, 2]), reps=20)
        flops = 4 * B * H * N * N * d
        print(json.dumps(dict(op=f"MasaCtrl mutual self-attention B=4 H=8 N={N} d={d}", reference_fp32_eager_ms=round(ms_ref, 3),
                              fused_bf16_ms=round(ms_ours, 4), speedup=round(ms_ref / ms_ours, 1), reference_TFLOPs=round(flops / ms_ref / 1e9, 1),
                              fused_TFLOPs=round(flops / ms_ours / 1e9, 1), reference_probs_GiB=round(2 * B * H * N * N * 4 / 2 ** 30, 2))), flush=True)
    mapper = torch.eye(77, device=dev)[None]
    alpha = torch.ones(1, 1, 1, 77, device=dev)
    edit = ops.CrossEdit(ops.IEF_EDIT_REPLACE, 1, mapper=mapper.contiguous())
    al = torch.ones(1, 77, device=dev)
    for N, d in ((4096, 40), (1024, 80), (256, 160)):
        q = torch.randn(B, N, H * d, device=dev)
        k, v = (torch.randn(B, 77, H * d, device=dev) for _ in range(2))
        qh, kh, vh = (head_to_batch(t, H).contiguous() for t in (q, k, v))
        scale = d ** -0.5
        ms_ref = time_call(lambda: p2p_cross_reference(qh, kh, vh, H, scale, mapper, alpha))
        qb, kb, vb = (t.to(torch.bfloat16) for t in (q, k, v))
        o = torch.empty_like(qb)
        ms_ours = time_call(lambda: ops.cross_attention_edit(qb, kb, vb, H, scale, edit=edit, step_alpha=al, base_row=[-1, -1, -1, 2], edit_slot=[0] * 4, out=o), reps=20)
        print(json.dumps(dict(op=f"P2P cross-attention + replace edit B=4 H=8 N={N} d={d}", reference_fp32_eager_ms=round(ms_ref, 3),
                              fused_bf16_ms=round(ms_ours, 4), speedup=round(ms_ref / ms_ours, 1))), flush=True)


if __name__ == "__main__":
    main()
