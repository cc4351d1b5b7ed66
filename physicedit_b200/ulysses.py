"""Sequence-parallel (Ulysses / head-parallel) DiT forward: ONE image on N GPUs (the 2048^2 latency case).

The reference has no within-image parallelism for this model (its only sequence-parallel code is the xfuser USP path of the unrelated Wan
video DiT, distributed/xdit_context_parallel.py:102-126: all-to-all around attention, all-gather of the sequence at the end).  The same
decomposition here, built for NVLink 5 / NVSwitch instead of NCCL calls:

  * every rank owns a contiguous slice of the joint [text; image] token rows and runs LN+modulate, the projections and the MLPs on its rows only;
  * attention needs all rows but only 24 / N heads per rank.  The two all-to-alls around it are NOT separate passes: the QKV GEMM's epilogue
    (per-head RMSNorm + RoPE already happen there) stores each head group straight into the q / k / v buffer of the rank that owns it, and the
    attention kernel's epilogue stores each output row straight into the attention buffer of the rank that owns the row -- peer-GPU memory mapped
    into the process through torch's symmetric memory (`pe_gemm_seg.{q,k,v}_route`, `pe_attention_fwd_routed`): P2P stores over NVLink issued
    from the tcgen05 kernels themselves, overlapped tile by tile with the math;
  * two device-side barriers per block (`_SymmetricMemory.barrier`, ~7 us) order the stores against their consumers; no NCCL call in the loop.

Every GEMM row and every (head, 256-row) attention item is computed by the same instructions as on one GPU, so the result is bit-identical to
the single-GPU forward (tested on 2 GPUs).  bf16, N in {1, 2, 3, 4, 6, 8} (divisors of 24 up to the route-table size).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from . import native as nv

DIM, NUM_HEADS, HEAD_DIM = 3072, 24, 128


class UlyssesContext:
    """Symmetric workspaces and the row partition for one process group."""

    def __init__(self, group=None, device=None):
        import torch.distributed._symmetric_memory as symm
        self.symm = symm
        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.N = dist.get_rank(self.group), dist.get_world_size(self.group)
        if NUM_HEADS % self.N or self.N > 8:
            raise ValueError(f"the head-parallel mode needs a group size that divides {NUM_HEADS} and is <= 8 (got {self.N})")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.Hn = NUM_HEADS // self.N
        self._ws: Dict[Tuple[int, int], "UlyssesWorkspace"] = {}

    def bounds(self, S: int) -> List[int]:
        """Row partition of the joint sequence: equal chunks rounded to 128 rows (GEMM m-tile height), the last rank takes the tail."""
        per = math.ceil(S / self.N / 128) * 128
        return [min(r * per, S) for r in range(self.N)] + [S]

    def symm_alloc(self, shape):
        t = self.symm.empty(*shape, dtype=torch.bfloat16, device=self.device)
        h = self.symm.rendezvous(t, self.group)
        return t, h

    def workspace(self, S_img: int, T: int) -> "UlyssesWorkspace":
        key = (S_img, T)
        if key not in self._ws:
            if len(self._ws) >= 4:
                self._ws.pop(next(iter(self._ws)))
            self._ws[key] = UlyssesWorkspace(self, S_img, T)
        return self._ws[key]


class UlyssesWorkspace:
    def __init__(self, ctx: UlyssesContext, S_img: int, T: int):
        S = S_img + T
        b = ctx.bounds(S)
        self.S, self.T, self.S_img, self.b = S, T, S_img, b
        self.lo, self.hi = b[ctx.rank], b[ctx.rank + 1]
        n_loc, n_max = self.hi - self.lo, max(b[i + 1] - b[i] for i in range(ctx.N))
        bf = dict(dtype=torch.bfloat16, device=ctx.device)
        C = ctx.Hn * HEAD_DIM
        # peer-visible buffers: q / k / v hold ALL rows of the local heads, att holds the local rows of ALL heads
        self.qkv, self.h_qkv = ctx.symm_alloc((3, S, C))
        self.att, self.h_att = ctx.symm_alloc((max(n_max, 1), DIM))
        self.out_tok, self.h_out = ctx.symm_alloc((S_img, 64))
        es = 2
        qkv_ptrs, att_ptrs, out_ptrs = list(self.h_qkv.buffer_ptrs), list(self.h_att.buffer_ptrs), list(self.h_out.buffer_ptrs)
        # QKV routes for a segment whose first row is global row g0: peer p's q / k / v base + g0 * C
        self.qkv_base = [(qkv_ptrs[p], qkv_ptrs[p] + S * C * es, qkv_ptrs[p] + 2 * S * C * es) for p in range(ctx.N)]
        # attention output routes: owner i gets rows [b[i], b[i+1]) at its local row (row - b[i]), columns of MY heads
        self.att_routes = [att_ptrs[i] - b[i] * DIM * es + ctx.rank * C * es for i in range(ctx.N)]
        self.route_end = [b[i + 1] for i in range(ctx.N)]
        self.out_ptrs = out_ptrs
        self.x = torch.empty(max(n_loc, 1), DIM, **bf)
        self.xhat = torch.empty(max(n_loc, 1), DIM, **bf)
        self.h = torch.empty(max(n_loc, 1), 4 * DIM, **bf)
        self.x_full = torch.empty(S, DIM, **bf)                    # input stage (img_in / txt_in are cheap: computed on every rank)
        self.tok = torch.empty(S_img, 64, **bf)
        self.txt_n = torch.empty(T, 3584, **bf)


def _segments(ws: UlyssesWorkspace):
    """(text part, image part) of this rank's rows as (global lo, global hi) pairs; either may be empty."""
    t = (ws.lo, min(ws.hi, ws.T)) if ws.lo < ws.T else None
    i = (max(ws.lo, ws.T), ws.hi) if ws.hi > ws.T else None
    return t, i


def run_block_sp(eng, ctx: UlyssesContext, ws: UlyssesWorkspace, i: int, mods: torch.Tensor, rope: torch.Tensor) -> None:
    """One double-stream block on this rank's rows (DiTEngine.run_block with the attention head-parallel across the group)."""
    nat, blk = eng.nat, eng.dit.transformer_blocks[i]
    a = blk.attn
    flags = nv.GEMM_FLAG_CTA_PAIR if eng.use_cta_pair else 0
    mi, mt = mods[0], mods[1]
    D, C = DIM, ctx.Hn * HEAD_DIM
    seg_t, seg_i = _segments(ws)
    n_loc = ws.hi - ws.lo
    x, xhat, hbuf = ws.x[:n_loc], ws.xhat[:n_loc], ws.h[:n_loc]
    n_txt = (seg_t[1] - seg_t[0]) if seg_t else 0                   # local rows [0, n_txt) are text, the rest image

    def ln(which):
        o = 3 * D * which
        nat.tag = "ln_mod"
        nat.layernorm_modulate2(x, xhat, n_txt, mt[o:o + D], mt[o + D:o + 2 * D], mi[o:o + D], mi[o + D:o + 2 * D])

    def segs(build):
        out = []
        if seg_i:
            out.append(build(False, slice(n_txt, n_loc), seg_i[0]))
        if seg_t:
            out.append(build(True, slice(0, n_txt), seg_t[0]))
        return out

    ln(0)
    (wq_i, wq_t), (bq_i, bq_t) = eng.qkv_w[i], eng.qkv_b[i]

    def qkv_seg(is_txt, rows, g0):
        off = g0 * C * 2
        routes = tuple([ws.qkv_base[p][w] + off for p in range(ctx.N)] for w in range(3))
        mine = ws.qkv[:, g0:g0 + (rows.stop - rows.start)]
        return dict(a=xhat[rows], w=wq_t if is_txt else wq_i, bias=bq_t if is_txt else bq_i, out=mine[0], out_k=mine[1], out_v=mine[2],
                    norm_q_w=(a.norm_added_q if is_txt else a.norm_q).weight, norm_k_w=(a.norm_added_k if is_txt else a.norm_k).weight,
                    rope=rope[g0:g0 + (rows.stop - rows.start)], routes=routes)
    nat.tag = "gemm_qkv"
    nat.gemm(segs(qkv_seg), 3 * D, D, nv.EPI_QKV_NORM_ROPE, flags)       # epilogue stores every head group into its owner's q / k / v (NVLink P2P)
    ws.h_qkv.barrier(channel=0)
    nat.tag = "attention"
    nat.attention_routed(ws.qkv[0], ws.qkv[1], ws.qkv[2], ctx.Hn, 1.0 / math.sqrt(HEAD_DIM), ws.route_end, ws.att_routes, DIM, eng.attn_flags & 3)
    ws.h_att.barrier(channel=1)
    att = ws.att[:n_loc]
    nat.tag = "gemm_out"
    nat.gemm(segs(lambda is_txt, rows, g0: dict(a=att[rows], w=(a.to_add_out if is_txt else a.to_out[0]).weight, bias=(a.to_add_out if is_txt else a.to_out[0]).bias,
                                                out=x[rows], gate=(mt if is_txt else mi)[2 * D:3 * D])), D, D, nv.EPI_GATE_RESIDUAL, flags)
    ln(1)
    im, tm = blk.img_mlp.net, blk.txt_mlp.net
    nat.tag = "gemm_up"
    nat.gemm(segs(lambda is_txt, rows, g0: dict(a=xhat[rows], w=(tm if is_txt else im)[0].proj.weight, bias=(tm if is_txt else im)[0].proj.bias, out=hbuf[rows])),
             4 * D, D, nv.EPI_BIAS_GELU_SIGMOID, flags)
    nat.tag = "gemm_down"
    nat.gemm(segs(lambda is_txt, rows, g0: dict(a=hbuf[rows], w=(tm if is_txt else im)[2].weight, bias=(tm if is_txt else im)[2].bias, out=x[rows],
                                                gate=(mt if is_txt else mi)[5 * D:6 * D])), D, 4 * D, nv.EPI_GATE_RESIDUAL, flags)


def forward_sp(eng, ctx: UlyssesContext, latents_list: Sequence[torch.Tensor], timestep_bf16: torch.Tensor, prompt_emb: torch.Tensor,
               out_latents: torch.Tensor, t_key: Optional[float] = None, rope_sampling: bool = False) -> torch.Tensor:
    """DiTEngine.forward across the group: same arguments on every rank, the full velocity on every rank."""
    nat, dit = eng.nat, eng.dit
    T = prompt_emb.shape[0]
    shapes = [(1, l.shape[-2] // 2, l.shape[-1] // 2) for l in latents_list]
    S_img = sum(h * w for _, h, w in shapes)
    ws = ctx.workspace(S_img, T)
    rope = eng.rope(shapes, T, rope_sampling)
    temb, mods, out_mod = eng.conditioning(timestep_bf16, t_key)
    off = 0
    for l, (_, h, w) in zip(latents_list, shapes):
        nat.patchify(l.reshape(16, l.shape[-2], l.shape[-1]), ws.tok[off:off + h * w])
        off += h * w
    nat.gemm([dict(a=ws.tok, w=dit.img_in.weight, bias=dit.img_in.bias, out=ws.x_full[T:])], DIM, 64, nv.EPI_BIAS)
    nat.rmsnorm(prompt_emb, ws.txt_n, dit.txt_norm.weight, dit.txt_norm.eps)
    nat.gemm([dict(a=ws.txt_n, w=dit.txt_in.weight, bias=dit.txt_in.bias, out=ws.x_full[:T])], DIM, 3584, nv.EPI_BIAS)
    n_loc = ws.hi - ws.lo
    ws.x[:n_loc].copy_(ws.x_full[ws.lo:ws.hi])
    for i in range(len(dit.transformer_blocks)):
        run_block_sp(eng, ctx, ws, i, mods[i], rope)
    # norm_out + proj_out on the local noise tokens, written into every rank's token buffer (the final all-gather, 512 KiB in total)
    n0 = shapes[0][1] * shapes[0][2]
    lo, hi = max(ws.lo, T), min(ws.hi, T + n0)
    if hi > lo:
        rows = slice(lo - ws.lo, hi - ws.lo)
        nat.layernorm_modulate(ws.x[rows], ws.xhat[rows], out_mod[0, DIM:], out_mod[0, :DIM])
        tok = torch.empty(hi - lo, 64, dtype=torch.bfloat16, device=ctx.device)
        nat.gemm([dict(a=ws.xhat[rows], w=dit.proj_out.weight, bias=dit.proj_out.bias, out=tok)], 64, DIM, nv.EPI_BIAS)
        for p in range(ctx.N):
            ws.h_out.get_buffer(p, (S_img, 64), torch.bfloat16)[lo - T:hi - T].copy_(tok)
    ws.h_out.barrier(channel=2)
    nat.unpatchify(ws.out_tok[:n0], out_latents.reshape(16, out_latents.shape[-2], out_latents.shape[-1]))
    ws.h_out.barrier(channel=3)                  # nobody overwrites out_tok (next forward) before every rank has read it
    return out_latents
