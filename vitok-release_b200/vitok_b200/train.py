"""Training step of the B200-native AE (BASELINE config 5; reference: scripts/train_vae.py:304-320,371-372).

``AE.forward`` in training mode routes here: one ``torch.autograd.Function`` runs the encoder + decoder forward
with hand-written sm_100a kernels while keeping what the backward pass needs, and its ``backward`` produces the
gradient of every parameter with the kernels of csrc/vtk_train.cu, csrc/vtk_attention_bwd.cu and the tcgen05
GEMM (data gradients against the weights as stored, weight gradients on transposed operands).  PyTorch only owns memory,
streams, the autograd graph edge, the random draw of stochastic depth and -- under DDP (train_vae.py:172) -- the NCCL
gradient all-reduce.

Also here: ``charbonnier_loss`` (train_vae.py:314-320) and ``FusedAdamW`` (train_vae.py:200-208: AdamW with fp32 master
weights and fp32 moments, multi-tensor).

Saved per block (bf16 unless noted): x (block input), h = RMSNorm(x), zraw = [h Wqkv^T | pad | h W1^T], qkv (normed + roped),
a2 = [attention out | silu(g) v | pad], lse (fp32 [M, heads]), y = a2 [out_proj | fc2]^T.
``checkpoint = k`` (ae.py:159-160,202-205,231-233): blocks with ``i % k == 0`` keep only x (and their stochastic-depth draw)
and are re-run from it inside the backward pass -- config 5 at 8 x 1024 tokens per GPU saves ~0.9 GB per block, 40 GB for
44 blocks; with checkpoint = 1 the saved set shrinks to 50 MB per block.
``drop_path_rate > 0`` (ae.py:15-30,65,143-152): decoder block i drops its whole update for an image with probability
``rate * i / (depth - 1)`` and rescales the others by 1 / keep_prob; the draw uses torch's CUDA generator exactly as the
reference does (``keep_prob + torch.rand((B, 1, 1), dtype)`` floored), one draw per decoder block in block order.
Any ``patch_mask`` is accepted under the sdpa backend: tokens are permuted so that each image's valid tokens come first
(the model is equivariant to token order: positions travel in row_idx / col_idx), which turns every mask into the prefix
mask the attention backward understands; the output is permuted back.
"""
from __future__ import annotations

import ctypes
from typing import Dict, List, Optional

import torch

from . import _lib

BF = torch.bfloat16


def _check(rc):
    _lib.check(rc)


def _linear(a: torch.Tensor, lda: int, w: torch.Tensor, bias: Optional[torch.Tensor], M: int, N: int, K: int,
            out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out [M, N] = a[:, :K] @ w[:N, :K]^T (+ bias) on the tcgen05 GEMM.  ``a`` / ``w`` may have a row pitch > K."""
    if out is None:
        out = torch.empty(M, N, dtype=BF, device=a.device)
    _check(_lib.load().vtk_linear_bf16(a.data_ptr(), lda, w.data_ptr(), w.stride(0), _lib.ptr(bias), out.data_ptr(), out.stride(0),
                                       M, N, K, _lib.stream_ptr()))
    return out


def _wgrad(dy: torch.Tensor, ld_dy: int, n1: int, x: torch.Tensor, ld_x: int, n2: int, M: int,
           out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """dW [n1, n2] = dy[:, :n1]^T @ x[:, :n2] on the transposed-operand GEMM (both tiles read MN-major; no transpose pass)."""
    if out is None:
        out = torch.empty(n1, n2, dtype=BF, device=dy.device)
    _check(_lib.load().vtk_linear_tn_bf16(dy.data_ptr(), ld_dy, x.data_ptr(), ld_x, out.data_ptr(), n2, n1, n2, M, _lib.stream_ptr()))
    return out


def _dgrad(dy: torch.Tensor, ld_dy: int, w: torch.Tensor, M: int, n_in: int, n_out: int, out: Optional[torch.Tensor] = None,
           accumulate: bool = False) -> torch.Tensor:
    """dX [M, n_in] (+)= dy[:, :n_out] @ w for a weight w [n_out, n_in] used as stored (B operand read MN-major).
    ``accumulate``: the GEMM epilogue adds to what ``out`` holds (fp32 sum, one bf16 rounding) instead of a separate add pass."""
    if out is None:
        out = torch.empty(M, n_in, dtype=BF, device=dy.device)
    _check(_lib.load().vtk_linear_nn_acc_bf16(dy.data_ptr(), ld_dy, w.data_ptr(), w.stride(0), out.data_ptr(), out.stride(0), M, n_in,
                                              n_out, 1 if accumulate else 0, _lib.stream_ptr()))
    return out


def _linear2(a: torch.Tensor, lda: int, w0: torch.Tensor, w1: torch.Tensor, M: int, N: int) -> torch.Tensor:
    """out [M, N] = a[:, :K0] @ w0^T + a[:, K0:K0+K1] @ w1^T: out_proj(attn) + fc2(act) in ONE accumulator, reading both weights as
    stored (no packed [out_proj | fc2] copy: the optimizer changes them every step)."""
    out = torch.empty(M, N, dtype=BF, device=a.device)
    _check(_lib.load().vtk_linear2_bf16(a.data_ptr(), lda, w0.data_ptr(), w0.stride(0), w1.data_ptr(), w1.stride(0), out.data_ptr(), N,
                                        M, N, w0.shape[1], w1.shape[1], _lib.stream_ptr()))
    return out


def _colsum(x: torch.Tensor, ld: int, M: int, C: int) -> torch.Tensor:
    out = torch.zeros(C, dtype=torch.float32, device=x.device)
    _check(_lib.load().vtk_colsum(x.data_ptr(), ld, out.data_ptr(), M, C, _lib.stream_ptr()))
    return out


class _GradSync:
    """Data-parallel gradient averaging overlapped with the backward loop.

    torch DDP (scripts/train_vae.py:172) also works with this model, but because ``model(batch)`` is ONE autograd node
    all gradients reach DDP's reducer together, after the last backward kernel, and the 10 GB all-reduce of the 5B
    model runs un-overlapped.  With ``enable_grad_sync(model)`` the backward loop itself issues an NCCL all-reduce
    (AVG) for each block's weight gradients as soon as its wgrad GEMMs are queued -- NCCL runs them on its own stream
    while the next block's kernels execute -- and the remaining small tensors go out as one flat bucket at the end.
    The gradients autograd hands to the optimizer are already averaged.
    """

    def __init__(self, group=None, min_numel: int = 1 << 16):
        import torch.distributed as dist
        self.dist, self.group, self.min_numel = dist, group, min_numel
        self.world = dist.get_world_size(group)
        self.avg = dist.get_backend(group) == "nccl"     # ReduceOp.AVG is NCCL-only; elsewhere (gloo in the CPU tests): SUM, then scale
        self.pending, self.small = [], []

    def _all_reduce(self, t: torch.Tensor, async_op: bool):
        if self.avg:
            return self.dist.all_reduce(t, op=self.dist.ReduceOp.AVG, group=self.group, async_op=async_op)
        w = self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group, async_op=async_op)
        return (w, t) if async_op else t.div_(self.world)

    def reduce(self, t: torch.Tensor) -> None:
        if self.world == 1:
            return
        if t.numel() >= self.min_numel and t.is_contiguous():
            self.pending.append(self._all_reduce(t, True))
        else:
            self.small.append(t)

    def finish(self, named: Dict[str, torch.Tensor]) -> None:
        """Average everything not yet reduced (small or strided tensors) as one flat fp32 bucket, then make the
        current stream wait for every outstanding all-reduce."""
        if self.world == 1:
            return
        seen = {id(t) for t in self.small}
        rest = list(self.small) + [t for t in named.values() if t is not None and not getattr(t, "_vtk_reduced", False)
                                   and id(t) not in seen]
        if rest:
            flat = torch.cat([t.reshape(-1).float() for t in rest])
            self._all_reduce(flat, False)
            off = 0
            for t in rest:
                t.copy_(flat[off:off + t.numel()].view_as(t))
                off += t.numel()
        for w in self.pending:
            if isinstance(w, tuple):      # SUM fallback: scale once the sum has arrived
                w[0].wait()
                w[1].div_(self.world)
            else:
                w.wait()
        self.pending, self.small = [], []


def enable_grad_sync(model, process_group=None, broadcast_parameters: bool = True, reserve_sms: Optional[int] = None):
    """Turn on overlapped data-parallel gradient averaging for ``model`` (see _GradSync).  Call once after
    ``torch.distributed.init_process_group``; do NOT also wrap the model in DistributedDataParallel.  With
    ``broadcast_parameters`` the parameters of rank 0 are copied to every rank first (what DDP's constructor does).
    ``reserve_sms``: SMs the persistent GEMM / attention kernels leave free for NCCL's all-reduce CTAs, which run next to the
    backward kernels (include/vitok_b200.h: vtk_set_flag "reserve_sms"); None keeps the process-wide setting."""
    import os
    import torch.distributed as dist
    if reserve_sms is not None:
        _lib.set_flag("reserve_sms", int(reserve_sms))
    if (process_group is None and dist.get_world_size() > 1 and dist.get_backend() == "nccl"
            and os.environ.get("VTK_GRAD_SYNC_HIPRIO", "1") != "0"):
        # The per-block all-reduces run NEXT TO the backward kernels: give them their own communicator on a high-priority stream so
        # that their CTAs are placed as soon as SMs free up instead of queueing behind the next GEMM's grid (VTK_GRAD_SYNC_HIPRIO=0:
        # the default process group).
        try:
            opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
            process_group = dist.new_group(backend="nccl", pg_options=opts)
        except Exception:     # older torch builds without the option: stay on the default group
            process_group = None
    if broadcast_parameters and dist.get_world_size(process_group) > 1:
        with torch.no_grad():
            for p in model.parameters():
                dist.broadcast(p.data, src=dist.get_global_rank(process_group, 0) if process_group is not None else 0,
                               group=process_group)
                torch.autograd.graph.increment_version(p)
    model._grad_sync_group = (process_group,)
    return model


class _SideWeights:
    """bf16 weights of one side (encoder / decoder) as the training kernels read them, taken straight from the parameters
    every step (the optimizer has just changed them) -- nothing is repacked, transposed or concatenated:
      * qkv_proj.weight [3D, D] and fc1.weight [2Hf, D] feed two forward GEMMs that write the column ranges [0, 3D) and
        [qp, qp + 2Hf) of zraw; their data gradients come from the same tensors read MN-major (vtk_linear_nn_acc_bf16, the
        second one accumulating into the first one's output) and their weight gradients from the transposed-operand GEMM
        (vtk_linear_tn_bf16), in the parameters' own row order;
      * out_proj.weight [D, D] and fc2.weight [D, Hf] feed ONE forward GEMM over the concatenated activations
        (vtk_linear2_bf16: two B tensor maps, one accumulator).
    bf16 parameters are used in place (no copy); fp32 parameters (the reference recipe under autocast) are cast once per step.
    """

    def __init__(self, model, side: int):
        from .models.ae import _Scale
        sd = {s[0]: s for s in model._sides()}[side]
        _, lin_a, lin_b, blocks, width, heads = sd
        dev = lin_a.weight.device
        D = width

        def b16(t):
            return t.detach().to(BF).contiguous()

        self.blocks = []
        self.drop = []
        for blk in blocks:
            gamma = b16(blk.layer_scale.gamma) if isinstance(blk.layer_scale, _Scale) else torch.ones(D, dtype=BF, device=dev)
            self.blocks.append((b16(blk.attn.qkv_proj.weight), b16(blk.ffn.fc1.weight), b16(blk.attn.out_proj.weight),
                                b16(blk.ffn.fc2.weight), b16(blk.norm1.weight), b16(blk.attn.norm_q.weight),
                                b16(blk.attn.norm_k.weight), gamma))
            self.drop.append(float(getattr(blk, "drop_path_rate", 0.0) or 0.0))
        self.wa, self.ba, self.wb, self.bb = b16(lin_a.weight), b16(lin_a.bias), b16(lin_b.weight), b16(lin_b.bias)


def _block_forward(lib, blk, x: torch.Tensor, rope, kv, pf, m8, B: int, N: int, dims, window: int, keep, keep_prob: float):
    """One transformer block (ae.py:55-65) with everything the backward needs: returns (x_new, (h, zraw, qkv, a2, lse, y))."""
    M, D, heads, d, Hf, qp, ZP, kp = dims
    wqkv, w1, w_o, w_2, n1, nq, nk, gamma = blk
    st = _lib.stream_ptr()
    dev = x.device
    h = torch.empty(M, D, dtype=BF, device=dev)
    _check(lib.vtk_rmsnorm_bf16(x.data_ptr(), D, n1.data_ptr(), h.data_ptr(), D, M, D, 1e-6, st))
    zraw = torch.empty(M, ZP, dtype=BF, device=dev)       # [q | k | v | pad | fc1 value | fc1 gate]
    if qp > 3 * D:                                        # (3D is a multiple of 256 for every preset: no pad, no fill)
        zraw[:, 3 * D:qp].zero_()
    _linear(h, D, wqkv, None, M, 3 * D, D, out=zraw[:, :3 * D])
    _linear(h, D, w1, None, M, 2 * Hf, D, out=zraw[:, qp:])
    qkv = torch.empty(M, 3 * D, dtype=BF, device=dev)
    _check(lib.vtk_qk_norm_rope_fwd(zraw.data_ptr(), ZP, nq.data_ptr(), nk.data_ptr(), rope.data_ptr(), qkv.data_ptr(), 3 * D,
                                    M, heads, d, 1e-6, st))
    a2 = torch.empty(M, kp, dtype=BF, device=dev)
    if kp > D + Hf:
        a2[:, D + Hf:].zero_()
    lse = torch.empty(M, heads, dtype=torch.float32, device=dev)
    base = qkv.data_ptr()
    _check(lib.vtk_attention_bf16(base, base + 2 * D, base + 4 * D, 3 * D, a2.data_ptr(), kp, _lib.ptr(kv),
                                  _lib.ptr(m8), _lib.ptr(pf), B, N, heads, d, 1 if m8 is not None else 0, window,
                                  lse.data_ptr(), st))
    _check(lib.vtk_swiglu_fwd(zraw.data_ptr(), ZP, qp, a2.data_ptr() + 2 * D, kp, M, Hf, 1, st))
    y = _linear2(a2, kp, w_o, w_2, M, D)
    x_new = torch.empty(M, D, dtype=BF, device=dev)
    _check(lib.vtk_resid_fwd_dp(x.data_ptr(), y.data_ptr(), gamma.data_ptr(), x_new.data_ptr(), M, D, _lib.ptr(keep), N,
                                float(keep_prob), st))
    return x_new, (h, zraw, qkv, a2, lse, y)


def _side_forward(lib, model, sw: _SideWeights, side: int, xin: torch.Tensor, row, col, m8, B: int, N: int, saved: Dict):
    width = model.encoder_width if side == 0 else model.decoder_width
    heads = model.encoder_heads if side == 0 else model.decoder_heads
    d = width // heads
    M, D = B * N, width
    from .models.ae import _ffn_hidden
    Hf = _ffn_hidden(width, model.mlp_factor)
    qp = ((3 * D + 255) // 256) * 256
    ZP = qp + 2 * Hf
    kp = (D + Hf + 63) // 64 * 64
    dims = (M, D, heads, d, Hf, qp, ZP, kp)
    st = _lib.stream_ptr()
    dev = xin.device
    cin = xin.shape[-1]
    x = _linear(xin, cin, sw.wa, sw.ba, M, D, cin)
    rope = kv = pf = None
    if sw.blocks:
        inv = (1.0 / (model.rope_theta ** (torch.arange(0, d // 2, 2).float() / (d // 2)))).to(dev)
        rope = torch.zeros((M + 31) // 32 * 32, 2 * d, dtype=BF, device=dev)
        _check(lib.vtk_rope_table(row.data_ptr(), col.data_ptr(), inv.data_ptr(), rope.data_ptr(), M, d, st))
        if m8 is not None:
            kv, pf = _lib.kv_len(m8.view(B, N))     # train_forward has made every mask a prefix mask
    window = int(model.sw) if (model.sw and m8 is None) else -1
    ck = int(getattr(model, "checkpoint", 0) or 0)
    layers = []
    for li, blk in enumerate(sw.blocks):
        keep, keep_prob = None, 1.0
        rate = sw.drop[li]
        if rate > 0.0:
            # stochastic depth (ae.py:15-30): the reference's own draw -- keep_prob + rand((B, 1, 1), dtype) floored -- from
            # torch's CUDA generator, one per block in block order, so a seeded run drops the same images as the reference
            keep_prob = 1.0 - rate
            keep = (keep_prob + torch.rand((B, 1, 1), dtype=BF, device=dev)).floor_().float().reshape(B).contiguous()
        x_new, inner = _block_forward(lib, blk, x, rope, kv, pf, m8, B, N, dims, window, keep, keep_prob)
        # activation checkpointing (ae.py:159-160): block i with i % checkpoint == 0 keeps its input only
        layers.append((x, None if (ck > 0 and li % ck == 0) else inner, keep, keep_prob))
        x = x_new
    saved.update(dict(layers=layers, x_final=x, rope=rope, kv=kv, pf=pf, m8=m8, window=window, xin=xin, dims=dims))
    return x


def _side_backward(lib, model, sw: _SideWeights, side: int, dx: torch.Tensor, saved: Dict, B: int, N: int, need_dxin: bool,
                   sync: Optional[_GradSync] = None):
    """dx = gradient w.r.t. the output of the last block.  Returns (block grads, dW_a, db_a, d_input)."""
    dims = saved["dims"]
    M, D, heads, d, Hf, qp, ZP, kp = dims
    st = _lib.stream_ptr()
    dev = dx.device
    rope, kv, window = saved["rope"], saved["kv"], saved["window"]
    grads: List[Dict[str, torch.Tensor]] = []
    for li in range(len(sw.blocks) - 1, -1, -1):
        wqkv, w1, w_o, w_2, n1, nq, nk, gamma = sw.blocks[li]
        x, inner, keep, keep_prob = saved["layers"][li]
        if inner is None:      # checkpointed block: re-run its forward from x (same stochastic-depth draw)
            _, inner = _block_forward(lib, sw.blocks[li], x, rope, kv, saved["pf"], saved["m8"], B, N, dims, window, keep, keep_prob)
        h, zraw, qkv, a2, lse, y = inner
        saved["layers"][li] = None                                                 # free this block's activations as we go
        dy = torch.empty(M, D, dtype=BF, device=dev)
        dgamma = torch.zeros(D, dtype=torch.float32, device=dev)
        _check(lib.vtk_resid_bwd_dp(dx.data_ptr(), y.data_ptr(), gamma.data_ptr(), dy.data_ptr(), dgamma.data_ptr(), M, D,
                                    _lib.ptr(keep), N, float(keep_prob), st))
        # the four weight gradients of the block live in ONE flat buffer so that they leave as one all-reduce
        n_qkv, n_fc1, n_o, n_f2 = 3 * D * D, 2 * Hf * D, D * D, D * Hf
        flat = torch.empty(n_qkv + n_fc1 + n_o + n_f2, dtype=BF, device=dev)
        dw_qkv = flat[:n_qkv].view(3 * D, D)
        dw_fc1 = flat[n_qkv:n_qkv + n_fc1].view(2 * Hf, D)
        dw_o = flat[n_qkv + n_fc1:n_qkv + n_fc1 + n_o].view(D, D)
        dw_f2 = flat[n_qkv + n_fc1 + n_o:].view(D, Hf)
        da2 = torch.empty(M, kp, dtype=BF, device=dev)                            # d[attn | act] = dy @ [out_proj | fc2]
        _dgrad(dy, D, w_o, M, D, D, out=da2[:, :D])
        _dgrad(dy, D, w_2, M, Hf, D, out=da2[:, D:D + Hf])
        _wgrad(dy, D, D, a2, kp, D, M, out=dw_o)                                  # out_proj.weight grad [D, D] = dy^T attn
        _wgrad(dy, D, D, a2[:, D:], kp, Hf, M, out=dw_f2)                         # fc2.weight grad [D, Hf] = dy^T act
        dz = torch.empty(M, ZP, dtype=BF, device=dev)
        if qp > 3 * D:
            dz[:, 3 * D:qp].zero_()
        _check(lib.vtk_swiglu_bwd(da2.data_ptr() + 2 * D, kp, zraw.data_ptr(), ZP, qp, dz.data_ptr(), ZP, M, Hf, 1, st))
        delta = torch.empty(M, heads, dtype=torch.float32, device=dev)
        _check(lib.vtk_attn_delta(a2.data_ptr(), kp, da2.data_ptr(), kp, delta.data_ptr(), M, heads, d, st))
        qb, zb = qkv.data_ptr(), dz.data_ptr()
        _check(lib.vtk_attention_bwd_bf16(qb, qb + 2 * D, qb + 4 * D, 3 * D, da2.data_ptr(), kp, lse.data_ptr(), delta.data_ptr(),
                                          zb, zb + 2 * D, zb + 4 * D, ZP, _lib.ptr(kv), B, N, heads, d,
                                          1 if kv is not None else 0, window, st))
        dwqk = torch.zeros(2, d, dtype=torch.float32, device=dev)
        _check(lib.vtk_qk_norm_rope_bwd(dz.data_ptr(), ZP, zraw.data_ptr(), ZP, nq.data_ptr(), nk.data_ptr(), rope.data_ptr(),
                                        dwqk.data_ptr(), M, heads, d, 1e-6, st))
        dh = _dgrad(dz, ZP, wqkv, M, D, 3 * D)                                    # dz[:, :3D] @ Wqkv ...
        _dgrad(dz[:, qp:], ZP, w1, M, D, 2 * Hf, out=dh, accumulate=True)         # ... + dz[:, qp:] @ W1, added in the GEMM epilogue
        _wgrad(dz, ZP, 3 * D, h, D, D, M, out=dw_qkv)                             # qkv_proj.weight grad [3D, D]
        _wgrad(dz[:, qp:], ZP, 2 * Hf, h, D, D, M, out=dw_fc1)                    # fc1.weight grad [2Hf, D], fc1's own row order
        dx_in = torch.empty(M, D, dtype=BF, device=dev)
        dw1 = torch.zeros(D, dtype=torch.float32, device=dev)
        _check(lib.vtk_rmsnorm_bwd(x.data_ptr(), dh.data_ptr(), n1.data_ptr(), dx.data_ptr(), dx_in.data_ptr(), dw1.data_ptr(),
                                   M, D, 1e-6, st))
        bg = dict(
            qkv=dw_qkv, fc1=dw_fc1, out=dw_o, fc2=dw_f2,
            norm1=dw1, norm_q=dwqk[0], norm_k=dwqk[1], gamma=dgamma)
        if sync is not None:          # this block's weight gradients leave for the all-reduce while the next block computes
            sync.reduce(flat)
            for k in ("qkv", "fc1", "out", "fc2"):
                bg[k]._vtk_reduced = True
        grads.append(bg)
        dx = dx_in
    grads.reverse()
    xin = saved["xin"]
    cin = xin.shape[-1]
    dwa = _wgrad(dx, D, D, xin.reshape(M, cin), cin, cin, M)                     # [D, cin]
    dba = _colsum(dx, D, M, D)
    dxin = _dgrad(dx, D, sw.wa, M, cin, D) if need_dxin else None               # [M, cin]
    return grads, dwa, dba, dxin


class AETrainFunction(torch.autograd.Function):
    """patches -> reconstructed patches with gradients for every AE parameter (reference AE.forward, ae.py:245-251)."""

    @staticmethod
    def forward(ctx, model, patches, row, col, mask, *params):
        with _lib.device_of(patches, *params):
            return AETrainFunction._forward(ctx, model, patches, row, col, mask, *params)

    @staticmethod
    def _forward(ctx, model, patches, row, col, mask, *params):
        lib = _lib.load()
        B, N, P = patches.shape
        dev = patches.device
        if (B * N) % 8:
            raise ValueError("vitok_b200 training: B * N must be a multiple of 8 (token count is the K dimension of the wgrad GEMMs)")
        enc_w, dec_w = _SideWeights(model, 0), _SideWeights(model, 1)
        row = row.to(device=dev, dtype=torch.int64).contiguous()
        col = col.to(device=dev, dtype=torch.int64).contiguous()
        m8 = mask.to(dev).bool().contiguous().view(torch.uint8) if mask is not None else None
        xin = (_lib.cast_to_bf16(patches) if patches.dtype == torch.float32 else patches).contiguous().view(B * N, P)
        M, C = B * N, model.channels_per_token
        s_enc, s_dec = {}, {}
        x = _side_forward(lib, model, enc_w, 0, xin, row, col, m8, B, N, s_enc)
        zlin = _linear(x, x.shape[1], enc_w.wb, enc_w.bb, M, C, x.shape[1])
        z = torch.empty(M, C, dtype=BF, device=dev)
        _check(lib.vtk_layernorm_fwd(zlin.data_ptr(), z.data_ptr(), M, C, 1e-6, _lib.stream_ptr()))
        xd = _side_forward(lib, model, dec_w, 1, z, row, col, m8, B, N, s_dec)
        out = _linear(xd, xd.shape[1], dec_w.wb, dec_w.bb, M, P, xd.shape[1])
        # activations live on ctx (not save_for_backward: they are workspace of the backward kernels, freed block by block while it
        # runs); a second backward through the same graph is therefore refused with a clear error in backward()
        ctx.model, ctx.enc_w, ctx.dec_w, ctx.s_enc, ctx.s_dec = model, enc_w, dec_w, s_enc, s_dec
        ctx.zlin, ctx.shape, ctx.names = zlin, (B, N, P, C), [n for n, _ in model.named_parameters()]
        ctx.param_dtypes = [p.dtype for p in params]
        ctx.patches_dtype = patches.dtype
        return out.view(B, N, P)

    @staticmethod
    def backward(ctx, dout):
        if ctx.s_dec is None:
            raise RuntimeError("vitok_b200 training: backward was already run through this forward pass -- its activations are "
                               "released while the backward kernels consume them (retain_graph / double backward is not "
                               "supported; run model(batch) again)")
        with _lib.device_of(dout):
            return AETrainFunction._backward(ctx, dout)

    @staticmethod
    def _backward(ctx, dout):
        lib = _lib.load()
        model, enc_w, dec_w = ctx.model, ctx.enc_w, ctx.dec_w
        B, N, P, C = ctx.shape
        M = B * N
        dout = dout.to(BF).contiguous().view(M, P)
        g: Dict[str, torch.Tensor] = {}
        # decoder: to_pixels, blocks, decoder_embed
        xd = ctx.s_dec["x_final"]
        Dd = xd.shape[1]
        g["to_pixels.weight"] = _wgrad(dout, P, P, xd, Dd, Dd, M)
        g["to_pixels.bias"] = _colsum(dout, P, M, P)
        dx = _dgrad(dout, P, dec_w.wb, M, Dd, P)
        grp = getattr(model, "_grad_sync_group", None)
        sync = _GradSync(grp[0]) if grp is not None else None
        blocks, dwa, dba, dz = _side_backward(lib, model, dec_w, 1, dx, ctx.s_dec, B, N, True, sync)
        g["decoder_embed.weight"], g["decoder_embed.bias"] = dwa, dba
        for i, bg in enumerate(blocks):
            _name_block(g, f"decoder_blocks.{i}.", bg)
        # latent bottleneck: LayerNorm (no affine) + to_code
        dzlin = torch.empty(M, C, dtype=BF, device=dout.device)
        _check(lib.vtk_layernorm_bwd(ctx.zlin.data_ptr(), dz.data_ptr(), dzlin.data_ptr(), M, C, 1e-6, _lib.stream_ptr()))
        xe = ctx.s_enc["x_final"]
        De = xe.shape[1]
        g["to_code.weight"] = _wgrad(dzlin, C, C, xe, De, De, M)
        g["to_code.bias"] = _colsum(dzlin, C, M, C)
        dx = _dgrad(dzlin, C, enc_w.wb, M, De, C)
        need_dpatches = bool(ctx.needs_input_grad[1])
        blocks, dwa, dba, dpatches = _side_backward(lib, model, enc_w, 0, dx, ctx.s_enc, B, N, need_dpatches, sync)
        g["patch_embed.weight"], g["patch_embed.bias"] = dwa, dba
        for i, bg in enumerate(blocks):
            _name_block(g, f"encoder_blocks.{i}.", bg)
        if sync is not None:
            sync.finish(g)
        grads = []
        for name, dt in zip(ctx.names, ctx.param_dtypes):
            t = g.get(name)
            grads.append(None if t is None else t.to(dt).contiguous())
        ctx.s_enc = ctx.s_dec = ctx.zlin = None
        if dpatches is not None:
            dpatches = dpatches.view(B, N, P).to(ctx.patches_dtype)
        return (None, dpatches, None, None, None, *grads)


def _name_block(g: Dict[str, torch.Tensor], prefix: str, bg: Dict[str, torch.Tensor]) -> None:
    g[prefix + "attn.qkv_proj.weight"] = bg["qkv"]
    g[prefix + "attn.out_proj.weight"] = bg["out"]
    g[prefix + "ffn.fc1.weight"] = bg["fc1"]
    g[prefix + "ffn.fc2.weight"] = bg["fc2"]
    g[prefix + "norm1.weight"] = bg["norm1"]
    g[prefix + "attn.norm_q.weight"] = bg["norm_q"]
    g[prefix + "attn.norm_k.weight"] = bg["norm_k"]
    g[prefix + "layer_scale.gamma"] = bg["gamma"]


def _prefix_permutation(mask: torch.Tensor):
    """(perm, inv) [B, N] int64 such that mask.gather(1, perm) is a prefix mask (valid tokens first, order kept) and
    x.gather(1, perm).gather(1, inv) == x; None when the mask already is a prefix mask (what patchify emits, ops.py:262-263)."""
    m = mask.bool()
    n_valid = m.sum(dim=1, keepdim=True)
    ar = torch.arange(m.shape[1], device=m.device)[None, :]
    if bool((m == (ar < n_valid)).all()):
        return None
    perm = torch.sort((~m).to(torch.int8), dim=1, stable=True).indices
    inv = torch.empty_like(perm)
    inv.scatter_(1, perm, ar.expand_as(perm).contiguous())
    return perm, inv


def train_forward(model, patch_dict: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """AE.forward in training mode (ae.py:245-251): same dict contract as encode -> decode."""
    if not (model.is_encoder and model.is_decoder):
        raise NotImplementedError("vitok_b200: the training step needs both halves (encoder=True, decoder=True)")
    patches = patch_dict["patches"]
    if not patches.is_cuda:
        raise RuntimeError("vitok_b200.AE: inputs must be CUDA tensors (there is no CPU path)")
    mask = patch_dict.get("patch_mask") if model.attn_backend == "sdpa" else None
    params = [p for _, p in model.named_parameters()]
    row, col = patch_dict["row_idx"], patch_dict["col_idx"]
    pp = _prefix_permutation(mask.to(patches.device)) if mask is not None else None
    if pp is None:
        out = AETrainFunction.apply(model, patches, row, col, mask, *params)
    else:
        # a mask with holes: move every image's valid tokens to the front (pure index plumbing; the model only sees positions
        # through row_idx / col_idx, which travel with the tokens), run the prefix-mask path, and move the outputs back
        perm, inv = pp
        pe = perm[:, :, None].expand(-1, -1, patches.shape[-1])
        out_p = AETrainFunction.apply(model, patches.gather(1, pe), row.to(perm.device).gather(1, perm), col.to(perm.device).gather(1, perm),
                                      mask.to(perm.device).gather(1, perm), *params)
        out = out_p.gather(1, inv[:, :, None].expand(-1, -1, out_p.shape[-1]))
    pdtype = params[0].dtype
    if pdtype == torch.float32 and not torch.is_autocast_enabled():
        out = out.float()
    return {
        "patch_mask": patch_dict.get("patch_mask"), "row_idx": patch_dict.get("row_idx"), "col_idx": patch_dict.get("col_idx"),
        "orig_height": patch_dict.get("orig_height"), "orig_width": patch_dict.get("orig_width"), "patches": out,
    }


class _Charbonnier(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target, mask, eps):
        B, N, P = pred.shape
        dev = pred.device
        with _lib.device_of(pred):
            pb = pred.to(BF).contiguous()
            tb = target.to(device=dev, dtype=BF).contiguous()
            m8 = nv = None
            if mask is not None:
                mb = mask.to(dev).bool().contiguous()
                m8 = mb.view(torch.uint8)
                nv = mb.sum(dim=1).clamp_min(1).to(torch.int32).contiguous()
            loss_sum = torch.zeros(B, dtype=torch.float32, device=dev)
            dpred = torch.empty_like(pb)
            _check(_lib.load().vtk_charbonnier(pb.data_ptr(), tb.data_ptr(), _lib.ptr(m8), _lib.ptr(nv), loss_sum.data_ptr(),
                                               dpred.data_ptr(), B, N, P, float(eps), _lib.stream_ptr()))
        ctx.dpred = dpred
        ctx.in_dtype = pred.dtype
        return loss_sum.mean()

    @staticmethod
    def backward(ctx, gout):
        dpred = ctx.dpred
        if dpred is None:
            raise RuntimeError("vitok_b200.charbonnier_loss: backward was already run through this loss (its gradient buffer is scaled in place)")
        ctx.dpred = None
        with _lib.device_of(dpred):      # d loss / d pred was written by the forward kernel; times the incoming gradient, in place
            gs = gout.detach().to(device=dpred.device, dtype=torch.float32).reshape(1).contiguous()
            _check(_lib.load().vtk_scale_by_dev(dpred.data_ptr(), gs.data_ptr(), dpred.numel(), _lib.stream_ptr()))
        return dpred.to(ctx.in_dtype), None, None, None


def charbonnier_loss(pred: torch.Tensor, target: torch.Tensor, patch_mask: Optional[torch.Tensor] = None,
                     eps: float = 1e-3) -> torch.Tensor:
    """sqrt(diff^2 + eps^2) averaged over pixels, then over each image's valid tokens, then over the batch
    (scripts/train_vae.py:314-320), forward and gradient in one kernel."""
    return _Charbonnier.apply(pred, target, patch_mask, eps)


class FusedAdamW(torch.optim.Optimizer):
    """Multi-tensor AdamW with the reference's optimizer precision (scripts/train_vae.py:200-208 keeps fp32 parameters and fp32
    Adam moments under autocast): every parameter has an fp32 MASTER copy and fp32 ``exp_avg`` / ``exp_avg_sq``; the update runs
    in fp32 on the master and the bf16 model parameter is re-emitted from it in the same pass, so an update smaller than half a
    bf16 ulp of the weight is not lost (at lr 1e-4 that is every weight above ~0.05 and all norm weights).  fp32 model
    parameters are their own master.  torch.optim.AdamW semantics (decoupled weight decay, bias correction); one kernel launch
    per 64 tensors of a param group; 28 bytes of HBM traffic per bf16 parameter.  ``state_dict()`` holds step, master (bf16
    parameters only), exp_avg, exp_avg_sq per parameter."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.load()
        for group in self.param_groups:
            b1, b2 = group["betas"]
            by_dev: Dict[torch.device, list] = {}
            for p in group["params"]:
                if p.grad is None:
                    continue
                if p.dtype not in (BF, torch.float32) or not p.is_cuda:
                    raise RuntimeError("FusedAdamW: parameters must be bf16 or fp32 CUDA tensors")
                if not p.is_contiguous():
                    raise RuntimeError("FusedAdamW: parameters must be contiguous")
                stt = self.state[p]
                if not stt:
                    stt["step"] = 0
                    if p.dtype == BF:
                        stt["master"] = p.detach().float()
                    stt["exp_avg"] = torch.zeros(p.shape, dtype=torch.float32, device=p.device)
                    stt["exp_avg_sq"] = torch.zeros(p.shape, dtype=torch.float32, device=p.device)
                stt["step"] += 1
                gr = p.grad
                if gr.dtype not in (BF, torch.float32):
                    gr = gr.float()
                gr = gr.contiguous()
                by_dev.setdefault(p.device, []).append((p, gr, stt))
            for dev, items in by_dev.items():
                # parameters of a group share lr / betas / eps / weight decay; their step counts agree unless a parameter
                # got its first gradient later -- launch per distinct step
                for step in sorted({it[2]["step"] for it in items}):
                    sel = [it for it in items if it[2]["step"] == step]
                    table = (_lib.AdamwTensor * len(sel))()
                    for e, (p, gr, stt) in zip(table, sel):
                        master = stt.get("master")
                        e.master = (master if master is not None else p).data_ptr()
                        e.p16 = p.data_ptr() if master is not None else None
                        e.g = gr.data_ptr()
                        e.m, e.v = stt["exp_avg"].data_ptr(), stt["exp_avg_sq"].data_ptr()
                        e.n = p.numel()
                        e.weight_decay = float(group["weight_decay"])
                        e.g_is_f32 = 1 if gr.dtype == torch.float32 else 0
                    with torch.cuda.device(dev):
                        _check(lib.vtk_adamw_multi(table, len(sel), float(group["lr"]), float(b1), float(b2), float(group["eps"]),
                                                   int(step), 1.0, _lib.stream_ptr()))
                    for p, _, _ in sel:
                        torch.autograd.graph.increment_version(p)   # the kernel wrote p in place: invalidate packed copies
        return loss
