"""Training step of the B200-native AE (BASELINE config 5; reference: scripts/train_vae.py:304-320,371-372).

``AE.forward`` in training mode routes here: one ``torch.autograd.Function`` runs the encoder + decoder forward
with hand-written sm_100a kernels while keeping what the backward pass needs, and its ``backward`` produces the
gradient of every parameter with the kernels of csrc/vtk_train.cu, csrc/vtk_attention_bwd.cu and the tcgen05
GEMM (dgrad against transposed weight copies, wgrad on transposed activations).  PyTorch only owns memory,
streams, the autograd graph edge and -- under DDP (train_vae.py:172) -- the NCCL gradient all-reduce.

Also here: ``charbonnier_loss`` (train_vae.py:314-320) and ``FusedAdamW`` (train_vae.py:200-208 semantics for
bf16 parameters), both single-kernel.

Saved per block (bf16 unless noted): x (block input), h = RMSNorm(x), zraw = h W_in^T in the packed column order
[q | k | v | pad | (value16, gate16)*], qkv (normed + roped), a2 = [attention out | silu(g) v | pad], lse (fp32
[M, heads]), y = a2 W_out^T.  ``checkpoint`` (ae.py:159-160) is accepted and ignored: at config 5 (8 x 1024
tokens per GPU) the saved set is ~0.9 GB per block, 40 GB for 44 blocks, well inside 180 GB.
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


def _wgrad(dy: torch.Tensor, ld_dy: int, n1: int, x: torch.Tensor, ld_x: int, n2: int, M: int) -> torch.Tensor:
    """dW [n1, n2] = dy[:, :n1]^T @ x[:, :n2] on the transposed-operand GEMM (both tiles read MN-major; no transpose pass)."""
    out = torch.empty(n1, n2, dtype=BF, device=dy.device)
    _check(_lib.load().vtk_linear_tn_bf16(dy.data_ptr(), ld_dy, x.data_ptr(), ld_x, out.data_ptr(), n2, n1, n2, M, _lib.stream_ptr()))
    return out


def _dgrad(dy: torch.Tensor, ld_dy: int, w: torch.Tensor, M: int, n_in: int, n_out: int, out: Optional[torch.Tensor] = None,
           accumulate_into: Optional[torch.Tensor] = None) -> torch.Tensor:
    """dX [M, n_in] = dy[:, :n_out] @ w for a weight w [n_out, n_in] used as stored (B operand read MN-major)."""
    if out is None:
        out = torch.empty(M, n_in, dtype=BF, device=dy.device)
    _check(_lib.load().vtk_linear_nn_bf16(dy.data_ptr(), ld_dy, w.data_ptr(), w.stride(0), out.data_ptr(), out.stride(0), M, n_in, n_out,
                                          _lib.stream_ptr()))
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


def enable_grad_sync(model, process_group=None, broadcast_parameters: bool = True):
    """Turn on overlapped data-parallel gradient averaging for ``model`` (see _GradSync).  Call once after
    ``torch.distributed.init_process_group``; do NOT also wrap the model in DistributedDataParallel.  With
    ``broadcast_parameters`` the parameters of rank 0 are copied to every rank first (what DDP's constructor does)."""
    import torch.distributed as dist
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
    every step (the optimizer has just changed them) -- nothing is repacked or transposed:
      * qkv_proj.weight [3D, D] and fc1.weight [2Hf, D] feed two forward GEMMs that write the column ranges [0, 3D) and
        [qp, qp + 2Hf) of zraw; their data gradients come from the same tensors read MN-major (vtk_linear_nn_bf16) and their
        weight gradients from the transposed-operand GEMM (vtk_linear_tn_bf16), in the parameters' own row order;
      * w_out [D, kp] = [out_proj | fc2 | 0-pad] is the one packed copy (one forward GEMM over the concatenated K).
    """

    def __init__(self, model, side: int):
        from .models.ae import _Scale, pack_w_out
        sd = {s[0]: s for s in model._sides()}[side]
        _, lin_a, lin_b, blocks, width, heads = sd
        dev = lin_a.weight.device
        D = width

        def b16(t):
            return t.detach().to(BF).contiguous()

        self.blocks = []
        for blk in blocks:
            wqkv, w1 = b16(blk.attn.qkv_proj.weight), b16(blk.ffn.fc1.weight)
            w_out = pack_w_out(b16(blk.attn.out_proj.weight), b16(blk.ffn.fc2.weight))
            gamma = b16(blk.layer_scale.gamma) if isinstance(blk.layer_scale, _Scale) else torch.ones(D, dtype=BF, device=dev)
            self.blocks.append((wqkv, w1, w_out, b16(blk.norm1.weight), b16(blk.attn.norm_q.weight), b16(blk.attn.norm_k.weight), gamma))
        self.wa, self.ba, self.wb, self.bb = b16(lin_a.weight), b16(lin_a.bias), b16(lin_b.weight), b16(lin_b.bias)


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
            kv, pf = _lib.kv_len(m8.view(B, N))
            if not bool(pf.all()):
                raise NotImplementedError("vitok_b200 training: patch_mask must be a prefix mask per image (what patchify emits, "
                                          "ops.py:262-263); the attention backward takes key lengths only")
    window = int(model.sw) if (model.sw and m8 is None) else -1
    layers = []
    for blk in sw.blocks:
        wqkv, w1, w_out, n1, nq, nk, gamma = blk
        h = torch.empty(M, D, dtype=BF, device=dev)
        _check(lib.vtk_rmsnorm_bf16(x.data_ptr(), D, n1.data_ptr(), h.data_ptr(), D, M, D, 1e-6, st))
        zraw = torch.empty(M, ZP, dtype=BF, device=dev)       # [q | k | v | pad | fc1 value | fc1 gate]
        if qp > 3 * D:
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
        y = _linear(a2, kp, w_out, None, M, D, D + Hf)
        x_new = torch.empty(M, D, dtype=BF, device=dev)
        _check(lib.vtk_resid_fwd(x.data_ptr(), y.data_ptr(), gamma.data_ptr(), x_new.data_ptr(), M, D, st))
        layers.append((x, h, zraw, qkv, a2, lse, y))
        x = x_new
    saved.update(dict(layers=layers, x_final=x, rope=rope, kv=kv, window=window, xin=xin,
                      dims=(M, D, heads, d, Hf, qp, ZP, kp)))
    return x


def _side_backward(lib, model, sw: _SideWeights, side: int, dx: torch.Tensor, saved: Dict, B: int, N: int, need_dxin: bool,
                   sync: Optional[_GradSync] = None):
    """dx = gradient w.r.t. the output of the last block.  Returns (block grads, dW_a, db_a, d_input)."""
    M, D, heads, d, Hf, qp, ZP, kp = saved["dims"]
    st = _lib.stream_ptr()
    dev = dx.device
    rope, kv, window = saved["rope"], saved["kv"], saved["window"]
    grads: List[Dict[str, torch.Tensor]] = []
    for li in range(len(sw.blocks) - 1, -1, -1):
        wqkv, w1, w_out, n1, nq, nk, gamma = sw.blocks[li]
        x, h, zraw, qkv, a2, lse, y = saved["layers"][li]
        dy = torch.empty(M, D, dtype=BF, device=dev)
        dgamma = torch.zeros(D, dtype=torch.float32, device=dev)
        _check(lib.vtk_resid_bwd(dx.data_ptr(), y.data_ptr(), gamma.data_ptr(), dy.data_ptr(), dgamma.data_ptr(), M, D, st))
        da2 = _dgrad(dy, D, w_out, M, kp, D)                                      # [M, kp]: d[attn | act | pad] = dy @ [out_proj | fc2 | 0]
        dw_o = _wgrad(dy, D, D, a2, kp, D, M)                                     # out_proj.weight grad [D, D] = dy^T attn
        dw_f2 = _wgrad(dy, D, D, a2[:, D:], kp, Hf, M)                            # fc2.weight grad [D, Hf] = dy^T act
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
        dh1 = _dgrad(dz[:, qp:], ZP, w1, M, D, 2 * Hf)                            # ... + dz[:, qp:] @ W1 (added inside rmsnorm_bwd's input below)
        dh.add_(dh1)
        dw_qkv = _wgrad(dz, ZP, 3 * D, h, D, D, M)                                # qkv_proj.weight grad [3D, D]
        dw_fc1 = _wgrad(dz[:, qp:], ZP, 2 * Hf, h, D, D, M)                       # fc1.weight grad [2Hf, D], fc1's own row order
        dx_in = torch.empty(M, D, dtype=BF, device=dev)
        dw1 = torch.zeros(D, dtype=torch.float32, device=dev)
        _check(lib.vtk_rmsnorm_bwd(x.data_ptr(), dh.data_ptr(), n1.data_ptr(), dx.data_ptr(), dx_in.data_ptr(), dw1.data_ptr(),
                                   M, D, 1e-6, st))
        bg = dict(
            qkv=dw_qkv, fc1=dw_fc1, out=dw_o, fc2=dw_f2,
            norm1=dw1, norm_q=dwqk[0], norm_k=dwqk[1], gamma=dgamma)
        if sync is not None:          # this block's big gradients leave for the all-reduce while the next block computes
            for k in ("qkv", "fc1", "out", "fc2"):
                sync.reduce(bg[k])
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
        ctx.model, ctx.enc_w, ctx.dec_w, ctx.s_enc, ctx.s_dec = model, enc_w, dec_w, s_enc, s_dec
        ctx.zlin, ctx.shape, ctx.names = zlin, (B, N, P, C), [n for n, _ in model.named_parameters()]
        ctx.param_dtypes = [p.dtype for p in params]
        return out.view(B, N, P)

    @staticmethod
    def backward(ctx, dout):
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
        blocks, dwa, dba, _ = _side_backward(lib, model, enc_w, 0, dx, ctx.s_enc, B, N, False, sync)
        g["patch_embed.weight"], g["patch_embed.bias"] = dwa, dba
        for i, bg in enumerate(blocks):
            _name_block(g, f"encoder_blocks.{i}.", bg)
        if sync is not None:
            sync.finish(g)
        grads = []
        for name, dt in zip(ctx.names, ctx.param_dtypes):
            t = g.get(name)
            grads.append(None if t is None else t.to(dt).contiguous())
        ctx.s_enc = ctx.s_dec = None
        return (None, None, None, None, None, *grads)


def _name_block(g: Dict[str, torch.Tensor], prefix: str, bg: Dict[str, torch.Tensor]) -> None:
    g[prefix + "attn.qkv_proj.weight"] = bg["qkv"]
    g[prefix + "attn.out_proj.weight"] = bg["out"]
    g[prefix + "ffn.fc1.weight"] = bg["fc1"]
    g[prefix + "ffn.fc2.weight"] = bg["fc2"]
    g[prefix + "norm1.weight"] = bg["norm1"]
    g[prefix + "attn.norm_q.weight"] = bg["norm_q"]
    g[prefix + "attn.norm_k.weight"] = bg["norm_k"]
    g[prefix + "layer_scale.gamma"] = bg["gamma"]


def train_forward(model, patch_dict: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """AE.forward in training mode (ae.py:245-251): same dict contract as encode -> decode."""
    if not (model.is_encoder and model.is_decoder):
        raise NotImplementedError("vitok_b200: the training step needs both halves (encoder=True, decoder=True)")
    patches = patch_dict["patches"]
    if not patches.is_cuda:
        raise RuntimeError("vitok_b200.AE: inputs must be CUDA tensors (there is no CPU path)")
    mask = patch_dict.get("patch_mask") if model.attn_backend == "sdpa" else None
    params = [p for _, p in model.named_parameters()]
    out = AETrainFunction.apply(model, patches, patch_dict["row_idx"], patch_dict["col_idx"], mask, *params)
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
        ctx.save_for_backward(dpred)
        ctx.in_dtype = pred.dtype
        return loss_sum.mean()

    @staticmethod
    def backward(ctx, gout):
        (dpred,) = ctx.saved_tensors
        return (dpred.float() * gout).to(ctx.in_dtype), None, None, None


def charbonnier_loss(pred: torch.Tensor, target: torch.Tensor, patch_mask: Optional[torch.Tensor] = None,
                     eps: float = 1e-3) -> torch.Tensor:
    """sqrt(diff^2 + eps^2) averaged over pixels, then over each image's valid tokens, then over the batch
    (scripts/train_vae.py:314-320), forward and gradient in one kernel."""
    return _Charbonnier.apply(pred, target, patch_mask, eps)


class FusedAdamW(torch.optim.Optimizer):
    """AdamW for bf16 parameters, one fused kernel per tensor (torch.optim.AdamW(fused=True) semantics as used at
    scripts/train_vae.py:200-208: decoupled weight decay, bias correction, bf16 moments, fp32 arithmetic)."""

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
            for p in group["params"]:
                if p.grad is None:
                    continue
                if p.dtype != BF or not p.is_cuda:
                    raise RuntimeError("FusedAdamW: parameters must be bf16 CUDA tensors (model.to('cuda', torch.bfloat16))")
                stt = self.state[p]
                if not stt:
                    stt["step"] = 0
                    stt["exp_avg"] = torch.zeros_like(p)
                    stt["exp_avg_sq"] = torch.zeros_like(p)
                stt["step"] += 1
                gr = p.grad.to(BF).contiguous()
                _check(lib.vtk_adamw_bf16(p.data_ptr(), gr.data_ptr(), stt["exp_avg"].data_ptr(), stt["exp_avg_sq"].data_ptr(),
                                          p.numel(), float(group["lr"]), float(b1), float(b2), float(group["eps"]),
                                          float(group["weight_decay"]), int(stt["step"]), 1.0, _lib.stream_ptr()))
                torch.autograd.graph.increment_version(p)   # the kernel wrote p in place: invalidate packed copies
        return loss
