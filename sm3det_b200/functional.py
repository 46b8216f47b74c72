"""autograd.Function wrappers: one per fused stage of the backbone, forward and hand-written backward.

Each Function only sequences C-ABI kernel calls (sm3det_b200.ops); tensors stay NHWC fp32 between
stages.  What each one replaces in the reference (mmrotate/models/backbones/convnext_moe.py):
  StemFn        dataset_stems['single'] + downsample_layers[0]            :783-791, :800-806
  DownsampleFn  Sequential(LayerNorm2d, Conv2d(2, stride 2))              :549-558, :806
  DenseBlockFn  ConvNeXtBlock._inner_forward with FFN                     :343-372, :397-405
  MoEBlockFn    ConvNeXtBlock._inner_forward with MoE_layer               :343-372, :226-293
  OutNormFn     norm{i}(x) channel_first incl. permute+contiguous         :811-817, :34-47
Backward follows SURVEY.md Appendix F (what autograd derives for the reference).

The only torch arithmetic left in this file is O(#parameters) glue on weight-sized tensors
(transposing 7x7 taps, flipping them for dgrad, multiplying a [C] vector by gamma).
"""
import torch
from torch.autograd import Function

from . import ops
from .ops import (EPI_AUXSTORE, EPI_COLSCALE, EPI_DGELU, EPI_GELU, EPI_RESID, EPI_ROWSCALE, LN_NCHW, LN_NHWC, LN_PATCH2)


def _zeros_like_param(p):
    return torch.zeros(p.shape, device=p.device, dtype=torch.float32)


def _taps(dww):            # [C,1,7,7] -> [49][C]
    return dww.reshape(dww.shape[0], 49).t().contiguous()


def _taps_flipped(dww):    # correlation taps for dgrad
    return dww.flip(2, 3).reshape(dww.shape[0], 49).t().contiguous()


@ops.captures_precision
class StemFn(Function):
    @staticmethod
    def forward(ctx, x, w, b, lnw, lnb, eps, ps):
        C0 = w.shape[0]
        wt = w.reshape(C0, -1).t().contiguous()
        train = any(ctx.needs_input_grad)
        x = x.contiguous().float()
        y, conv, stats = ops.stem_fwd(x, wt, b, lnw, lnb, eps, ps, save=train)
        if train:
            ctx.save_for_backward(x, conv, stats, lnw)
            ctx.ps = ps
            ctx.wshape = tuple(w.shape)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, conv, stats, lnw = ctx.saved_tensors
        dy = dy.contiguous()
        C0 = lnw.shape[0]
        T = conv.numel() // C0
        dlnw, dlnb = _zeros_like_param(lnw), _zeros_like_param(lnw)
        du = ops.layernorm_bwd(dy, conv, stats, lnw, dlnw, dlnb, tokens=T, C=C0)
        # weight gradient on the tensor cores: patches gathered once (im2col of the non-overlapping ps x ps patches, columns
        # ordered (kh, kw, ci), zero padded to a multiple of 32) and reduced over all tokens by the split-K GEMM -- 5x faster
        # than the SIMT stem_wgrad kernel at 1024^2 (0.25 vs 1.15 ms per step)
        N, Cin, H, W = x.shape
        ps = ctx.ps
        K = Cin * ps * ps
        Kp = (K + 31) // 32 * 32
        if C0 % 8 == 0:
            col, _, _ = ops.im2col(x, N=N, H=H, W=W, Cin=Cin, ks=ps, stride=ps, pad=0, Kp=Kp, nchw=True)
            dw2 = torch.zeros((C0, Kp), device=x.device, dtype=torch.float32)
            ops.linear_wgrad(du, col, dw2)
            db = torch.zeros((C0,), device=x.device, dtype=torch.float32)
            ops.colsum(du, db, rows=T, Cc=C0)
            dw = dw2[:, :K].reshape(C0, ps, ps, Cin).permute(0, 3, 1, 2).contiguous()
            return None, dw, db, dlnw, dlnb, None, None
        dwt = torch.zeros((K, C0), device=x.device, dtype=torch.float32)
        db = torch.zeros((C0,), device=x.device, dtype=torch.float32)
        ops.stem_wgrad(x, du, dwt, db, ctx.ps)
        dw = dwt.t().reshape(ctx.wshape).contiguous()
        return None, dw, db, dlnw, dlnb, None, None


@ops.captures_precision
class DownsampleFn(Function):
    @staticmethod
    def forward(ctx, x, lnw, lnb, w, b, eps):
        N, H, W, C = x.shape
        Co = w.shape[0]
        T = N * H * W
        train = any(ctx.needs_input_grad)
        xn = torch.empty((T // 4, 4 * C), device=x.device, dtype=torch.float32)
        _, stats = ops.layernorm_fwd(x, lnw, lnb, eps, tokens=T, C=C, out=xn, out_mode=LN_PATCH2, H=H, W=W,
                                     save_stats=train)
        w2 = w.permute(0, 2, 3, 1).reshape(Co, 4 * C).contiguous()      # [Co, (kh, kw, ci)]
        # w2 is a per-call re-ordered copy of the conv weight: split it once here (a few us) so both operands take the
        # bulk-copy main loop (4x the throughput of the in-kernel split on these shapes)
        y = ops.linear_fwd(xn, w2, b, packed=ops.pack_weight(w2, transposed=False))
        if train:
            ctx.save_for_backward(x, stats, xn, lnw, w2)
            ctx.dims = (N, H, W, C, Co)
        return y.view(N, H // 2, W // 2, Co)

    @staticmethod
    def backward(ctx, dy):
        x, stats, xn, lnw, w2 = ctx.saved_tensors
        N, H, W, C, Co = ctx.dims
        T = N * H * W
        dy2 = dy.contiguous().view(T // 4, Co)
        dxn = ops.linear_dgrad(dy2, w2, packed=ops.pack_weight(w2, transposed=True))
        dw2 = torch.zeros_like(w2)
        ops.linear_wgrad(dy2, xn, dw2)
        db = torch.zeros((Co,), device=x.device, dtype=torch.float32)
        ops.colsum(dy2, db, rows=T // 4, Cc=Co)
        dlnw, dlnb = _zeros_like_param(lnw), _zeros_like_param(lnw)
        dx = ops.layernorm_bwd(dxn, x, stats, lnw, dlnw, dlnb, tokens=T, C=C, in_mode=LN_PATCH2, H=H, W=W)
        dw = dw2.view(Co, 2, 2, C).permute(0, 3, 1, 2).contiguous()
        return dx.view(N, H, W, C), dlnw, dlnb, dw, db, None


@ops.captures_precision
class OutNormFn(Function):
    @staticmethod
    def forward(ctx, x, w, b, eps):
        N, H, W, C = x.shape
        T = N * H * W
        train = any(ctx.needs_input_grad)
        y = torch.empty((N, C, H, W), device=x.device, dtype=torch.float32)
        _, stats = ops.layernorm_fwd(x, w, b, eps, tokens=T, C=C, out=y, out_mode=LN_NCHW, H=H, W=W, save_stats=train)
        if train:
            ctx.save_for_backward(x, stats, w)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, stats, w = ctx.saved_tensors
        N, H, W, C = x.shape
        dw, db = _zeros_like_param(w), _zeros_like_param(w)
        dx = ops.layernorm_bwd(dy.contiguous(), x, stats, w, dw, db, tokens=N * H * W, C=C, in_mode=LN_NCHW, H=H, W=W)
        return dx.view(N, H, W, C), dw, db, None


def _block_front(x, dww, dwb, lnw, lnb, eps, train):
    N, H, W, C = x.shape
    u = ops.dwconv7(x, _taps(dww), dwb)
    v, stats = ops.layernorm_fwd(u, lnw, lnb, eps, tokens=N * H * W, C=C, save_stats=train)
    return u, v.view(N * H * W, C), stats


def _block_front_bwd(dv, dout, x, u, stats, dww, lnw):
    """LN backward -> depthwise dgrad (+ shortcut gradient) and the depthwise/LN parameter grads."""
    N, H, W, C = x.shape
    T = N * H * W
    dlnw, dlnb = _zeros_like_param(lnw), _zeros_like_param(lnw)
    du = ops.layernorm_bwd(dv, u, stats, lnw, dlnw, dlnb, tokens=T, C=C).view(N, H, W, C)
    dx = ops.dwconv7(du, _taps_flipped(dww), None, resid=dout)
    ddwt = torch.zeros((49, C), device=x.device, dtype=torch.float32)
    ddwb = torch.zeros((C,), device=x.device, dtype=torch.float32)
    ops.dwconv7_wgrad(x, du, ddwt, ddwb)
    ddww = ddwt.t().reshape(C, 1, 7, 7).contiguous()
    return dx, ddww, ddwb, dlnw, dlnb


@ops.captures_precision
class DenseBlockFn(Function):
    """Dense ConvNeXt block.  Narrow stages (ops.ffn_chunk: C <= 192) run the FFN forward as the fused tcgen05 kernel of
    csrc/ffn_fused.cu (GEMM1 -> GELU -> GEMM2 on chip; the hidden tensor is written once, as fp32 h, only when a backward
    follows).  The backward is the GEMM sequence (dgrad2 -> act_pack -> wgrads / dgrad1): fused recompute kernels for it
    were built and measured slower (narrow tcgen05 MMAs are paced by the 64-byte/clk operand fetch, not by N --
    profiles/r02_mma_microbench.txt).  Wider stages keep GEMM -> act_pack -> GEMM."""

    @staticmethod
    def forward(ctx, x, dww, dwb, lnw, lnb, w1, b1, w2, b2, gamma, row_scale, eps, packs):
        N, H, W, C = x.shape
        T = N * H * W
        # autograd.Function.forward runs with grad mode off and needs_input_grad reflects requires_grad only: the caller
        # says whether a backward can follow (it also chose which weight images to provide on that basis)
        train = any(ctx.needs_input_grad) and packs.get('grad', True)
        fused = packs.get('fused')
        # packs['shortcut'] = False: return the branch gamma * ffn(...) alone (ConvNeXt_DA gates it before the shortcut add)
        ctx.shortcut = packs.get('shortcut', True)
        resid = x.view(T, C) if ctx.shortcut else None
        if fused is not None:
            # LayerNorm writes the FFN's A-operand image directly (the separate split pass never exists); the fused kernel
            # keeps the hidden tensor on chip and, when a backward follows, stores the pre-activation h once for it
            u = ops.dwconv7(x, _taps(dww), dwb)
            v_img, v, stats = ops.layernorm_fwd_img(u, lnw, lnb, eps, tokens=T, C=C, save_stats=train, want_f32=train)
            res = ops.ffn_fused_fwd(v_img, packs['w1_c'][0], packs['w2_n'][0], b1, b2, T=T, C=C, chunk=fused['fwd'],
                                    gamma=gamma, row_scale=row_scale, resid=resid, want_aux=train, want_h=train)
            if train:
                ctx.save_for_backward(x, u, stats, v, res[2], res[1], dww, lnw, w1, w2, gamma, row_scale)
                ctx.packs = packs
            return res[0].view(N, H, W, C)
        u, v, stats = _block_front(x, dww, dwb, lnw, lnb, eps, train)
        # GEMM1 stores the pre-activation only; GELU runs in the HBM-bound act_pack kernel, which emits the result
        # directly as GEMM2's pre-split A operand (fp32 `a` never exists)
        h = ops.linear_fwd(v, w1, b1, packed=packs.get('w1'))
        a_k, _, _ = ops.act_pack(h, rows=T, width=4 * C, mode=ops.ACT_GELU, want_k=True)
        y2 = torch.empty((T, C), device=x.device, dtype=torch.float32) if train else None
        epi = (EPI_COLSCALE | (EPI_RESID if resid is not None else 0) | (EPI_ROWSCALE if row_scale is not None else 0)
               | (EPI_AUXSTORE if train else 0))
        out = ops.linear_fwd(None, w2, b2, rows=T, a_packed=a_k, epilogue=epi, aux_out=y2, col_scale=gamma,
                             row_scale=row_scale, resid=resid, packed=packs.get('w2'))
        if train:
            ctx.save_for_backward(x, u, stats, v, h, y2, dww, lnw, w1, w2, gamma, row_scale)
            ctx.packs = packs
        return out.view(N, H, W, C)

    @staticmethod
    def backward(ctx, dout):
        x, u, stats, v, h, y2, dww, lnw, w1, w2, gamma, rs = ctx.saved_tensors
        N, H, W, C = x.shape
        T = N * H * W
        dout = dout.contiguous()
        dz = dout.view(T, C)
        dev = x.device
        if rs is None:
            csum, dgamma = ops.colstat(dz, rows=T, Cc=C, y=y2)            # one pass: sum dz and sum dz * y2
        else:
            dgamma = torch.zeros((C,), device=dev, dtype=torch.float32)
            ops.colsum(dz, dgamma, rows=T, Cc=C, b=y2, row_scale=rs)
            csum = torch.zeros((C,), device=dev, dtype=torch.float32)
            ops.colsum(dz, csum, rows=T, Cc=C, row_scale=rs)
        db2 = csum * gamma
        w2g = ops.scale_rows(w2, row_scale=gamma)                   # gamma[c] * W2[c, :]
        da = ops.linear_dgrad(dz, w2g, epilogue=(EPI_ROWSCALE if rs is not None else 0), row_scale=rs,
                              packed=ops.pack_weight(w2g, transposed=True))
        # dh = da * gelu'(h) goes straight into the two operand images (dgrad1's A, wgrad1's A) + db1 column sums
        db1 = torch.zeros((4 * C,), device=dev, dtype=torch.float32)
        # one pass over h: dh = da * gelu'(h) as dgrad1's / wgrad1's operands (+ db1) and a = gelu(h) as wgrad2's operand
        dh_k, dh_mn, a_mn = ops.act_pack(h, rows=T, width=4 * C, mode=ops.ACT_BWD, da=da, want_k=True, mn_tile=128,
                                      mn_tile2=ops._pick_bn(4 * C), colsum=db1)
        del da
        dzs = dz if rs is None else ops.scale_rows(dz, row_scale=rs)
        dw2 = torch.zeros_like(w2)
        ops.linear_wgrad(dzs, None, dw2, rows=T, row_scale=gamma, x_packed=a_mn)
        del a_mn
        dw1 = torch.zeros_like(w1)
        ops.linear_wgrad(None, v, dw1, rows=T, dy_packed=dh_mn)
        dv = ops.linear_dgrad(None, w1, rows=T, a_packed=dh_k, packed=ctx.packs.get('w1_t'))
        dx, ddww, ddwb, dlnw, dlnb = _block_front_bwd(dv, dout if ctx.shortcut else None, x, u, stats, dww, lnw)
        return dx, ddww, ddwb, dlnw, dlnb, dw1, db1, dw2, db2, dgamma, None, None, None


def stack_expert_params(params):
    """Make E same-shaped parameters views of one contiguous [E, ...] buffer (grouped-GEMM layout).

    Parameter objects (and therefore optimizer state, named_parameters() and state_dict keys) are
    untouched; only ``.data`` is re-pointed.  No-op when they are already adjacent in memory.
    """
    base = params[0]
    step = base.numel() * base.element_size()
    if all(p.is_contiguous() and p.data_ptr() == base.data_ptr() + i * step for i, p in enumerate(params)):
        return
    with torch.no_grad():
        flat = torch.stack([p.data for p in params]).contiguous()
        for i, p in enumerate(params):
            p.data = flat[i]


@ops.captures_precision
class MoEBlockFn(Function):
    """x -> dwconv -> LN -> router/plan/assign -> grouped expert GEMMs -> combine (+gamma, +shortcut)."""

    @staticmethod
    def forward(ctx, x, dww, dwb, lnw, lnb, gamma, wp, bp, sim, tau, w_noise, row_scale, noise, eps, E, k, record, packs,
                *experts):
        N, H, W, C = x.shape
        T = N * H * W
        w1s, b1s, w2s, b2s = experts[0:E], experts[E:2 * E], experts[2 * E:3 * E], experts[3 * E:4 * E]
        train = any(ctx.needs_input_grad)
        u, v, stats = _block_front(x, dww, dwb, lnw, lnb, eps, train)
        r = ops.moe_router(v, wp, bp, sim, tau, T=T, Cc=C, E=E, k=k, w_noise=w_noise, noise=noise, save=train)
        plan = ops.moe_plan(r['partials'], T=T, E=E, k=k)
        slot_of, pair_token = ops.moe_assign(r['top_idx'], plan, T=T, E=E, k=k)
        R = plan['max_rows']
        grouped = (plan['tile_group'], plan['num_m_tiles'])
        h = ops.linear_fwd(v, w1s[0], b1s[0], row_index=pair_token, rows=R, grouped=grouped, w_group_stride=4 * C * C,
                           bias_group_stride=4 * C, packed=packs.get('w1'))
        a_k, _, _ = ops.act_pack(h, rows=R, width=4 * C, mode=ops.ACT_GELU, want_k=True, live_tiles=plan['num_m_tiles'])
        o = ops.linear_fwd(None, w2s[0], b2s[0], rows=R, a_packed=a_k, grouped=grouped, w_group_stride=4 * C * C,
                           bias_group_stride=C, packed=packs.get('w2'))
        ctx.shortcut = packs.get('shortcut', True)
        out, y = ops.moe_combine(o, slot_of, r['top_idx'], r['top_gate'], gamma, x.view(T, C) if ctx.shortcut else None,
                                 row_scale, T=T, Cc=C, k=k, want_y=record is not None)
        if record is not None:
            record.append(dict(v=v, top_idx=r['top_idx'], top_gate=r['top_gate'], importance=plan['importance'],
                               load=plan['load'], loss=plan['loss'], y=y, counts=plan['counts']))
        if train:
            ctx.noisy = noise is not None     # gates depend on w_noise whenever noise was added, also for k == E
            ctx.save_for_backward(x, u, stats, v, h, o, dww, lnw, gamma, wp, sim, tau, row_scale, r['top_idx'],
                                  r['top_gate'], r['logits'], r['p'], slot_of, pair_token, plan['importance'],
                                  plan['seg_begin'], plan['seg_end'], plan['tile_group'], plan['num_m_tiles'],
                                  w1s[0], w2s[0], noise, r['sigma'], r['top_vals'], r['top_idx_m'], plan['load'],
                                  w_noise)
            ctx.E, ctx.k, ctx.R = E, k, R
            ctx.packs = packs
            ctx.has_noise_param = w_noise is not None
        return out.view(N, H, W, C), plan['loss'].reshape(())

    @staticmethod
    def backward(ctx, dout, dloss):
        (x, u, stats, v, h, o, dww, lnw, gamma, wp, sim, tau, rs, top_idx, top_gate, logits, p, slot_of, pair_token,
         importance, seg_begin, seg_end, tile_group, num_m_tiles, w1, w2, noise, sigma, top_vals, top_idx_m, load,
         w_noise) = ctx.saved_tensors
        E, k, R = ctx.E, ctx.k, ctx.R
        N, H, W, C = x.shape
        T = N * H * W
        dev = x.device
        dout = dout.contiguous()
        dz = dout.view(T, C)
        grouped = (tile_group, num_m_tiles)
        segs = (seg_begin, seg_end)
        # combine / layer scale / shortcut
        d_o = torch.zeros((R, C), device=dev, dtype=torch.float32)
        dgamma = torch.zeros((C,), device=dev, dtype=torch.float32)
        dgate = ops.moe_combine_bwd(dz, o, slot_of, top_idx, top_gate, gamma, rs, d_o, dgamma, T=T, Cc=C, k=k)
        # experts (grouped over the padded expert segments)
        da = ops.linear_dgrad(d_o, w2, grouped=grouped, w_group_stride=4 * C * C, packed=ctx.packs.get('w2_t'))
        db1s = torch.zeros((E, 4 * C), device=dev, dtype=torch.float32)
        # one pass over h: dh = da * gelu'(h) as dgrad1's / wgrad1's operands (+ db1) and a = gelu(h) as wgrad2's operand
        dh_k, dh_mn, a_mn = ops.act_pack(h, rows=R, width=4 * C, mode=ops.ACT_BWD, da=da, want_k=True, mn_tile=128,
                                      mn_tile2=ops._pick_bn(4 * C), colsum=db1s, live_tiles=num_m_tiles, tile_group=tile_group)
        del da
        dw2s = torch.zeros((E, C, 4 * C), device=dev, dtype=torch.float32)
        ops.linear_wgrad(d_o, None, dw2s, rows=R, segs=segs, num_groups=E, x_packed=a_mn)
        del a_mn
        db2s = torch.zeros((E, C), device=dev, dtype=torch.float32)
        ops.colsum(d_o, db2s, rows=R, Cc=C, segs=segs, groups=E)
        dw1s = torch.zeros((E, 4 * C, C), device=dev, dtype=torch.float32)
        ops.linear_wgrad(None, v, dw1s, rows=R, x_row_index=pair_token, segs=segs, num_groups=E, dy_packed=dh_mn)
        dxp = torch.empty((R, C), device=dev, dtype=torch.float32)   # every row gather_sum reads (live slots) is written by the GEMM
        ops.linear_dgrad(None, w1, rows=R, a_packed=dh_k, out=dxp, grouped=grouped, w_group_stride=4 * C * C,
                         packed=ctx.packs.get('w1_t'))
        # router
        P = wp.shape[0]
        dtau = torch.zeros((1,), device=dev, dtype=torch.float32)
        dsim = torch.zeros((P, E), device=dev, dtype=torch.float32)
        lscale = dloss.reshape(1).contiguous().float()
        noisy = dict(noise=noise, sigma=sigma, top_vals=top_vals, top_idx_m=top_idx_m, load=load) if ctx.noisy else None
        dp, dr = ops.moe_router_bwd(p, sim, tau, top_idx, top_gate, dgate, logits, importance, lscale, dsim, dtau, T=T,
                                    E=E, k=k, noisy=noisy)
        dwp = torch.zeros_like(wp)
        ops.linear_wgrad(dp, v, dwp)
        dbp = torch.zeros((P,), device=dev, dtype=torch.float32)
        ops.colsum(dp, dbp, rows=T, Cc=P)
        dv_r = ops.linear_dgrad(dp, wp, packed=ctx.packs.get('wp_t'))
        dwn = None
        if ctx.noisy:
            # r = v @ w_noise is an [T,C]x[C,E] product with E < 32: run it as a 32-wide zero-padded GEMM pair
            wn_t = torch.zeros((32, C), device=dev, dtype=torch.float32)
            wn_t[:E] = w_noise.t()
            dwn_t = torch.zeros((32, C), device=dev, dtype=torch.float32)
            ops.linear_wgrad(dr, v, dwn_t)                             # [32,C] = dr^T v
            dwn = dwn_t[:E].t().contiguous()
            dv_r = ops.linear_dgrad(dr, wn_t, epilogue=EPI_RESID, resid=dv_r)
        dv = ops.gather_sum(dxp, slot_of, dv_r, T=T, Cc=C, k=k)
        dx, ddww, ddwb, dlnw, dlnb = _block_front_bwd(dv, dout if ctx.shortcut else None, x, u, stats, dww, lnw)
        if dwn is None and ctx.has_noise_param:
            dwn = torch.zeros((C, E), device=dev, dtype=torch.float32)
        grads_e = [dw1s[e] for e in range(E)] + [db1s[e] for e in range(E)] + [dw2s[e] for e in range(E)] + \
                  [db2s[e] for e in range(E)]
        return (dx, ddww, ddwb, dlnw, dlnb, dgamma, dwp, dbp, dsim, dtau, dwn, None, None, None, None, None, None, None,
                *grads_e)


@ops.captures_precision
class DAGateFn(Function):
    """ConvNeXt_DA block tail (convnext_moe_DA.py:400-401): out = shortcut + row_scale * s[n, c] * y  with y the gamma-scaled
    FFN branch [N,H,W,C] and s = DALayer's per-sample channel gate [N,C]; also returns nothing else -- the squeeze
    (per-sample mean of y) is SampleMeanFn.  Per sample one `affine` launch (N is the per-GPU batch)."""

    @staticmethod
    def forward(ctx, y, x, s, row_scale):
        N, H, W, C = y.shape
        y, x = y.contiguous(), x.contiguous()
        hw = H * W
        out = torch.empty_like(x)
        # per-sample scale vector: s[n] (* the sample's drop-path factor: row_scale is constant over a sample's tokens)
        sc = s if row_scale is None else s * row_scale.view(N, hw)[:, :1]
        sc = sc.contiguous()
        for n in range(N):
            ops.affine(y[n].view(hw, C), a1=sc[n], add=x[n].view(hw, C), out=out[n].view(hw, C))
        ctx.save_for_backward(y, sc, s, row_scale)
        return out

    @staticmethod
    def backward(ctx, d):
        y, sc, s, rs = ctx.saved_tensors
        N, H, W, C = y.shape
        hw = H * W
        d = d.contiguous()
        dy = torch.empty_like(y)
        dsc = torch.zeros((N, C), device=y.device, dtype=torch.float32)
        for n in range(N):
            ops.affine(d[n].view(hw, C), a1=sc[n], out=dy[n].view(hw, C))
            ops.colsum(d[n].view(hw, C), dsc[n], rows=hw, Cc=C, b=y[n].view(hw, C))       # sum_hw d * y
        ds = dsc if rs is None else dsc * rs.view(N, hw)[:, :1]
        return dy, d, ds, None


@ops.captures_precision
class SampleMeanFn(Function):
    """DALayer's squeeze: AdaptiveAvgPool2d(1) over an NHWC tensor -> [N, C]."""

    @staticmethod
    def forward(ctx, y):
        N, H, W, C = y.shape
        y = y.contiguous()
        m = torch.zeros((N, C), device=y.device, dtype=torch.float32)
        for n in range(N):
            ops.colsum(y[n].view(H * W, C), m[n], rows=H * W, Cc=C)
        ctx.shape = (N, H, W, C)
        return m / float(H * W)

    @staticmethod
    def backward(ctx, dm):
        N, H, W, C = ctx.shape
        return (dm / float(H * W)).view(N, 1, 1, C).expand(N, H, W, C).contiguous()
