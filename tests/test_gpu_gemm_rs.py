"""Row-stream GEMM engine (csrc/mt_gemm_rs.cu) against fp64 torch on the same bf16 operands: every (N, K, epilogue) combination the
encoder uses -- forward projections MFT/multiTransformer.py:19-20 (FFN), :47-65 (Q/K/V/out) and their input gradients -- grouped over
modality stacks (per-group weights, biases, dropout keys, LayerNorm gains), with ragged row counts for the single-group form, the same
pair-hash dropout masks as oracle/dropout_rng.py, and the fused LayerNorm (:81-91, unbiased std, eps on std)."""
import pytest
import torch

from multimodal_transformer_b200 import _lib

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
RELU = 1


def _run(G, Mg, N, K, *, bkm=1, c_f32=0, bias=False, act=0, p=0.0, gate=False, res=False, colsum=False, ln=False, seed=77, site=40, scale=1.0, res_off=0.5, res_scale=2.0):
    from oracle.dropout_rng import keep_mask
    L = _lib.lib()
    gen = torch.Generator().manual_seed(G * 1000 + Mg + N * 3 + K * 7 + int(p * 100))
    rows = G * Mg
    A = (torch.randn(rows, K, generator=gen) * scale).bfloat16()
    W = (torch.randn(G, N, K, generator=gen) / K ** 0.5).bfloat16()             # logical B_g(n, k)
    b = torch.randn(G, N, generator=gen) if bias else None
    gt = torch.randn(rows, N, generator=gen).bfloat16() if gate else None
    r = (torch.randn(rows, N, generator=gen) * res_scale + res_off) if res else None
    la = (torch.randn(G, N, generator=gen) * 0.5 + 1.0) if ln else None
    lb = torch.randn(G, N, generator=gen) if ln else None
    cs0 = torch.randn(G, N, generator=gen) if colsum else None
    want = torch.empty(rows, N, dtype=torch.float64)
    for g in range(G):
        sl = slice(g * Mg, (g + 1) * Mg)
        y = A[sl].double() @ W[g].double().t()
        if bias:
            y = y + b[g].double()
        if act == RELU:
            y = torch.relu(y)
        if p > 0:
            y = y * keep_mask(seed, site + 512 * g, (Mg, N), p).double() / (1.0 - float(torch.tensor(p, dtype=torch.float32)))
        if gate:
            y = torch.where(gt[sl].double() > 0, y * 1.25, torch.zeros_like(y))
        if res:
            y = y + r[sl].double()
        want[sl] = y
    Wd = (W if bkm else W.transpose(1, 2).contiguous()).to(DEV)
    Ad = A.to(DEV)
    C = torch.full((rows, N), float('nan'), device=DEV, dtype=torch.float32 if c_f32 else torch.bfloat16)
    lnout = torch.full((rows, N), float('nan'), device=DEV, dtype=torch.bfloat16) if ln else None
    d = lambda t: None if t is None else t.to(DEV)
    bd, gd, rd, lad, lbd, csd = d(b), d(gt), d(r), d(la), d(lb), d(cs0)
    rc = L.mt_gemm_rs(G, Mg, N, K, _lib.ptr(Ad), _lib.ptr(Wd), bkm, _lib.ptr(C), c_f32, _lib.ptr(bd), act, p, seed, site, _lib.ptr(gd), 1.25,
                      _lib.ptr(rd), _lib.ptr(csd), _lib.ptr(lnout), _lib.ptr(lad), _lib.ptr(lbd), None)
    _lib.check(rc)
    torch.cuda.synchronize()
    got = C.double().cpu()
    tol = (2e-5 if c_f32 else 6e-3) * max(1.0, want.abs().max().item())
    err = (got - want).abs().max().item()
    assert err <= tol, f'C: max err {err} (tol {tol})'
    if colsum:
        wcs = cs0.double() + torch.stack([want[g * Mg:(g + 1) * Mg].sum(0) for g in range(G)])      # summed before the bf16 rounding
        e = (csd.double().cpu() - wcs).abs().max().item()
        assert e <= 2e-5 * max(1.0, wcs.abs().max().item()) * Mg ** 0.5, f'colsum err {e}'
    if ln:
        x = got                                        # the fp32 output the kernel normalises
        mean = x.mean(-1, keepdim=True); std = x.std(-1, keepdim=True)
        wl = torch.cat([la[g].double() * (x[g * Mg:(g + 1) * Mg] - mean[g * Mg:(g + 1) * Mg]) / (std[g * Mg:(g + 1) * Mg] + 1e-6) + lb[g].double()
                        for g in range(G)])
        e = (lnout.double().cpu() - wl).abs().max().item()
        assert e <= 6e-3 * max(1.0, wl.abs().max().item()), f'layernorm err {e}'


# (N, K, kwargs): the encoder's forward projections
FWD = [
    (768, 256, dict(bias=True)),
    (256, 256, dict(bias=True, c_f32=1, res=True)),
    (256, 256, dict(bias=True, c_f32=1, res=True, p=0.1)),
    (128, 256, dict(bias=True, act=RELU)),
    (128, 256, dict(bias=True, act=RELU, p=0.1)),
    (256, 128, dict(bias=True, c_f32=1, res=True)),
    (256, 128, dict(bias=True, c_f32=1, res=True, p=0.1)),
]
BWD = [
    (128, 256, dict(bkm=0, gate=True, colsum=True)),
    (256, 128, dict(bkm=0)),
    (256, 256, dict(bkm=0)),
]
LN = [
    (256, 128, dict(bias=True, c_f32=1, res=True, ln=True)),
    (256, 128, dict(bias=True, c_f32=1, res=True, ln=True, p=0.1)),
    # output projection + LayerNorm: two 128-column slices per row tile, the row moments cross the CTA pair through distributed shared memory
    (256, 256, dict(bias=True, c_f32=1, res=True, ln=True)),
    (256, 256, dict(bias=True, c_f32=1, res=True, ln=True, p=0.1)),
]


@pytest.mark.parametrize('N,K,kw', FWD + BWD + LN)
def test_rs_single_group_small(N, K, kw):
    _run(1, 128, N, K, **kw)


@pytest.mark.parametrize('N,K,kw', FWD + BWD + LN)
def test_rs_single_group_ragged_rows(N, K, kw):
    _run(1, 1000, N, K, **kw)


@pytest.mark.parametrize('N,K,kw', FWD + BWD + LN)
def test_rs_three_groups_many_tiles_per_cta(N, K, kw):
    """G = 3 with more row tiles than CTAs per (group, slice) pair: ring wrap-around, accumulator hand-off, staging reuse."""
    _run(3, 128 * 70, N, K, **kw)


def test_rs_two_groups_one_tile():
    _run(2, 128, 768, 256, bias=True)
    _run(2, 256, 256, 256, bias=True, c_f32=1, res=True, p=0.3)


def test_rs_layernorm_offset_rows():
    """Rows with a mean far from zero relative to their spread: the shifted single-pass moments must not cancel."""
    _run(1, 256, 256, 128, bias=True, c_f32=1, res=True, ln=True, scale=0.01, res_off=30.0, res_scale=0.5)
    _run(1, 256, 256, 256, bias=True, c_f32=1, res=True, ln=True, scale=0.01, res_off=30.0, res_scale=0.5)


def test_rs_rejects_unsupported():
    L = _lib.lib()
    A = torch.zeros(128, 192, device=DEV, dtype=torch.bfloat16); W = torch.zeros(256, 192, device=DEV, dtype=torch.bfloat16)
    C = torch.zeros(128, 256, device=DEV, dtype=torch.bfloat16)
    rc = L.mt_gemm_rs(1, 128, 256, 192, _lib.ptr(A), _lib.ptr(W), 1, _lib.ptr(C), 0, None, 0, 0.0, 0, 0, None, 1.0, None, None, None, None, None, None)
    assert rc == 5


def test_grouped_stacks_equal_per_stack_calls():
    """MultiTransformer through the grouped encoder call (one launch per projection / LayerNorm / weight gradient for all modality
    stacks) against the same model run one stack at a time: same dropout seeds -> same masks; fp32 mode agrees to round-off, bf16 mode
    to the GEMM engines' summation order."""
    import numpy as np
    import multimodal_transformer_b200 as mtb
    from oracle import fill
    from tests import util
    MODS = ['acoustic', 'image', 'linguistic']
    dims = {'acoustic': 88, 'image': 256, 'linguistic': 300}
    N, B, T = 2, 6, 128
    sd = util.filled_sd(util.mods_shapes('MFT.MultiTransformer', N), 91)
    inputs, mask, target, lengths = fill.make_batch(B, T, dims, 91)
    t = lambda a: torch.from_numpy(np.asarray(a))
    try:
        for mode, tol, gtol in (('fp32', 2e-6, 2e-5), ('bf16', 2e-2, 6e-2)):
            mtb.set_compute_dtype(mode)
            res = []
            for grouped in (True, False):
                old = mtb.set_grouped_stacks(grouped)
                try:
                    model = mtb.MultiTransformer(MODS, dims, N=N).train(); model.load_state_dict(sd)
                    mtb.fix_seed(1234)
                    pred = model({k: t(v).to(DEV) for k, v in inputs.items()}, t(mask).to(DEV), lengths)
                    (((pred - t(target).to(DEV)) ** 2).sum() / sum(lengths)).backward()
                    res.append((pred.detach().float().cpu(), {k: p.grad.detach().cpu() for k, p in model.named_parameters() if p.grad is not None}))
                finally:
                    mtb.set_grouped_stacks(old)
            if mode == 'bf16':      # ... and against the streaming GEMM engine with the stand-alone attention-backward preparation pass
                old = _lib.lib().mt_tune(5, 1)
                try:
                    model = mtb.MultiTransformer(MODS, dims, N=N).train(); model.load_state_dict(sd)
                    mtb.fix_seed(1234)
                    pred = model({k: t(v).to(DEV) for k, v in inputs.items()}, t(mask).to(DEV), lengths)
                    (((pred - t(target).to(DEV)) ** 2).sum() / sum(lengths)).backward()
                    res.append((pred.detach().float().cpu(), {k: p.grad.detach().cpu() for k, p in model.named_parameters() if p.grad is not None}))
                finally:
                    _lib.lib().mt_tune(5, old)
                (ps, gs_) = res[2]
                gmax_s = max(v.abs().max().item() for v in gs_.values())
                assert (res[0][0] - ps).abs().max().item() <= tol * max(1.0, ps.abs().max().item())
                for k in gs_:
                    e = (res[0][1][k] - gs_[k]).abs().max().item()
                    assert e <= gtol * max(gs_[k].abs().max().item(), 1e-3 * gmax_s), ('streaming engine', k, e)
            (pg, gg), (pu, gu) = res[0], res[1]
            assert (pg - pu).abs().max().item() <= tol * max(1.0, pu.abs().max().item()), mode
            assert gg.keys() == gu.keys()
            gmax = max(v.abs().max().item() for v in gu.values())
            for k in gu:
                e = (gg[k] - gu[k]).abs().max().item()
                assert e <= gtol * max(gu[k].abs().max().item(), 1e-3 * gmax), (mode, k, e)
    finally:
        mtb.set_compute_dtype('fp32'); mtb.fix_seed(None)
