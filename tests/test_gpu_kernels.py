"""GPU parity tests of the individual kernels THROUGH THE C ABI (ctypes) against the CPU oracle.

Tolerances: 1e-5 relative (+1e-6 absolute near zero) for float32 log-probs / log-dets / gradients (north_star);
bit-exact for neighbour-selection outputs / indices and MC accept / reject decisions.
"""
import ctypes as C
import os

import numpy as np
import pytest

from helpers import assert_close
from oracle import dists as odists
from oracle import mappings as omap
from oracle import mcmc as omc
from oracle import rqs as orqs
from oracle import vae as ovae

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), 'golden')


def T(v, a, dtype=np.float32):
    return v.Tensor.from_numpy(np.ascontiguousarray(a, dtype=dtype))


def _raw(rng, n, K, scale=1.5):
    return (rng.normal(0, scale, (n, K)).astype(np.float32), rng.normal(0, scale, (n, K)).astype(np.float32),
            rng.normal(0, scale, (n, K - 1)).astype(np.float32))


# ------------------------------------------------------------------------------------------------ K1: RQS
def _cond_tol(fldj, val, rel=1e-5):
    """Float32 knot positions carry ~ulp(range) absolute error, which a bin of slope dy/dx multiplies: the bound on a
    value error is rel * max(1, |value|) * (1 + local derivative).  The reference's own float32 TF ops have the same
    conditioning."""
    return rel * np.maximum(1.0, np.abs(val)) * (1.0 + np.exp(fldj))


@pytest.mark.parametrize('K', [32, 20, 8, 47, 2])
@pytest.mark.parametrize('n', [1, 127, 129, 1000])
def test_rqs_forward_inverse_match_oracle(vms, K, n):
    v = vms
    c = v._abi.ctx()
    for scale in (0.5, 1.5):  # conditioner-like logits, and wild ones (bin slopes up to ~1e3)
        rng = np.random.default_rng(K * 1000 + n)
        rw, rh, rs = _raw(rng, n, K, scale)
        x = rng.uniform(-11, 11, n).astype(np.float32)
        x[::17] = 10.5  # clearly outside: identity, zero log-det
        d = [T(v, a) for a in (x, rw, rh, rs)]
        f64 = lambda a: a.astype(np.float64)
        for name, ofn, inv in (('vms_rqs_forward', orqs.rqs_forward_raw, False), ('vms_rqs_inverse', orqs.rqs_inverse_raw, True)):
            y, l = v.Tensor((n, )), v.Tensor((n, ))
            getattr(c.lib, name)(d[0].ptr, d[1].ptr, d[2].ptr, d[3].ptr, n, K, -10.0, 10.0, y.ptr, l.ptr, c.stream)
            # reference in float64 from the same float32 logits (the float32 oracle itself carries rounding noise)
            yo, lo = ofn(f64(x), f64(rw), f64(rh), f64(rs), -10.0, 10.0)
            y32, l32 = ofn(x, rw, rh, rs, -10.0, 10.0)
            yg, lg = y.numpy(), l.numpy()
            oob = (x <= -10) | (x >= 10)
            assert np.array_equal(yg[oob], x[oob]) and np.all(lg[oob] == 0)
            slope_ldj = -lo if inv else lo  # log dy/dx of the map being evaluated ... of its output w.r.t. knots
            tol = _cond_tol(np.abs(lo), yo)
            err = np.abs(yg - yo)
            assert np.all(err <= tol + 2e-6), '%s value K=%d scale=%g: worst %g (tol %g)' % (
                name, K, scale, err.max(), tol[np.argmax(err - tol)])
            # the kernel is as close to float64 truth as the float32 oracle is (same algorithm, same conditioning)
            ref_err = np.abs(y32 - yo)
            assert np.median(err) <= 2 * np.median(ref_err) + 1e-6
            lerr, lref = np.abs(lg - lo), np.abs(l32 - lo)
            assert np.median(lerr) <= 2 * np.median(lref) + 2e-6, (np.median(lerr), np.median(lref))
            assert np.quantile(lerr, 0.99) <= 4 * np.quantile(lref, 0.99) + 2e-5, (np.quantile(lerr, 0.99), np.quantile(lref, 0.99))
            if scale == 0.5:  # north_star tolerance on conditioner-like parameters
                assert_close(yg, yo, rtol=2e-5, atol=2e-5, what='%s value K=%d' % (name, K))
                assert_close(lg, lo, rtol=2e-5, atol=3e-5 if K <= 32 else 6e-5, what='%s ldj K=%d' % (name, K))


def test_rqs_boundary_is_identity_or_last_bin(vms):
    """x exactly on the range edge: TFP's last knot is a float32 cumsum, so either answer (identity or last bin) can
    occur; both give y ~ x and a finite log-det."""
    v = vms
    c = v._abi.ctx()
    n, K = 64, 32
    rng = np.random.default_rng(0)
    rw, rh, rs = _raw(rng, n, K, 0.5)
    x = np.where(np.arange(n) % 2 == 0, 10.0, -10.0).astype(np.float32)
    d = [T(v, a) for a in (x, rw, rh, rs)]
    y, l = v.Tensor((n, )), v.Tensor((n, ))
    c.lib.vms_rqs_forward(d[0].ptr, d[1].ptr, d[2].ptr, d[3].ptr, n, K, -10.0, 10.0, y.ptr, l.ptr, c.stream)
    assert np.allclose(y.numpy(), x, atol=1e-4) and np.all(np.isfinite(l.numpy()))
    assert np.all(l.numpy()[x == -10.0] == 0)  # x <= range_min is always outside


def test_rqs_round_trip_at_scale(vms):
    """Size-independent property at 2M elements: inverse(forward(x)) == x, ildj == -fldj, monotone."""
    v = vms
    c = v._abi.ctx()
    n, K = 1 << 21, 32
    rng = np.random.default_rng(0)
    rw, rh, rs = _raw(rng, n, K, 0.5)
    x = rng.uniform(-10.5, 10.5, n).astype(np.float32)
    dx, dw, dh, ds = (T(v, a) for a in (x, rw, rh, rs))
    y, fl, xb, il = v.Tensor((n, )), v.Tensor((n, )), v.Tensor((n, )), v.Tensor((n, ))
    c.lib.vms_rqs_forward(dx.ptr, dw.ptr, dh.ptr, ds.ptr, n, K, -10.0, 10.0, y.ptr, fl.ptr, c.stream)
    c.lib.vms_rqs_inverse(y.ptr, dw.ptr, dh.ptr, ds.ptr, n, K, -10.0, 10.0, xb.ptr, il.ptr, c.stream)
    xb, fl, il = xb.numpy(), fl.numpy(), il.numpy()
    # a float32 round trip through a bin of slope s loses ~ulp(10)/s; bound the error by the local derivative
    err = np.abs(xb - x)
    bound = 1e-5 * (1 + np.exp(-fl)) * (1 + np.exp(fl)) + 2e-5  # knot error x slope going out, / slope coming back
    worst = np.argsort(err - bound)[-3:]
    # TFP's float32 quadratic solve (2c / (-b - sqrt(b^2 - 4ac))) loses digits near double roots: allow a 1e-4 tail
    assert np.mean(err <= bound) > 0.9999, [(float(x[i]), float(xb[i]), float(fl[i])) for i in worst]
    assert err.max() < 2e-3 and np.median(err) < 3e-6
    assert np.all(np.abs(il + fl) <= 2e-4 + 1e-4 * np.abs(fl)), np.abs(il + fl).max()


@pytest.mark.parametrize('Dt,K', [(1, 32), (2, 20), (3, 8), (5, 32)])
def test_rqs_strided_event_sum(vms, Dt, K):
    """Coupling-layer form: column slices of a [B, D] tensor, fused [w|h|s] parameter rows, event-summed ldj."""
    v = vms
    c = v._abi.ctx()
    B, D = 333, Dt + 2
    rng = np.random.default_rng(Dt * 100 + K)
    ldr = Dt * (3 * K - 1)
    raw = rng.normal(0, 0.5, (B, ldr)).astype(np.float32)
    x = rng.uniform(-11, 11, (B, D)).astype(np.float32)
    prev = rng.normal(size=B).astype(np.float32)
    nw = Dt * K
    rw, rh, rs = raw[:, :nw].reshape(B, Dt, K), raw[:, nw:2 * nw].reshape(B, Dt, K), raw[:, 2 * nw:].reshape(B, Dt, K - 1)
    dX, dR = T(v, x), T(v, raw)
    for inv, ofn in ((0, orqs.rqs_forward_raw), (1, orqs.rqs_inverse_raw)):
        out = T(v, x)
        lsum = T(v, prev)
        lel = v.Tensor((B, Dt))
        a = v._abi.RqsArgs(B, Dt, K, -10.0, 10.0, dX.ptr + 4, D, dR.ptr, ldr, dR.ptr + 4 * nw, ldr, dR.ptr + 8 * nw, ldr,
                           out.ptr + 4, D, lel.ptr, lsum.ptr, 1, inv)
        c.lib.vms_rqs_apply(C.byref(a), c.stream)
        yo, lo = ofn(x[:, 1:1 + Dt].astype(np.float64), rw.astype(np.float64), rh.astype(np.float64), rs.astype(np.float64),
                     -10.0, 10.0)
        got = out.numpy()
        assert np.array_equal(got[:, 0], x[:, 0]) and np.array_equal(got[:, 1 + Dt:], x[:, 1 + Dt:])
        assert_close(got[:, 1:1 + Dt], yo, rtol=2e-5, atol=2e-5, what='strided value')
        assert_close(lel.numpy(), lo, rtol=2e-5, atol=3e-5, what='strided per-element ldj')
        assert_close(lsum.numpy(), prev + lo.sum(-1), rtol=2e-5, atol=5e-5, what='accumulated event sum')


@pytest.mark.parametrize('K', [32, 20, 6])
@pytest.mark.parametrize('inverse_dir', [0, 1])
def test_rqs_backward_matches_oracle(vms, K, inverse_dir):
    v = vms
    c = v._abi.ctx()
    n = 777
    rng = np.random.default_rng(K + inverse_dir)
    rw, rh, rs = _raw(rng, n, K, 0.5)
    x = rng.uniform(-10.5, 10.5, n).astype(np.float32)
    g_out, g_ldj = rng.normal(size=n).astype(np.float32), rng.normal(size=n).astype(np.float32)
    d = [T(v, a) for a in (x, rw, rh, rs, g_out, g_ldj)]
    g_in, g_w, g_h, g_s = v.Tensor((n, )), v.Tensor((n, K)), v.Tensor((n, K)), v.Tensor((n, K - 1))
    c.lib.vms_rqs_backward(d[0].ptr, d[1].ptr, d[2].ptr, d[3].ptr, n, K, -10.0, 10.0, inverse_dir, d[4].ptr, d[5].ptr,
                           g_in.ptr, g_w.ptr, g_h.ptr, g_s.ptr, c.stream)
    f64 = lambda a: a.astype(np.float64)
    o = orqs.rqs_backward_raw(f64(x), f64(rw), f64(rh), f64(rs), -10.0, 10.0, f64(g_out), f64(g_ldj),
                              inverse_dir=bool(inverse_dir))
    for got, want, nm in zip((g_in, g_w, g_h, g_s), o, ('g_in', 'g_raw_w', 'g_raw_h', 'g_raw_s')):
        scale = np.abs(want).max()
        assert_close(got.numpy(), want, rtol=5e-5, atol=1e-5 * max(1.0, scale), what='%s K=%d inv=%d' % (nm, K, inverse_dir))
        rel = np.linalg.norm(got.numpy() - want) / np.linalg.norm(want)
        assert rel < 1e-5, (nm, rel)


# ------------------------------------------------------------------------------------------------ K2/K3: dense
@pytest.mark.parametrize('B,K,N,act', [(4096, 6, 200, 1), (4096, 200, 4, 0), (1000, 100, 95, 0), (130, 1, 100, 2),
                                       (65, 17, 130, 2), (1, 3, 5, 1)])
def test_dense_forward_backward(vms, B, K, N, act):
    v = vms
    c = v._abi.ctx()
    rng = np.random.default_rng(B + K + N)
    x = rng.normal(size=(B, K)).astype(np.float32)
    W = (rng.normal(size=(K, N)) / np.sqrt(K)).astype(np.float32)
    b = rng.normal(size=N).astype(np.float32)
    g = rng.normal(size=(B, N)).astype(np.float32)
    dx, dW, db, dg = (T(v, a) for a in (x, W, b, g))
    out = v.Tensor((B, N))
    c.lib.vms_dense_forward(dx.ptr, K, dW.ptr, db.ptr, B, K, N, act, None, 0, None, 0, out.ptr, N, c.stream)
    pre = x.astype(np.float64) @ W.astype(np.float64) + b
    want = [pre, np.maximum(pre, 0), np.tanh(pre)][act]
    assert_close(out.numpy(), want, rtol=1e-5, atol=1e-5, what='dense forward')
    ws = v.Tensor((max(1, c.lib.vms_dense_backward_workspace(B, K, N, 0) // 4), ))
    gx, gW, gb = v.Tensor((B, K)), v.Tensor((K, N)), v.Tensor((N, ))
    c.lib.vms_dense_backward(dx.ptr, K, dW.ptr, B, K, N, act, out.ptr, N, dg.ptr, N, None, 0, None, 0, gx.ptr, K, 0,
                             gW.ptr, gb.ptr, None, 0, None, 0, ws.ptr, c.stream)
    o = out.numpy().astype(np.float64)
    gpre = g * [np.ones_like(o), (o > 0).astype(np.float64), 1 - o * o][act]
    sc = np.sqrt(B)
    assert_close(gx.numpy(), gpre @ W.T.astype(np.float64), rtol=1e-5, atol=2e-5, what='dense g_x')
    assert_close(gW.numpy(), x.T.astype(np.float64) @ gpre, rtol=1e-5, atol=2e-6 * sc, what='dense g_W')
    assert_close(gb.numpy(), gpre.sum(0), rtol=1e-5, atol=2e-6 * sc, what='dense g_b')


@pytest.mark.parametrize('B,K,N,act,pad', [(16384, 100, 95, 0, 0), (10001, 100, 95, 2, 0), (20000, 36, 48, 1, 5),
                                           (9000, 64, 128, 0, 0)])
def test_dense_forward_tensor_core_path(vms, B, K, N, act, pad):
    """Large batches route to the tcgen05 3 x TF32 GEMM (csrc/gemm_tc.cu): float32-grade parity with the float64
    product (the TF32 split keeps 22 mantissa bits; the TMEM accumulation adds ~5e-6 of the dot product's scale), a
    ragged last tile, padded K / N, a strided output, and agreement with the FFMA kernel on a small-batch slice."""
    v = vms
    c = v._abi.ctx()
    rng = np.random.default_rng(B + K + N)
    x = rng.normal(size=(B, K)).astype(np.float32)
    W = (rng.normal(size=(K, N)) / np.sqrt(K)).astype(np.float32)
    b = rng.normal(size=N).astype(np.float32)
    dx, dW, db = T(v, x), T(v, W), T(v, b)
    ldo = N + pad
    out = v.Tensor.zeros((B, ldo))
    before = v._abi.launch_count()
    c.lib.vms_dense_forward(dx.ptr, K, dW.ptr, db.ptr, B, K, N, act, None, 0, None, 0, out.ptr, ldo, c.stream)
    assert v._abi.launch_count() - before == 1
    got = out.numpy()
    pre = x.astype(np.float64) @ W.astype(np.float64) + b
    want = [pre, np.maximum(pre, 0), np.tanh(pre)][act]
    scale = np.sqrt((x.astype(np.float64)**2) @ (W.astype(np.float64)**2))   # magnitude of the terms of each dot product
    assert np.all(np.abs(got[:, :N] - want) <= 1e-5 * np.abs(want) + 8e-6 * scale + 1e-6), np.abs(got[:, :N] - want).max()
    if pad:
        assert np.all(got[:, N:] == 0)   # the strided form must not touch the padding columns
    # the FFMA kernel (small-batch path) on the first rows gives the same numbers to float32 accuracy
    small = v.Tensor((1000, N))
    c.lib.vms_dense_forward(dx.ptr, K, dW.ptr, db.ptr, 1000, K, N, act, None, 0, None, 0, small.ptr, N, c.stream)
    assert_close(got[:1000, :N], small.numpy(), rtol=2e-5, atol=2e-5, what='tcgen05 vs FFMA dense forward')


def test_dense_ones_input_and_conditional(vms):
    v = vms
    c = v._abi.ctx()
    rng = np.random.default_rng(5)
    B, N, Cn = 300, 70, 3
    W = rng.normal(size=(1, N)).astype(np.float32)
    b = rng.normal(size=N).astype(np.float32)
    cond = rng.normal(size=(B, Cn)).astype(np.float32)
    Wc = rng.normal(size=(Cn, N)).astype(np.float32)
    dW, db, dc, dWc = (T(v, a) for a in (W, b, cond, Wc))
    out = v.Tensor((B, N))
    c.lib.vms_dense_forward(None, 1, dW.ptr, db.ptr, B, 1, N, 2, dc.ptr, Cn, dWc.ptr, Cn, out.ptr, N, c.stream)
    assert_close(out.numpy(), np.tanh(W[0] + b + cond.astype(np.float64) @ Wc), rtol=1e-5, atol=1e-6, what='ones+cond')
    with pytest.raises(ValueError):
        c.lib.vms_dense_forward(None, 1, dW.ptr, db.ptr, B, 2, N, 0, None, 0, None, 0, out.ptr, N, c.stream)


# ------------------------------------------------------------------------------------------------ K4/K5: log-probs
def test_blockwise_log_prob_and_params(vms):
    v = vms
    c = v._abi.ctx()
    rng = np.random.default_rng(6)
    B = 5000
    # dofs: Normal, VonMises, Normal, VonMises, contiguous per-dof parameter groups (dists.py:210)
    kinds = ['normal', 'vonmises', 'normal', 'vonmises']
    params = rng.normal(0, 2, (B, 10)).astype(np.float32)
    params[:50, 4] = rng.uniform(5, 60, 50)  # large concentrations: second Chebyshev branch of i0e
    x = rng.uniform(-np.pi, np.pi, (B, 4)).astype(np.float32)
    want = odists.independent_blockwise_log_prob(x.astype(np.float64), params.astype(np.float64), kinds)
    i32 = lambda a: (C.c_int32 * len(a))(*a)
    kind, loc, loc2, sc = i32([0, 1, 0, 1]), i32([0, 2, 5, 7]), i32([-1, 3, -1, 8]), i32([1, 4, 6, 9])
    dX, dP = T(v, x), T(v, params)
    lp = v.Tensor((B, ))
    c.lib.vms_blockwise_log_prob(dX.ptr, 4, dP.ptr, 10, B, 4, kind, loc, loc2, sc, 2, lp.ptr, 0, c.stream)
    assert_close(lp.numpy(), want, rtol=1e-5, atol=1e-5, what='blockwise log_prob')
    L, S = v.Tensor((B, 4)), v.Tensor((B, 4))
    c.lib.vms_blockwise_params(dP.ptr, 10, B, 4, kind, loc, loc2, sc, 2, L.ptr, S.ptr, c.stream)
    t = odists.param_transform('vonmises', params[:, 2:5].astype(np.float64))
    assert_close(L.numpy()[:, 1], t['loc'], rtol=1e-6, atol=1e-6, what='vonmises loc')
    assert_close(S.numpy()[:, 1], t['concentration'], rtol=1e-6, atol=1e-7, what='vonmises concentration')


def test_param_transform_known_answers(vms):
    """tests/test_dists.py:15-30 through the host API."""
    d = vms.dists
    t = d.make_param_transform(d.Normal)(np.zeros((1, 2), np.float32))
    assert t['loc'].numpy()[0] == 0 and np.isclose(t['scale'].numpy()[0], np.log(2.0), rtol=1e-6)
    t = d.make_param_transform(d.VonMises)(np.array([[0.0, -1.0, 0.0]], np.float32))
    assert np.isclose(t['loc'].numpy()[0], np.pi, rtol=1e-6)
    assert np.isclose(t['concentration'].numpy()[0], np.log(2.0), rtol=1e-6)


def test_normal_sample_log_prob_and_backward(vms):
    v = vms
    c = v._abi.ctx()
    rng = np.random.default_rng(7)
    B, D = 4097, 6
    p = rng.normal(size=(B, 2 * D)).astype(np.float32)
    eps = rng.normal(size=(B, D)).astype(np.float32)
    dP, dE = T(v, p), T(v, eps)
    z, lp = v.Tensor((B, D)), v.Tensor((B, ))
    c.lib.vms_normal_sample_log_prob(dP.ptr, 2 * D, 0, D, 1, dE.ptr, B, D, z.ptr, D, lp.ptr, c.stream)
    loc, scale = odists.independent_normal_params(p, D)
    zo = odists.normal_sample(loc, scale, eps)
    assert_close(z.numpy(), zo, rtol=1e-6, atol=1e-6, what='reparameterised sample')
    assert_close(lp.numpy(), odists.normal_log_prob(zo.astype(np.float64), loc.astype(np.float64),
                                                    scale.astype(np.float64)).sum(-1), rtol=1e-5, atol=1e-5, what='logq')
    g = rng.normal(size=B).astype(np.float32)
    x = rng.normal(size=(B, D)).astype(np.float32)
    dG, dX = T(v, g), T(v, x)
    gx, gp = v.Tensor((B, D)), v.Tensor((B, 2 * D))
    c.lib.vms_normal_log_prob_backward(dX.ptr, D, dP.ptr, 2 * D, 0, D, 1, dG.ptr, B, D, gx.ptr, D, 0, gp.ptr, 2 * D,
                                       c.stream)
    ox, ol, os_ = ovae._normal_lp_bwd(x.astype(np.float64), loc.astype(np.float64), scale.astype(np.float64),
                                      g.astype(np.float64))
    assert_close(gx.numpy(), ox, rtol=1e-5, atol=1e-5, what='d logp / dx')
    assert_close(gp.numpy()[:, :D], ol, rtol=1e-5, atol=1e-5, what='d logp / dloc')
    assert_close(gp.numpy()[:, D:], os_ * orqs.sigmoid(p[:, D:].astype(np.float64)), rtol=1e-5, atol=1e-5, what='d/draw')


@pytest.mark.parametrize('B', [1, 100, 8192, 8193, 300000])
def test_kl_and_mean_reductions(vms, B):
    v = vms
    c = v._abi.ctx()
    rng = np.random.default_rng(B)
    a, b = rng.normal(size=B).astype(np.float32), rng.normal(size=B).astype(np.float32)
    dA, dB = T(v, a), T(v, b)
    out = v.Tensor((1, ))
    c.lib.vms_kl_mean(dA.ptr, dB.ptr, B, 2.5, out.ptr, c.stream)
    ref = 2.5 * np.mean(a.astype(np.float64) - b)
    assert abs(out.numpy()[0] - ref) <= 1e-5 * max(1.0, abs(ref)) + 3e-7 * np.sqrt(B)
    first = out.numpy()[0]
    c.lib.vms_kl_mean(dA.ptr, dB.ptr, B, 2.5, out.ptr, c.stream)
    assert out.numpy()[0] == first  # deterministic
    c.lib.vms_scaled_mean(dA.ptr, B, -1.0, out.ptr, c.stream)
    assert abs(out.numpy()[0] + np.mean(a.astype(np.float64))) <= 1e-5 + 3e-7 * np.sqrt(B)
    # exact identities of tests/test_losses.py:55-70: weight linearity is a single float32 multiply of the same mean
    o1, o100 = v.Tensor((1, )), v.Tensor((1, ))
    c.lib.vms_kl_mean(dA.ptr, dB.ptr, B, 1.0, o1.ptr, c.stream)
    c.lib.vms_kl_mean(dA.ptr, dB.ptr, B, 100.0, o100.ptr, c.stream)
    assert o100.numpy()[0] == np.float32(100.0) * o1.numpy()[0]


# ------------------------------------------------------------------------------------------------ K6: selection
def _select(v, coords, ref, cutoff, k, box=None, info=None, splits=None, per_row=False):
    c = v._abi.ctx()
    if splits is None:
        B, N = coords.shape[0], coords.shape[1]
    else:
        B, N = len(splits) - 1, 0
    dC = T(v, coords.reshape(-1, 3))
    dS = None if splits is None else T(v, splits, np.int64)
    dR = T(v, ref)
    dB = None if box is None else T(v, box)
    dI = None if info is None else T(v, info.reshape(-1, info.shape[-1]))
    P = 0 if info is None else info.shape[-1]
    oxyz, oidx = v.Tensor((B, k, 3)), v.Tensor((B, k), np.int32)
    oinfo = None if info is None else v.Tensor((B, k, P))
    c.lib.vms_dist_select(dC.ptr, None if dS is None else dS.ptr, B, N, dR.ptr, None if dB is None else dB.ptr,
                          1 if per_row else 0, float(np.float32(cutoff**2)), k, None if dI is None else dI.ptr, P, oxyz.ptr,
                          None if oinfo is None else oinfo.ptr, oidx.ptr, c.stream)
    fast = v.Tensor((B, k, 3))  # same call without indices: the single-pass candidate path
    c.lib.vms_dist_select(dC.ptr, None if dS is None else dS.ptr, B, N, dR.ptr, None if dB is None else dB.ptr,
                          1 if per_row else 0, float(np.float32(cutoff**2)), k, None, 0, fast.ptr, None, None, c.stream)
    return oxyz.numpy(), (None if oinfo is None else oinfo.numpy()), oidx.numpy(), fast.numpy()


@pytest.mark.parametrize('N,k,cutoff', [(100, 50, 3.0), (1000, 10, 3.0), (1000, 50, 1.0), (5000, 64, 6.0), (30, 50, 3.0),
                                        (3000, 50, 100.0)])
def test_dist_select_bit_exact(vms, N, k, cutoff):
    rng = np.random.default_rng(N + k)
    B, L = 24, 10.0
    coords = rng.uniform(0, L, (B, N, 3)).astype(np.float32)
    coords[:, 5] = coords[:, 3]  # exact duplicates: ties resolved by the lower index
    ref = rng.uniform(0, L, (B, 3)).astype(np.float32)
    info = rng.normal(size=(B, N, 2)).astype(np.float32)
    box = np.array([L, L, L], np.float32)
    for bx, per_row in ((None, False), (box, False), (np.tile(box * np.float32(1.1), (B, 1)), True)):
        want = omap.distance_selection(coords, ref, cutoff, k, box_lengths=bx, particle_info=info, return_indices=True)
        xyz, oinfo, idx, fast = _select(vms, coords, ref, cutoff, k, box=bx, info=info, per_row=per_row)
        assert np.array_equal(idx, want[2]), 'neighbour indices differ'
        assert np.array_equal(xyz, want[0]) and np.array_equal(oinfo, want[1])
        assert np.array_equal(fast, want[0])


def test_dist_select_ragged_with_empty_row(vms):
    """tests/test_mappings.py:88-98 and tests/test_models.py:265-308: 0..N particles per row."""
    rng = np.random.default_rng(3)
    lens = np.array([7, 0, 60, 49, 1, 0, 200, 50])
    rows = [rng.uniform(0, 10, (n, 3)).astype(np.float32) for n in lens]
    irows = [rng.normal(size=(n, 3)).astype(np.float32) for n in lens]
    ref = rng.uniform(0, 10, (len(lens), 3)).astype(np.float32)
    box = np.array([10.0, 10.0, 10.0], np.float32)
    splits = np.concatenate([[0], np.cumsum(lens)])
    want = omap.distance_selection(rows, ref, 3.0, 50, box_lengths=box, particle_info=irows, return_indices=True)
    xyz, oinfo, idx, fast = _select(vms, np.concatenate(rows), ref, 3.0, 50, box=box, info=np.concatenate(irows),
                                    splits=splits)
    assert np.array_equal(xyz, want[0]) and np.array_equal(oinfo, want[1]) and np.array_equal(fast, want[0])
    assert np.array_equal(idx, want[2])
    assert np.all(xyz[1] == 0) and np.all(xyz[5] == 0)


def test_dist_select_layer_identities(vms):
    """Reference test identities through the layer API (tests/test_mappings.py:62-86)."""
    m = vms.mappings
    rng = np.random.default_rng(4)
    coords = rng.uniform(0, 10, (8, 100, 3)).astype(np.float32)
    ref = rng.uniform(0, 10, (8, 1, 3)).astype(np.float32)
    box = np.array([10.0, 10.0, 10.0], np.float32)
    ds = m.DistanceSelection(3.0)
    assert ds.sq_cut == 9.0
    out = ds(coords, ref).numpy()
    assert out.shape == (8, 50, 3)
    stored = m.DistanceSelection(3.0, box_lengths=box)(coords, ref).numpy()
    per_call = ds(coords, ref, box_lengths=np.tile(box, (8, 1))).numpy()
    assert np.array_equal(stored, per_call) and not np.array_equal(stored, out)
    info = rng.normal(size=(8, 100, 2)).astype(np.float32)
    sel, sinfo = m.DistanceSelection(3.0, max_included=10)(coords, ref, particle_info=info)
    assert sel.shape == (8, 10, 3) and sinfo.shape == (8, 10, 2)
    rag = m.DistanceSelection(3.0)([coords[0, :5], coords[1, :0], coords[2]], ref[:3]).numpy()
    assert rag.shape == (3, 50, 3) and np.all(rag[1] == 0)


# ------------------------------------------------------------------------------------------------ K7: MC acceptance
@pytest.mark.parametrize('tag', ['c4a', 'flow'])
def test_mc_accept_bit_exact_vs_reference_mcmc(vms, tag):
    """Decisions from the reference's own mcmc.py (golden) reproduced by vms_mc_accept from the same inputs."""
    v = vms
    c = v._abi.ctx()
    g = np.load(os.path.join(GOLD, 'mcmc_reference_%s.npz' % tag))
    n_acc = v.Tensor.zeros((1, ), np.uint64)
    for s in range(5):
        B, D = g['x_old_%d' % s].shape
        d = dict(e_new=T(v, g['e_new_%d' % s], np.float64), e_old=T(v, g['e_old_%d' % s], np.float64),
                 fwd=T(v, g['fwd_%d' % s]), rev=T(v, g['rev_%d' % s]), lr=T(v, g['log_rand_%d' % s], np.float64),
                 x_old=T(v, g['x_old_%d' % s]), x_new=T(v, g['x_new_%d' % s]))
        e_out, acc = v.Tensor((B, ), np.float64), v.Tensor((B, ), np.uint8)
        c.lib.vms_mc_accept(d['e_new'].ptr, d['e_old'].ptr, d['fwd'].ptr, d['rev'].ptr, d['lr'].ptr, B, D,
                            d['x_old'].ptr, d['x_new'].ptr, e_out.ptr, acc.ptr, n_acc.ptr, c.stream)
        assert np.array_equal(acc.numpy().astype(bool), g['acc_%d' % s])
        assert np.array_equal(d['x_new'].numpy(), g['configs_%d' % s])
        assert np.array_equal(e_out.numpy(), g['energies_%d' % s])
    assert float(n_acc.numpy()[0]) == float(g['num_acc'])


def test_mc_accept_f32_bit_exact_vs_reference_mcmc_c4b(vms):
    """A float32 energy callback (tfp log_prob, MC notebook cell 38) makes NumPy evaluate mcmc.py:116 in float32:
    vms_mc_accept_f32 reproduces the reference driver's decisions from the same inputs; vms_energy_gmm against the oracle."""
    v = vms
    c = v._abi.ctx()
    g = np.load(os.path.join(GOLD, 'mcmc_reference_c4b.npz'))
    n_acc = v.Tensor.zeros((1, ), np.uint64)
    lw, loc, sc = T(v, np.log(omc.GMM_PROBS)), T(v, omc.GMM_LOCS), T(v, omc.GMM_SCALES)
    for s in range(5):
        B, D = g['x_old_%d' % s].shape
        assert g['e_new_%d' % s].dtype == np.float32
        d = dict(e_new=T(v, g['e_new_%d' % s]), e_old=T(v, g['e_old_%d' % s]), fwd=T(v, g['fwd_%d' % s]),
                 rev=T(v, g['rev_%d' % s]), lr=T(v, g['log_rand_%d' % s], np.float64), x_old=T(v, g['x_old_%d' % s]),
                 x_new=T(v, g['x_new_%d' % s]))
        e_dev = v.Tensor((B, ), np.float32)
        c.lib.vms_energy_gmm(d['x_new'].ptr, B, D, 3, lw.ptr, loc.ptr, sc.ptr, e_dev.ptr, c.stream)
        np.testing.assert_allclose(e_dev.numpy(), g['e_new_%d' % s], rtol=2e-6, atol=2e-6)
        e_out, acc = v.Tensor((B, ), np.float32), v.Tensor((B, ), np.uint8)
        c.lib.vms_mc_accept_f32(d['e_new'].ptr, d['e_old'].ptr, d['fwd'].ptr, d['rev'].ptr, d['lr'].ptr, B, D,
                                d['x_old'].ptr, d['x_new'].ptr, e_out.ptr, acc.ptr, n_acc.ptr, c.stream)
        assert np.array_equal(acc.numpy().astype(bool), g['acc_%d' % s])
        assert np.array_equal(d['x_new'].numpy(), g['configs_%d' % s])
        assert np.array_equal(e_out.numpy(), g['energies_%d' % s])
    assert float(n_acc.numpy()[0]) == float(g['num_acc'])


def test_mc_accept_random_and_energy(vms):
    v = vms
    c = v._abi.ctx()
    rng = np.random.default_rng(11)
    B, D = 100003, 6
    x_old, x_new = rng.normal(size=(B, D)).astype(np.float32), rng.normal(size=(B, D)).astype(np.float32)
    means = np.linspace(-2, 2, D)
    dE_old, dE_new = v.Tensor((B, ), np.float64), v.Tensor((B, ), np.float64)
    dXo, dXn, dM = T(v, x_old), T(v, x_new), T(v, means, np.float64)
    c.lib.vms_energy_quadratic(dXo.ptr, B, D, dM.ptr, dE_old.ptr, c.stream)
    c.lib.vms_energy_quadratic(dXn.ptr, B, D, dM.ptr, dE_new.ptr, c.stream)
    e_old, e_new = omc.quadratic_energy(x_old), omc.quadratic_energy(x_new)
    assert np.array_equal(dE_old.numpy(), e_old) and np.array_equal(dE_new.numpy(), e_new)
    fwd, rev = rng.normal(size=B).astype(np.float32) * 5, rng.normal(size=B).astype(np.float32) * 5
    log_u = np.log(rng.random(B))
    want = omc.accept(e_new, e_old, fwd, rev, log_u)
    acc, e_out, n_acc = v.Tensor((B, ), np.uint8), v.Tensor((B, ), np.float64), v.Tensor.zeros((1, ), np.uint64)
    dF, dRv, dU = T(v, fwd), T(v, rev), T(v, log_u, np.float64)  # keep alive: freed tensors return to the pool
    c.lib.vms_mc_accept(dE_new.ptr, dE_old.ptr, dF.ptr, dRv.ptr, dU.ptr, B, D, dXo.ptr,
                        dXn.ptr, e_out.ptr, acc.ptr, n_acc.ptr, c.stream)
    assert np.array_equal(acc.numpy().astype(bool), want) and int(n_acc.numpy()[0]) == int(want.sum())
    assert np.array_equal(dXn.numpy(), np.where(want[:, None], x_new, x_old))
    assert np.array_equal(e_out.numpy(), np.where(want, e_new, e_old))


# ------------------------------------------------------------------------------------------------ K8: Adam
def test_adam_matches_oracle(vms):
    v = vms
    c = v._abi.ctx()
    rng = np.random.default_rng(12)
    n = 44396
    theta = rng.normal(size=n).astype(np.float32)
    th_o, m_o, v_o = theta.copy(), np.zeros(n, np.float32), np.zeros(n, np.float32)
    dT, dM, dV = T(v, theta), v.Tensor.zeros((n, )), v.Tensor.zeros((n, ))
    for t in range(1, 4):
        parts = rng.normal(size=(3, n)).astype(np.float32)
        dP = T(v, parts)
        c.lib.vms_adam_step(dT.ptr, dP.ptr, 3, 0.5, dM.ptr, dV.ptr, n, t, 1e-3, 0.9, 0.999, 1e-7, c.stream)
        g = ((parts[0] + parts[1]) + parts[2]) * np.float32(0.5)
        ovae.adam_step(th_o, g, m_o, v_o, t)
    assert_close(dT.numpy(), th_o, rtol=2e-6, atol=2e-7, what='adam theta')  # FMA contraction vs NumPy mul+add
    assert_close(dM.numpy(), m_o, rtol=2e-6, atol=2e-7, what='adam m')


# ------------------------------------------------------------------------------------------------ blockwise sampling
def test_blockwise_sample_normal_and_von_mises(vms):
    """vms_blockwise_sample (tfp Blockwise.sample over the per-dof distributions of dists.py:210-217 / :602-610): Normal dofs
    reproduce eps * scale + loc for a given eps; von Mises dofs (tfp's Best-Fisher rejection sampler on a Philox stream)
    are checked distributionally against scipy's von Mises CDF (Kolmogorov-Smirnov) and the analytic mean resultant
    length I1(k) / I0(k); samples are reproducible for a seed and wrapped to [-pi, pi)."""
    import ctypes as C
    from scipy import special, stats
    v = vms
    c = v._abi.ctx()
    B, D = 100000, 4
    rng = np.random.default_rng(8)
    kappa = np.array([0.01, 1.0, 50.0], np.float32)
    loc_vm = np.array([0.3, -2.5, 3.0], np.float32)
    # params: [sin, cos, concentration] x 3 von Mises dofs, then [loc, scale] of one Normal dof; identity scale mode
    P = np.zeros((B, 11), np.float32)
    for i in range(3):
        P[:, 3 * i], P[:, 3 * i + 1], P[:, 3 * i + 2] = 2.0 * np.sin(loc_vm[i]), 2.0 * np.cos(loc_vm[i]), kappa[i]
    P[:, 9] = rng.normal(size=B).astype(np.float32)
    P[:, 10] = rng.uniform(0.5, 2.0, B).astype(np.float32)
    i32 = lambda a: (C.c_int32 * len(a))(*a)
    kind, loc, loc2, sc = i32([1, 1, 1, 0]), i32([0, 3, 6, 9]), i32([1, 4, 7, -1]), i32([2, 5, 8, 10])
    eps = rng.standard_normal((B, D), dtype=np.float32)
    Pt, et = v.Tensor.from_numpy(P), v.Tensor.from_numpy(eps)

    def draw(seed, with_eps):
        out = v.Tensor((B, D))
        c.lib.vms_blockwise_sample(Pt.ptr, 11, B, D, kind, loc, loc2, sc, 0, et.ptr if with_eps else None, D, seed, out.ptr,
                                   D, c.stream)
        return out.numpy()

    x = draw(1234, True)
    assert np.array_equal(x, draw(1234, True)) and not np.array_equal(x[:, :3], draw(1235, True)[:, :3])
    assert_close(x[:, 3], eps[:, 3] * P[:, 10] + P[:, 9], rtol=1e-6, atol=1e-6, what='Normal dof from eps')
    assert np.all(x[:, :3] >= -np.pi - 1e-6) and np.all(x[:, :3] <= np.pi + 1e-6)
    for i in range(3):
        dlt = np.angle(np.exp(1j * (x[:, i].astype(np.float64) - loc_vm[i])))
        ks = stats.kstest(dlt, stats.vonmises(kappa[i]).cdf).statistic
        assert ks < 2.2 / np.sqrt(B), ('von Mises KS', i, ks)
        R = np.abs(np.mean(np.exp(1j * dlt)))
        want = special.i1e(kappa[i]) / special.i0e(kappa[i])
        assert abs(R - want) < 5.0 / np.sqrt(B), ('mean resultant length', i, R, want)
    n = draw(77, False)[:, 3]  # Normal dof from the Philox stream
    zs = (n - P[:, 9]) / P[:, 10]
    assert abs(zs.mean()) < 5.0 / np.sqrt(B) and abs(zs.var() - 1.0) < 0.03
    assert stats.kstest(zs, 'norm').statistic < 2.2 / np.sqrt(B)


def test_deterministic_log_prob_on_device(vms):
    v = vms
    loc = np.arange(12, dtype=np.float32).reshape(4, 3)
    d = v._protocols.Deterministic(v.Tensor.from_numpy(loc))
    x = loc.copy()
    x[2, 1] += 1.0
    lp = d.log_prob(v.Tensor.from_numpy(x)).numpy()
    assert np.array_equal(lp, np.array([0.0, 0.0, -np.inf, 0.0], np.float32))


@pytest.mark.parametrize('D,B,accumulate', [(6, 3 * 65536 + 77, 0), (3, 65536 + 256, 1), (2, 65536, 0), (8, 70001, 0)])
def test_normal_rows_log_prob_streaming_ring(vms, D, B, accumulate):
    """Planar Normal rows at streaming sizes: full 256-row chunks go through the bulk-copy ring (`normal_rows_lp_ring_kernel`),
    the ragged tail through the thread-per-row kernel; float64 evaluation of tfp's Normal.log_prob with the softplus + eps
    scale of dists.py:56-78 on every row."""
    v = vms
    c = v._abi.ctx()
    rng = np.random.default_rng(D * 1000 + B % 1000)
    x = rng.standard_normal((B, D), dtype=np.float32)
    p = rng.standard_normal((B, 2 * D), dtype=np.float32)
    lp0 = rng.standard_normal(B).astype(np.float32)
    xd, pd, lp = v.Tensor.from_numpy(x), v.Tensor.from_numpy(p), v.Tensor.from_numpy(lp0)
    i32 = lambda a: (C.c_int32 * len(a))(*a)
    c.lib.vms_blockwise_log_prob(xd.ptr, D, pd.ptr, 2 * D, B, D, i32([0] * D), i32(list(range(D))), i32([-1] * D),
                                 i32(list(range(D, 2 * D))), 2, lp.ptr, accumulate, c.stream)
    x64, p64 = x.astype(np.float64), p.astype(np.float64)
    sc = np.logaddexp(0.0, p64[:, D:]) + float(np.finfo(np.float32).eps)
    want = (-0.5 * ((x64 - p64[:, :D]) / sc)**2 - 0.5 * np.log(2 * np.pi) - np.log(sc)).sum(axis=1)
    if accumulate:
        want = want + lp0
    assert_close(lp.numpy(), want, rtol=1e-5, atol=2e-5, what='normal rows log_prob, ring + tail')
