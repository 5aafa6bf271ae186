"""GPU tests of the geometric-algebra attention path (mappings.py:480-762, models.py:470-572): kernels and
layers against the NumPy restatement `oracle/gaa.py` (parity unpinned: the arithmetic lives in the un-vendored
geometric-algebra-attention package), the one-kernel forward against the op-by-op path, reverse mode against float64 finite
differences of the oracle, and ports of the reference's tests (tests/test_mappings.py:101-159, tests/test_models.py:265-308).
"""
import os

import numpy as np
import pytest

from helpers import assert_close
from oracle import gaa as ogaa

pytestmark = pytest.mark.gpu


def _assign(t, a):
    import vaemolsim_b200 as v
    c = v._abi.ctx()
    a = np.ascontiguousarray(a, np.float32)
    assert tuple(a.shape) == tuple(t.shape) or a.size == t.size, (a.shape, t.shape)
    c.lib.vms_memcpy_h2d(t.ptr, a.ctypes.data, a.nbytes, c.stream)
    c.synchronize()


def _load_mlp(seq, w):
    """oracle [W1, b1, (gamma, beta,) W2, b2] -> Sequential [Dense, (LayerNormalization, Activation,) Dense]."""
    import vaemolsim_b200._protocols as P
    it = iter(w)
    for lay in seq.layers:
        if isinstance(lay, P.Dense):
            _assign(lay.kernel, next(it))
            _assign(lay.bias, next(it))
        elif isinstance(lay, P.LayerNormalization):
            _assign(lay.gamma, next(it))
            _assign(lay.beta, next(it))


def _load_attention(va, w):
    for k in range(2):
        _assign(va.merge_kernels[k], w['merge'][k])
        _assign(va.join_kernels[k], w['join'][k])
    _load_mlp(va.score_net, w['score'])
    _load_mlp(va.value_net, w['value'])


def _load_embedding(pe, w):
    _assign(pe.info_net.kernel, w['info'][0])
    _assign(pe.info_net.bias, w['info'][1])
    for blk, wb in zip(pe.block_list, w['blocks']):
        _load_attention(blk.attn, wb)
        _load_mlp(blk.nonlinearity, wb['nonlin'])
    _load_attention(pe.final_attn, w['final'])


def _cloud(rng, B, n, P, n_pad=0, dtype=np.float32):
    r = rng.uniform(-3, 3, (B, n, 3)).astype(dtype)
    info = np.round(rng.uniform(0, 1, (B, n, P))).astype(dtype)
    for b in range(B):  # DistanceSelection-style zero padding at the end of a cloud (b = 1: everything masked)
        k = n if b == 1 else min(n_pad + b, n - 1) if n_pad else 0
        if k:
            r[b, n - k:] = 0
            info[b, n - k:] = 0
    return r, info


class _FusedOff(object):
    def __enter__(self):
        self.old = os.environ.get('VMS_GAA_FUSED')
        os.environ['VMS_GAA_FUSED'] = '0'

    def __exit__(self, *exc):
        if self.old is None:
            del os.environ['VMS_GAA_FUSED']
        else:
            os.environ['VMS_GAA_FUSED'] = self.old


# ------------------------------------------------------------------------------------------------------ kernels
def test_pair_invariants_bit_exact_and_zero_mask(vms):
    v = vms
    c = v._abi.ctx()
    rng = np.random.default_rng(0)
    B, n = 5, 13
    r, _ = _cloud(rng, B, n, 2, n_pad=3)
    rd = v.Tensor.from_numpy(r)
    out = v.Tensor((B * n * n, 2))
    c.lib.vms_gaa_pair_invariants(rd.ptr, B, n, out.ptr, c.stream)
    want = ogaa.pair_invariants(r)
    assert want.dtype == np.float32
    assert np.array_equal(out.numpy().reshape(B, n, n, 2), want)
    assert np.array_equal(v.mappings.zero_mask(rd).numpy().astype(bool), ogaa.keras_mask(r))


@pytest.mark.parametrize('R,H,act', [(1000, 40, 'relu'), (37, 7, 'tanh'), (4099, 128, None), (5, 20, 'relu')])
def test_layernorm_forward_backward(vms, R, H, act):
    """y = act(LayerNormalization(x)) against the float64 oracle; g_x / g_gamma / g_beta against the analytic float64
    gradient of the same expression."""
    v = vms
    import vaemolsim_b200._protocols as P
    c = v._abi.ctx()
    rng = np.random.default_rng(R + H)
    x = rng.normal(0.3, 1.5, (R, H)).astype(np.float32)
    gamma = (1 + 0.3 * rng.normal(size=H)).astype(np.float32)
    beta = rng.normal(0, 0.2, H).astype(np.float32)
    gy = rng.normal(size=(R, H)).astype(np.float32)
    code = P.ACT[act]
    xd, gd, bd, gyd = [v.Tensor.from_numpy(a) for a in (x, gamma, beta, gy)]
    y, stats = v.Tensor((R, H)), v.Tensor((R, 2))
    c.lib.vms_layernorm_forward(xd.ptr, H, R, H, gd.ptr, bd.ptr, 1e-3, code, y.ptr, H, stats.ptr, c.stream)
    x64, g64, b64 = x.astype(np.float64), gamma.astype(np.float64), beta.astype(np.float64)
    pre = ogaa.layer_norm(x64, g64, b64)
    want = ogaa._act(act)(pre)
    assert_close(y.numpy(), want, rtol=1e-5, atol=2e-6, what='layer norm forward')
    gx, gg, gb = v.Tensor.zeros((R, H)), v.Tensor.zeros((H, )), v.Tensor.zeros((H, ))
    ws = v.Tensor((max(int(c.lib.vms_layernorm_backward_workspace(R, H)) // 4, 1), ))
    c.lib.vms_layernorm_backward(xd.ptr, H, R, H, gd.ptr, stats.ptr, code, y.ptr, H, gyd.ptr, H, gx.ptr, H, gg.ptr, gb.ptr,
                                 ws.ptr, c.stream)
    dact = {None: np.ones_like(pre), 'relu': (pre > 0).astype(np.float64), 'tanh': 1 - np.tanh(pre)**2}[act]
    gp = gy.astype(np.float64) * dact
    mean = x64.mean(-1, keepdims=True)
    rstd = 1 / np.sqrt(((x64 - mean)**2).mean(-1, keepdims=True) + 1e-3)
    xh = (x64 - mean) * rstd
    dxh = gp * g64
    want_gx = rstd * (dxh - dxh.mean(-1, keepdims=True) - xh * (dxh * xh).mean(-1, keepdims=True))
    assert_close(gx.numpy(), want_gx, rtol=2e-5, atol=2e-5, what='layer norm g_x')
    sc = np.sqrt(R)
    assert_close(gg.numpy(), (gp * xh).sum(0), rtol=1e-5, atol=1e-5 * sc, what='layer norm g_gamma')
    assert_close(gb.numpy(), gp.sum(0), rtol=1e-5, atol=1e-5 * sc, what='layer norm g_beta')


# ------------------------------------------------------------------------------------------------------- layers
def _make_attention(v, D, H, reduce, act):
    import vaemolsim_b200._protocols as P
    M = v.mappings
    va = M.VectorAttention(P.Sequential([P.Dense(H, activation=act), P.Dense(1)]), M._mlp_ln(H, D, act), reduce=reduce,
                           merge_fun='concat', join_fun='concat', rank=2)
    va.build([(None, None, 3), (None, None, D)])
    va.built = True
    return va


@pytest.mark.parametrize('B,n,D,H,reduce,act,masked', [
    (4, 6, 9, 40, False, 'relu', False), (4, 6, 9, 40, True, 'relu', True), (3, 50, 20, 40, False, 'relu', True),
    (3, 50, 20, 40, True, 'relu', True), (2, 100, 20, 40, True, 'relu', False), (5, 10, 20, 20, False, 'tanh', True),
    (3, 17, 32, 64, True, 'tanh', True), (2, 300, 12, 24, False, 'relu', True)])
def test_vector_attention_both_paths_match_oracle(vms, B, n, D, H, reduce, act, masked):
    """One VectorAttention layer (rank 2, concat / concat): the op-by-op path and the one-kernel forward against the float64
    oracle, with zero-padded clouds (one cloud fully masked: uniform attention, as softmax over equal -1e9 logits gives)."""
    v = vms
    rng = np.random.default_rng(B * n + D)
    w = ogaa.init_attention(rng, D, H)
    r, _ = _cloud(rng, B, n, 2, n_pad=2 if masked else 0)
    vals = rng.normal(0, 1, (B, n, D)).astype(np.float32)
    va = _make_attention(v, D, H, reduce, act)
    _load_attention(va, w)
    mask = v.mappings.zero_mask(v.Tensor.from_numpy(r)) if masked else None
    want = ogaa.vector_attention(r.astype(np.float64), vals.astype(np.float64), ogaa.cast(w, np.float64), reduce, act,
                                 ogaa.keras_mask(r) if masked else None)
    scale = np.abs(want).max()
    with _FusedOff():
        out_ops = va.call([v.Tensor.from_numpy(r), v.Tensor.from_numpy(vals)], mask=mask).numpy()
    assert_close(out_ops, want, rtol=1e-5, atol=1e-5 * scale, what='op-by-op attention vs float64 oracle')
    n0 = v._abi.launch_count()
    out_f = va.call([v.Tensor.from_numpy(r), v.Tensor.from_numpy(vals)], mask=mask).numpy()
    assert v._abi.launch_count() - n0 == 2, 'the fused forward is two launches: weight image + attention kernel'
    assert_close(out_f, want, rtol=1e-5, atol=1e-5 * scale, what='fused attention vs float64 oracle')
    assert_close(out_f, out_ops, rtol=1e-5, atol=1e-5 * scale, what='fused vs op-by-op')


def test_attention_block_and_particle_embedding_match_oracle(vms):
    """Shapes of tests/test_mappings.py:101-125 (block: coords (4, 6, 3), info (4, 6, 9); embedding: (4, 100, 3), (4, 100,
    9) -> (4, 20)), values against the float64 oracle, both paths."""
    v = vms
    M = v.mappings
    rng = np.random.default_rng(7)
    r, info = _cloud(rng, 4, 6, 9)
    blk = M.AttentionBlock()
    out = blk([v.Tensor.from_numpy(r), v.Tensor.from_numpy(info)])
    assert out.shape == info.shape
    wb = ogaa.init_block(rng, 9, 40)
    _load_attention(blk.attn, wb)
    _load_mlp(blk.nonlinearity, wb['nonlin'])
    want = ogaa.attention_block(r.astype(np.float64), info.astype(np.float64), ogaa.cast(wb, np.float64))
    got = blk([v.Tensor.from_numpy(r), v.Tensor.from_numpy(info)]).numpy()
    assert_close(got, want, rtol=1e-5, atol=1e-5 * np.abs(want).max(), what='AttentionBlock')
    blk20 = M.AttentionBlock(hidden_dim=20)
    assert blk20.hidden_dim == 20 and blk20([v.Tensor.from_numpy(r), v.Tensor.from_numpy(info)]).shape == info.shape

    r, info = _cloud(rng, 4, 100, 9, n_pad=40)
    pe = M.ParticleEmbedding(20)
    assert pe.mask_zero
    out = pe(v.Tensor.from_numpy(r), v.Tensor.from_numpy(info))
    assert out.shape == (4, pe.embedding_dim) and len(pe.block_list) == pe.num_blocks and hasattr(out, '_keras_mask')
    w = ogaa.init_embedding(rng, 9, 20)
    _load_embedding(pe, w)
    want = ogaa.particle_embedding(r.astype(np.float64), info.astype(np.float64), ogaa.cast(w, np.float64))
    got = pe(v.Tensor.from_numpy(r), v.Tensor.from_numpy(info)).numpy()
    assert_close(got, want, rtol=2e-5, atol=2e-5 * np.abs(want).max(), what='ParticleEmbedding (fused layers)')
    with _FusedOff():
        got_ops = pe(v.Tensor.from_numpy(r), v.Tensor.from_numpy(info)).numpy()
    assert_close(got_ops, want, rtol=2e-5, atol=2e-5 * np.abs(want).max(), what='ParticleEmbedding (op-by-op)')
    # masking changes the result (tests/test_mappings.py:127-148)
    no_mask = M.ParticleEmbedding(20, mask_zero=False)
    out_no = no_mask(v.Tensor.from_numpy(r), v.Tensor.from_numpy(info))
    assert no_mask.mask is None and not hasattr(out_no, '_keras_mask')
    _load_embedding(no_mask, w)
    out_no = no_mask(v.Tensor.from_numpy(r), v.Tensor.from_numpy(info)).numpy()
    want_no = ogaa.particle_embedding(r.astype(np.float64), info.astype(np.float64), ogaa.cast(w, np.float64), mask_zero=False)
    assert_close(out_no, want_no, rtol=2e-5, atol=2e-5 * np.abs(want_no).max(), what='ParticleEmbedding without mask')
    assert not np.all(out_no == got)


def _flat_params(w, out=None):
    out = [] if out is None else out
    if isinstance(w, dict):
        for k in sorted(w):
            _flat_params(w[k], out)
    elif isinstance(w, list):
        for u in w:
            _flat_params(u, out)
    else:
        out.append(w)
    return out


def test_particle_embedding_reverse_mode_matches_float64_finite_differences(vms):
    """d sum(cot * ParticleEmbedding(coords, info)) / d every weight tensor through the tape (dense, layer norm, pair merge,
    attention reverse-mode kernels) against central differences of the float64 oracle on sampled entries."""
    v = vms
    from vaemolsim_b200 import _autodiff
    M = v.mappings
    rng = np.random.default_rng(11)
    B, n, Pn, E, H = 3, 7, 4, 12, 16
    r, info = _cloud(rng, B, n, Pn, n_pad=1)
    w = ogaa.init_embedding(rng, Pn, E, hidden=H, num_blocks=1)
    cot = rng.normal(size=(B, E)).astype(np.float32)
    pe = M.ParticleEmbedding(E, hidden_dim=H, num_blocks=1, activation='tanh')
    pe(v.Tensor.from_numpy(r), v.Tensor.from_numpy(info))
    _load_embedding(pe, w)
    cot_d = v.Tensor.from_numpy(cot)
    c = v._abi.ctx()
    with _autodiff.Tape() as tape:
        out = pe(v.Tensor.from_numpy(r), v.Tensor.from_numpy(info))
        g = tape.grad(out)
        c.lib.vms_memcpy_d2d(g.ptr, cot_d.ptr, cot.nbytes, c.stream)
        for fn in reversed(tape.ops):
            fn()
    w64 = ogaa.cast(w, np.float64)
    r64, i64 = r.astype(np.float64), info.astype(np.float64)
    f = lambda: float((ogaa.particle_embedding(r64, i64, w64, 'tanh') * cot).sum())
    pairs = [(pe.info_net.kernel, w64['info'][0]), (pe.info_net.bias, w64['info'][1])]
    for va, wa in ((pe.block_list[0].attn, w64['blocks'][0]), (pe.final_attn, w64['final'])):
        pairs += [(va.merge_kernels[k], wa['merge'][k]) for k in range(2)] + [(va.join_kernels[k], wa['join'][k]) for k in range(2)]
        s, vn = va.score_net.layers, va.value_net.layers
        pairs += [(s[0].kernel, wa['score'][0]), (s[0].bias, wa['score'][1]), (s[1].kernel, wa['score'][2]),
                  (s[1].bias, wa['score'][3])]
        pairs += [(vn[0].kernel, wa['value'][0]), (vn[0].bias, wa['value'][1]), (vn[1].gamma, wa['value'][2]),
                  (vn[1].beta, wa['value'][3]), (vn[3].kernel, wa['value'][4]), (vn[3].bias, wa['value'][5])]
    nl = pe.block_list[0].nonlinearity.layers
    wn = w64['blocks'][0]['nonlin']
    pairs += [(nl[0].kernel, wn[0]), (nl[0].bias, wn[1]), (nl[1].gamma, wn[2]), (nl[1].beta, wn[3]), (nl[3].kernel, wn[4]),
              (nl[3].bias, wn[5])]
    h = 1e-6
    checked = 0
    for t, a in pairs:
        assert tape.has(t), 'no gradient reached a weight tensor'
        got = tape.grad(t).numpy().reshape(-1)
        flat = a.reshape(-1)
        for k in rng.choice(flat.size, size=min(4, flat.size), replace=False):
            old = flat[k]
            flat[k] = old + h
            fp = f()
            flat[k] = old - h
            fm = f()
            flat[k] = old
            want = (fp - fm) / (2 * h)
            assert abs(got[k] - want) <= 3e-5 * max(1.0, abs(want)) + 3e-5, (t.shape, k, got[k], want)
            checked += 1
    assert checked > 100
    tape.release()


def test_local_particle_descriptors_and_backmapping_only(vms):
    """Ports of tests/test_mappings.py:151-159 and tests/test_models.py:265-308: DistanceSelection (k = 10) + ParticleEmbedding
    behind LocalParticleDescriptors; BackmappingOnly over ragged FG + CG clouds of 0-65 particles with an
    AutoregressiveBlockwise (3 Normal + 3 von Mises) decoder: call / sample / log_prob / fit / evaluate / predict."""
    v = vms
    import vaemolsim_b200._protocols as P
    M, D, Mo, Lo = v.mappings, v.dists, v.models, v.losses
    v.set_seed(5)
    rng = np.random.default_rng(5)
    c4 = rng.uniform(-5, 5, (4, 100, 3)).astype(np.float32)
    ref4 = rng.uniform(-5, 5, (4, 1, 3)).astype(np.float32)
    info4 = np.round(rng.uniform(0, 1, (4, 100, 9))).astype(np.float32)
    pe = M.ParticleEmbedding(20)
    lpd = M.LocalParticleDescriptors(M.DistanceSelection(3.0, max_included=10, box_lengths=[10.0, 10.0, 10.0]), pe)
    out = lpd(c4, ref4, info4)
    assert out.shape == (4, pe.embedding_dim)
    # the same through the oracle chain (DistanceSelection oracle -> embedding oracle with the layer's own weights)
    from oracle import mappings as omap
    sel, sinfo = omap.distance_selection(c4, ref4, 3.0, 10, box_lengths=np.array([10.0, 10.0, 10.0], np.float32),
                                         particle_info=info4)
    w = ogaa.init_embedding(rng, 9, 20)
    _load_embedding(pe, w)
    want = ogaa.particle_embedding(sel.astype(np.float64), sinfo.astype(np.float64), ogaa.cast(w, np.float64))
    assert_close(lpd(c4, ref4, info4).numpy(), want, rtol=2e-5, atol=2e-5 * np.abs(want).max(), what='descriptors')

    n_b = 200
    ref = rng.uniform(-5, 5, (n_b, 1, 3)).astype(np.float32)
    n_particles = rng.integers(0, 50, size=n_b)
    coords, infos = [], []
    for nfg in n_particles:
        ncg = nfg // 3
        coords.append(np.concatenate([rng.uniform(-5, 5, (nfg, 3)), rng.uniform(-5, 5, (ncg, 3))]).astype(np.float32))
        infos.append(np.concatenate([np.tile([1.0, 0.0], (nfg, 1)), np.tile([0.0, 1.0], (ncg, 1))]).astype(np.float32).reshape(-1, 2))
    all_coords = M.RaggedTensor.from_rows(coords, inner=3)
    all_info = M.RaggedTensor.from_rows(infos, inner=2)
    mask_and_embed = M.LocalParticleDescriptors(M.DistanceSelection(3.0, max_included=10, box_lengths=[10.0, 10.0, 10.0]),
                                                M.ParticleEmbedding(20))
    decoder = Mo.MappingToDistribution(D.AutoregressiveBlockwise(6, [D.Normal] * 3 + [D.VonMises] * 3), name='decoder')
    backmap = Mo.BackmappingOnly(mask_and_embed, decoder)
    out = backmap([ref, all_coords, all_info])
    assert isinstance(out, P.Distribution)
    sample = out.sample()
    lp = out.log_prob(sample)
    assert sample.shape == (n_b, 6) and np.all(np.isfinite(lp.numpy()))
    target = rng.uniform(-3, 3, (n_b, 6)).astype(np.float32)
    backmap.compile(optimizer=Mo.Adam(learning_rate=1e-3), loss=Lo.LogProbLoss())
    before = backmap.evaluate([ref, all_coords, all_info], target, batch_size=20)
    hist = backmap.fit([ref, all_coords, all_info], target, batch_size=20, epochs=3)
    after = backmap.evaluate([ref, all_coords, all_info], target, batch_size=20)
    assert np.isfinite(hist['loss']).all() and after < before, (before, hist, after)
    pred = backmap.predict([ref, all_coords, all_info])
    assert pred.shape == (n_b, 6) and np.all(np.isfinite(pred))
