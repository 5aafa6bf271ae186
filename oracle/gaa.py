"""Geometric-algebra attention oracle -- NumPy restatement, test infrastructure only.

PARITY UNPINNED.  The arithmetic of `AttentionBlock` / `ParticleEmbedding` (`vaemolsim/mappings.py:480-688`) lives in the
third-party package `geometric-algebra-attention` (github.com/klarh/geometric_algebra_attention; `pyproject.toml:28`,
version NOT pinned, not vendored, not installable here) and in Keras (`Dense`, `LayerNormalization`, `Masking`).  This
file restates the PUBLISHED algorithm of `geometric_algebra_attention.base.VectorAttention` /
`.tensorflow.geometric_algebra` for exactly the configuration the reference constructs (`mappings.py:518-525,633-647`:
`rank=2, merge_fun='concat', join_fun='concat'`, default `invariant_mode='single'`), from the package's documentation
and source as published [recalled, no copy at hand]:

  * tuples (pairs) are laid out by `_get_broadcast_indices`: tuple position 0 varies along the LAST particle axis, position
    1 along the one before it.  Entry [b, i, j] of every pair tensor therefore belongs to the product r_j * r_i;
  * `vecvec(a, b)` = [a.b, a_x b_y - a_y b_x, a_x b_z - a_z b_x, a_y b_z - a_z b_y] (scalar + bivector);
    `vecvec_invariants(p)` = [p_0, |p_1..3|] (`custom_norm`: plain Euclidean norm, its custom gradient only guards 0/0);
  * `invar_values = value_net(invariants)`; `merged = v_j @ merge_kernel_0 + v_i @ merge_kernel_1` (`_merge_fun`,
    'concat'); `joined = invar_values @ join_kernel_1 + merged @ join_kernel_2` (`_join_fun`, 'concat');
    `new_values = joined` (plain VectorAttention: product weights 1); `scores = score_net(joined)`;
  * a position mask m [B, n] masks pair (i, j) unless m_i and m_j: masked scores are replaced by -1e9;
  * `reduce=False`: softmax over j for every i, output_i = sum_j attention_ij new_values_ij  -> [B, n, D];
    `reduce=True`: ONE softmax over all (i, j) of a cloud, output = sum_ij attention_ij new_values_ij -> [B, D].

Keras pieces: `Dense` (x @ kernel + bias), `LayerNormalization()` (last axis, epsilon 1e-3, gamma / beta, biased
variance), `Masking()` (mask_value 0: a particle is masked when ALL its coordinates are 0; values pass unchanged).

The reference-side structure follows `mappings.py`:
  :503-535  AttentionBlock.build: score_fun = Dense(H, act) Dense(1); value_fun = Dense(H) LN act Dense(D);
            nonlinearity = Dense(H) LN act Dense(D)
  :553-558  AttentionBlock.call: attn([coords, emb]) -> nonlinearity -> + emb
  :618-647  ParticleEmbedding.build: info_net = Dense(E); num_blocks AttentionBlocks; Masking; final VectorAttention
            (reduce=True)
  :666-680  ParticleEmbedding.call: mask coords; emb = info_net(info); blocks; final attention
"""
import numpy as np

LN_EPS = 1e-3      # tf.keras.layers.LayerNormalization default epsilon
MASKED_SCORE = -1e9


def _act(name):
    if name in (None, 'linear'):
        return lambda x: x
    if name == 'relu':
        return lambda x: np.maximum(x, 0)
    if name == 'tanh':
        return np.tanh
    raise ValueError(name)


def layer_norm(x, gamma, beta, eps=LN_EPS):
    mean = x.mean(axis=-1, keepdims=True)
    var = ((x - mean)**2).mean(axis=-1, keepdims=True)
    return (x - mean) / np.sqrt(var + x.dtype.type(eps)) * gamma + beta


def pair_invariants(r):
    """r [B, n, 3] -> [B, n, n, 2]: entry [b, i, j] = (r_j . r_i, |r_j ^ r_i|)."""
    a = r[:, None, :, :]  # tuple position 0: last particle axis (j)
    b = r[:, :, None, :]  # tuple position 1: axis i
    dot = a[..., 0] * b[..., 0] + a[..., 1] * b[..., 1] + a[..., 2] * b[..., 2]
    xy = a[..., 0] * b[..., 1] - a[..., 1] * b[..., 0]
    xz = a[..., 0] * b[..., 2] - a[..., 2] * b[..., 0]
    yz = a[..., 1] * b[..., 2] - a[..., 2] * b[..., 1]
    return np.stack([dot, np.sqrt(xy * xy + xz * xz + yz * yz)], axis=-1)


def keras_mask(coords):
    """tf.keras.layers.Masking(mask_value=0.0).compute_mask: any(coords != 0, axis=-1)."""
    return np.any(coords != 0, axis=-1)


def init_mlp(rng, n_in, hidden, n_out, layer_norm_=True, dtype=np.float32):
    def glorot(a, b):
        lim = np.sqrt(6.0 / (a + b))
        return rng.uniform(-lim, lim, size=(a, b)).astype(dtype)

    w = [glorot(n_in, hidden), rng.normal(0, 0.1, hidden).astype(dtype)]
    if layer_norm_:
        w += [(1 + 0.2 * rng.normal(size=hidden)).astype(dtype), rng.normal(0, 0.1, hidden).astype(dtype)]
    w += [glorot(hidden, n_out), rng.normal(0, 0.1, n_out).astype(dtype)]
    return w


def init_attention(rng, D, hidden, dtype=np.float32):
    """Weights of one VectorAttention(rank 2, concat / concat) with the reference's score / value networks."""
    sd = np.sqrt(2.0 / 2 / D)
    return {'merge': [rng.normal(0, sd, (D, D)).astype(dtype) for _ in range(2)],
            'join': [rng.normal(0, sd, (D, D)).astype(dtype) for _ in range(2)],
            'score': init_mlp(rng, D, hidden, 1, layer_norm_=False, dtype=dtype),
            'value': init_mlp(rng, 2, hidden, D, dtype=dtype)}


def init_block(rng, D, hidden, dtype=np.float32):
    w = init_attention(rng, D, hidden, dtype)
    w['nonlin'] = init_mlp(rng, D, hidden, D, dtype=dtype)
    return w


def init_embedding(rng, P, E, hidden=40, num_blocks=2, dtype=np.float32):
    lim = np.sqrt(6.0 / (P + E))
    return {'info': [rng.uniform(-lim, lim, (P, E)).astype(dtype), rng.normal(0, 0.1, E).astype(dtype)],
            'blocks': [init_block(rng, E, hidden, dtype) for _ in range(num_blocks)],
            'final': init_attention(rng, E, hidden, dtype)}


def cast(w, dtype):
    if isinstance(w, dict):
        return {k: cast(v, dtype) for k, v in w.items()}
    if isinstance(w, list):
        return [cast(v, dtype) for v in w]
    return np.asarray(w, dtype)


def mlp_ln(x, w, activation):
    """Dense(H) -> LayerNormalization -> Activation -> Dense(out)   (mappings.py:509-514, 526-531, 638-643)."""
    W1, b1, g, be, W2, b2 = w
    return _act(activation)(layer_norm(x @ W1 + b1, g, be)) @ W2 + b2


def vector_attention(r, v, w, reduce, activation='relu', mask=None, return_attention=False):
    """r [B, n, 3], v [B, n, D], mask [B, n] bool or None -> [B, n, D] (reduce False) or [B, D] (reduce True)."""
    B, n, D = v.shape
    inv = pair_invariants(r)                                  # [B, n, n, 2]
    iv = mlp_ln(inv, w['value'], activation)                  # [B, n, n, D]
    merged = (v @ w['merge'][0])[:, None, :, :] + (v @ w['merge'][1])[:, :, None, :]
    joined = iv @ w['join'][0] + merged @ w['join'][1]
    W1, b1, W2, b2 = w['score']
    scores = (_act(activation)(joined @ W1 + b1) @ W2 + b2)[..., 0]   # [B, n, n]
    if mask is not None:
        pm = mask[:, None, :] & mask[:, :, None]
        scores = np.where(pm, scores, scores.dtype.type(MASKED_SCORE))
    if reduce:
        s = scores.reshape(B, n * n)
        e = np.exp(s - s.max(axis=-1, keepdims=True))
        att = (e / e.sum(axis=-1, keepdims=True)).reshape(B, n, n)
        out = (att[..., None] * joined).sum(axis=(1, 2))
    else:
        e = np.exp(scores - scores.max(axis=-1, keepdims=True))
        att = e / e.sum(axis=-1, keepdims=True)
        out = (att[..., None] * joined).sum(axis=2)
    return (out, att) if return_attention else out


def attention_block(r, v, w, activation='relu', mask=None):
    """mappings.py:553-558."""
    new = vector_attention(r, v, w, False, activation, mask)
    return mlp_ln(new, w['nonlin'], activation) + v


def particle_embedding(r, info, w, activation='relu', mask_zero=True):
    """mappings.py:666-680."""
    mask = keras_mask(r) if mask_zero else None
    emb = info @ w['info'][0] + w['info'][1]
    for blk in w['blocks']:
        emb = attention_block(r, emb, blk, activation, mask)
    return vector_attention(r, emb, w['final'], True, activation, mask)
