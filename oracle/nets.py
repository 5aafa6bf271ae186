"""Dense / FCDeepNN / MADE oracle -- NumPy restatement, test infrastructure only.

Follows
  * `vaemolsim/mappings.py:90-123` (`FCDeepNN.build`: Dense(hidden, activation)... Dense(prod(target)) Reshape)
  * `vaemolsim/mappings.py:125-155` (`FCDeepNN.call`: flatten, cos/sin featurisation of periodic dofs `:144-149`)
  * `vaemolsim/flows.py:450-487`    (`MaskedSplineBijector.build`: three `tfp.bijectors.AutoregressiveNetwork`,
                                     hidden_units=[hidden_dim], tanh, optional conditional input)
  * `vaemolsim/dists.py:301-305`    (`AutoregressiveBlockwise.build`: AutoregressiveNetwork(max(param_nums), num_dofs))
and the published tensorflow-probability v0.23.0 `bijectors/masked_autoregressive.py`
(`_create_input_order`, `_create_degrees` with hidden_degrees='equal', `_create_masks`,
`_make_dense_autoregressive_masks`, `AutoregressiveNetwork.build/call`; third-party, not vendored).
Architecture pin: the notebook `summary()` parameter counts hold only when the
conditional input enters bias-free into EVERY layer -- checked in tests/test_oracle_nets.py.
"""
import numpy as np


# ----------------------------------------------------------------------------- initialisers
def glorot_uniform(rng, fan_in, fan_out, dtype=np.float32):
    """Keras 'glorot_uniform' (FCDeepNN default, mappings.py:50): U(-l, l), l = sqrt(6/(fan_in+fan_out))."""
    lim = np.sqrt(6.0 / (fan_in + fan_out))
    return rng.uniform(-lim, lim, size=(fan_in, fan_out)).astype(dtype)


def truncated_normal(rng, fan_in, fan_out, std=0.05, dtype=np.float32):
    """Keras 'truncated_normal' string initialiser (flows.py:109): N(0, 0.05) resampled beyond 2 sigma."""
    out = rng.normal(0.0, std, size=(fan_in, fan_out))
    bad = np.abs(out) > 2 * std
    while bad.any():
        out[bad] = rng.normal(0.0, std, size=int(bad.sum()))
        bad = np.abs(out) > 2 * std
    return out.astype(dtype)


# ----------------------------------------------------------------------------- dense
def dense(x, W, b=None, act=None):
    y = x @ W
    if b is not None:
        y = y + b
    if act == 'relu':
        y = np.maximum(y, 0)
    elif act == 'tanh':
        y = np.tanh(y)
    elif act is not None:
        raise ValueError(act)
    return y.astype(x.dtype)


def periodic_featurise(x, periodic_mask):
    """mappings.py:144-149: concat([x_nonperiodic, cos(x_p), sin(x_p)])."""
    pm = np.asarray(periodic_mask, dtype=bool)
    if not pm.any():
        return x
    return np.concatenate([x[:, ~pm], np.cos(x[:, pm]), np.sin(x[:, pm])], axis=-1).astype(x.dtype)


def fcdeepnn_init(rng, din, hidden, target_shape, periodic_mask=None, dtype=np.float32):
    """Weights of FCDeepNN: list of (W, b).  din is the flattened input size before featurisation."""
    n_p = int(np.sum(periodic_mask)) if periodic_mask is not None else 0
    sizes = [din + n_p] + list(hidden) + [int(np.prod(target_shape))]
    return [(glorot_uniform(rng, a, b, dtype), np.zeros(b, dtype)) for a, b in zip(sizes[:-1], sizes[1:])]


def fcdeepnn_forward(x, layers, target_shape, periodic_mask=None, activation='relu', return_hidden=False):
    """mappings.py:141-155 (batch_norm=False)."""
    out = x.reshape(x.shape[0], -1)
    if periodic_mask is not None:
        out = periodic_featurise(out, periodic_mask)
    hid = [out]
    for W, b in layers[:-1]:
        out = dense(out, W, b, activation)
        hid.append(out)
    out = dense(out, layers[-1][0], layers[-1][1], None)
    out = out.reshape((x.shape[0],) + tuple(target_shape))
    return (out, hid) if return_hidden else out


# ----------------------------------------------------------------------------- MADE
def create_input_order(event_size, input_order='left-to-right'):
    if isinstance(input_order, str):
        if input_order == 'left-to-right':
            return np.arange(1, event_size + 1)
        if input_order == 'right-to-left':
            return np.arange(event_size, 0, -1)
        raise ValueError("input_order %r not supported by the oracle" % input_order)
    order = np.array(input_order)
    if not np.all(np.sort(order) == np.arange(1, event_size + 1)):
        raise ValueError('Invalid input order')
    return order


def create_degrees(event_size, hidden_units, input_order='left-to-right'):
    """TFP `_create_degrees`, hidden_degrees='equal'."""
    degrees = [create_input_order(event_size, input_order)]
    for units in hidden_units:
        min_degree = min(int(np.min(degrees[-1])), event_size - 1)
        degrees.append(
            np.maximum(min_degree,
                       np.ceil(np.arange(1, units + 1) * (event_size - 1) / float(units + 1)).astype(np.int32)))
    return degrees


def made_masks(params, event_size, hidden_units, input_order='left-to-right'):
    """TFP `_make_dense_autoregressive_masks`: list of boolean [in, out] masks; last one tiled `params` wide."""
    deg = create_degrees(event_size, hidden_units, input_order)
    masks = [inp[:, None] <= out[None, :] for inp, out in zip(deg[:-1], deg[1:])]
    last = deg[-1][:, None] < deg[0][None, :]
    last = np.reshape(np.tile(last[..., None], [1, 1, params]), [last.shape[0], event_size * params])
    return masks + [last]


def made_init(rng, params, event_size, hidden_units, input_order='left-to-right', cond_size=0,
              kernel_init='glorot_uniform', dtype=np.float32):
    """Weights of an AutoregressiveNetwork: list of dict(W (pre-masked), b, Wc or None, mask)."""
    masks = made_masks(params, event_size, hidden_units, input_order)
    sizes = [event_size] + list(hidden_units) + [event_size * params]
    init = glorot_uniform if kernel_init == 'glorot_uniform' else truncated_normal
    layers = []
    for k, (a, b) in enumerate(zip(sizes[:-1], sizes[1:])):
        W = init(rng, a, b, dtype=dtype) * masks[k].astype(dtype)  # masked initializer + constraint
        Wc = init(rng, cond_size, b, dtype=dtype) if cond_size else None  # bias-free, every layer
        layers.append(dict(W=W, b=np.zeros(b, dtype), Wc=Wc, mask=masks[k]))
    return layers


def made_forward(x, layers, params, cond=None, activation='tanh'):
    """AutoregressiveNetwork.call: x [B, D] (cond [B, C]) -> [B, D, params]."""
    out = x
    for k, L in enumerate(layers):
        y = out @ L['W'] + L['b']
        if L['Wc'] is not None:
            if cond is None:
                raise ValueError('`conditional_input` must be passed as a named argument.')
            y = y + cond @ L['Wc']
        if k + 1 < len(layers):
            y = np.tanh(y) if activation == 'tanh' else (np.maximum(y, 0) if activation == 'relu' else y)
        out = y.astype(x.dtype)
    return out.reshape(x.shape[0], x.shape[1], params)


def made_param_count(layers):
    n = 0
    for L in layers:
        n += L['W'].size + L['b'].size + (L['Wc'].size if L['Wc'] is not None else 0)
    return n


def maf_block_orders(num_blocks, data_dim, order_seed):
    """flows.py:606-621: first 'right-to-left', last 'left-to-right', middle = seeded shuffles of 1..D."""
    rng = np.random.default_rng(order_seed)
    orders = []
    for i in range(num_blocks):
        if i == 0:
            orders.append('right-to-left')
        elif i == num_blocks - 1:
            orders.append('left-to-right')
        else:
            o = np.arange(start=1, stop=data_dim + 1)
            rng.shuffle(o)
            orders.append(o)
    return orders


# ------------------------------------------------------------------------------------------------ batch normalisation
def batch_norm_moments(x):
    """tf.nn.moments(x, axes=[0]): mean, then the mean of squared differences (biased variance).  [TF-recalled]"""
    mean = x.mean(axis=0)
    return mean, ((x - mean) ** 2).mean(axis=0)


def batch_norm_normalize(x, mean, var, gamma, beta, eps=1e-3):
    """tf.nn.batch_normalization as used by tf.keras.layers.BatchNormalization (mappings.py:113-114) and by the INVERSE
    of tfp.bijectors.BatchNormalization (flows.py:308-309): x * inv + (beta - mean * inv), inv = rsqrt(var + eps) * gamma;
    log-det per row = sum(log gamma) - 0.5 sum(log(var + eps)).  [TF/TFP-recalled]"""
    inv = gamma / np.sqrt(var + eps)
    return x * inv + (beta - mean * inv), float(np.sum(np.log(gamma)) - 0.5 * np.sum(np.log(var + eps)))


def batch_norm_denormalize(y, mean, var, gamma, beta, eps=1e-3):
    """Forward of tfp.bijectors.BatchNormalization: y * r + (mean - beta * r), r = sqrt(var + eps) / gamma.  [TFP-recalled]"""
    r = np.sqrt(var + eps) / gamma
    return y * r + (mean - beta * r), float(-(np.sum(np.log(gamma)) - 0.5 * np.sum(np.log(var + eps))))
