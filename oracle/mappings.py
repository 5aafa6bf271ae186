"""DistanceSelection oracle -- NumPy restatement with TF's float32 op order, test infrastructure only.

Follows `vaemolsim/mappings.py:362-455` (`DistanceSelection.call`) line by line:
  :404      local = coords - ref
  :408-412  local -= box * tf.round(local / box)          (round half to even; per-call box wins over stored box)
  :417      ragged -> dense, pad value float32.max         (applied AFTER the wrap, so pads give d^2 = +inf)
  :420-426  pad the particle axis up to max_included
  :429      d^2 = reduce_sum(local * local, axis=-1)       (separate Mul then Sum => ((x^2+y^2)+z^2), no FMA)
  :433      top_k(-d^2, k)                                 (descending, ties -> lower index first, int32 indices)
  :436      gather;  :440-441 zero where d^2 > cutoff^2 (sq_cut cast to float32 by TF)
  :443-453  same gather + mask on particle_info (pad 0.0)
"""
import numpy as np

F32MAX = np.float32(np.finfo(np.float32).max)


def _to_rows(coords, row_splits=None, width=3):
    if row_splits is not None:
        return [np.asarray(coords[row_splits[i]:row_splits[i + 1]], np.float32) for i in range(len(row_splits) - 1)]
    if isinstance(coords, np.ndarray) and coords.ndim == 3:
        return None
    return [np.asarray(c, np.float32).reshape(-1, width) for c in coords]


def distance_selection(coords, ref, cutoff, max_included=50, box_lengths=None, particle_info=None, row_splits=None,
                       return_indices=False):
    """coords: dense [B, N, 3] float32, or ragged (list of [n_i, 3], or flat values + row_splits [B+1]).

    box_lengths: None, [3] (stored box) or [B, 3] (per-call box).  particle_info like coords with last dim P.
    Returns select_coords [B, k, 3] (and select_info [B, k, P]) (and indices [B, k] int32).
    """
    k = int(max_included)
    sq_cut = np.float32(cutoff**2)
    rows = _to_rows(coords, row_splits)
    if rows is None:
        dense = np.asarray(coords, np.float32)
        B, nmax = dense.shape[0], dense.shape[1]
        lens = np.full(B, nmax)
    else:
        B = len(rows)
        lens = np.array([r.shape[0] for r in rows])
        nmax = int(lens.max()) if B else 0
    ref = np.asarray(ref, np.float32).reshape(B, 1, 3)
    if box_lengths is not None:
        box = np.asarray(box_lengths, np.float32)
        box = box.reshape(B, 1, 3) if box.size == 3 * B and box.size != 3 else np.broadcast_to(
            box.reshape(1, 1, 3), (B, 1, 3))
    npad = max(nmax, k)
    local = np.full((B, npad, 3), F32MAX, np.float32)
    for b in range(B) if rows is not None else [None]:
        if rows is None:
            loc = dense - ref
            if box_lengths is not None:
                loc = loc - box * np.rint(loc / box)
            local[:, :nmax] = loc
        else:
            loc = rows[b] - ref[b]
            if box_lengths is not None:
                loc = loc - box[b] * np.rint(loc / box[b])
            local[b, :lens[b]] = loc
    sq = local * local
    d2 = (sq[..., 0] + sq[..., 1]) + sq[..., 2]
    order = np.argsort(d2, axis=1, kind='stable')[:, :k].astype(np.int32)
    near_d2 = np.take_along_axis(d2, order, axis=1)
    sel = np.take_along_axis(local, order[..., None], axis=1)
    mask = (near_d2 <= sq_cut)[..., None]
    sel = np.where(mask, sel, np.float32(0))
    outs = [sel]
    if particle_info is not None:
        if isinstance(particle_info, np.ndarray):
            pw = particle_info.shape[-1]
        else:
            pw = next((np.asarray(r).shape[-1] for r in particle_info if np.asarray(r).ndim == 2), 1)
        irows = _to_rows(particle_info, row_splits, pw)
        if irows is None:
            pi = np.asarray(particle_info, np.float32)
            P = pi.shape[-1]
            info = np.zeros((B, npad, P), np.float32)
            info[:, :pi.shape[1]] = pi
        else:
            P = pw
            info = np.zeros((B, npad, P), np.float32)
            for b in range(B):
                info[b, :lens[b]] = irows[b].reshape(-1, P)
        sinfo = np.take_along_axis(info, order[..., None], axis=1)
        outs.append(np.where(mask, sinfo, np.float32(0)))
    if return_indices:
        outs.append(order)
    return outs[0] if len(outs) == 1 else tuple(outs)
