"""Host-side index arithmetic for the mask-guided sparse window attention.

The reference materialises rolled / pooled K,V per window with torch.roll + window_partition + cat
(model/modules/sparse_transformer.py:182-221, ~40 copy kernels and 2x140 MB per layer).  Here the same
key set is described once by a small int32 table (token coordinates per window) that the CUDA kernel
gathers through; nothing is materialised.
"""
import functools
import math

import numpy as np


def rolled_valid_index(window):
    """Indices (into the 4*wh*ww concatenated rolled windows) of tokens that lie outside the own
    window -- the ``valid_ind_rolled`` buffer of sparse_transformer.py:140-153."""
    wh, ww = window
    eh, ew = (wh + 1) // 2, (ww + 1) // 2
    keep = []
    for top, left in ((True, True), (True, False), (False, True), (False, False)):
        m = np.ones((wh, ww), dtype=bool)
        rs = slice(0, wh - eh) if top else slice(eh, wh)
        cs = slice(0, ww - ew) if left else slice(ew, ww)
        m[rs, cs] = False
        keep.append(m.reshape(-1))
    return np.nonzero(np.concatenate(keep))[0].astype(np.int64)


@functools.lru_cache(maxsize=32)
def window_key_table(grid_h, grid_w, window=(5, 9)):
    """int32 [n_windows, wh*ww + n_rolled]: linear token index (y*grid_w + x) on the *padded* token
    grid of every non-pooled key a masked window attends to, in the reference's key order:
    own tokens (sparse_transformer.py:178), then the 4 rolled neighbourhoods (tl,tr,bl,br; :182-197)
    filtered by ``valid_ind_rolled`` (:199-200).  torch.roll wraps around the padded grid."""
    wh, ww = window
    assert grid_h % wh == 0 and grid_w % ww == 0
    eh, ew = (wh + 1) // 2, (ww + 1) // 2
    nwh, nww = grid_h // wh, grid_w // ww
    base = np.arange(grid_h * grid_w, dtype=np.int64).reshape(grid_h, grid_w)

    def part(a):                    # window_partition :104-115 on an index map
        return a.reshape(nwh, wh, nww, ww).transpose(0, 2, 1, 3).reshape(nwh * nww, wh * ww)

    own = part(base)
    rolled = [part(np.roll(base, (sy, sx), (0, 1)))
              for sy, sx in ((-eh, -ew), (-eh, ew), (eh, -ew), (eh, ew))]
    rolled = np.concatenate(rolled, 1)[:, rolled_valid_index(window)]
    return np.ascontiguousarray(np.concatenate([own, rolled], 1).astype(np.int32))


def padded_grid(h, w, window=(5, 9)):
    wh, ww = window
    return math.ceil(h / wh) * wh, math.ceil(w / ww) * ww


def token_grid(hw, ks=(7, 7), st=(3, 3), pd=(3, 3)):
    """Soft-split token grid of a feature map (sparse_transformer.py:20-23)."""
    return tuple((hw[i] + 2 * pd[i] - ks[i]) // st[i] + 1 for i in range(2))
