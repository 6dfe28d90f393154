"""Compact training sets: what the device leaves behind for a saved search, kept as arrays.

The reference stores a data set as one JSON list of 13-element samples, each with a nested
(27,39,11) policy list and four mirror-augmented copies (ai.py:1613-1666, 1822-1829): 162 KB of text
and 5 ms of Python per saved search, i.e. ~200 searches/s per host core against 8 x 10^4 searches/s
that one GPU produces.  A CompactSet keeps per saved search the packed position (400 B), the outcome
value and the visited root children (move, visit fraction) -- about 0.6 KB -- and expands any batch of
samples into exactly the tensors the JSON path yields (`batch_tensors`, mirror variants included), on
the training device.  The JSON format stays the default of make_training_set (drop-in); this is the
format that keeps up with the engine (`Config(data_format="compact")` or make_training_set(...,
data_format="compact")).
"""
import numpy as np
import torch

from .const import MINOS, POLICY_SHAPE, POLICY_SIZE, PREVIEWS, policy_index_to_piece, policy_piece_to_index
from .state import GAME_DTYPE

_PIECE_MIRROR = np.array([3, 5, 2, 0, 4, 1, 6])  # Z<->S, L<->J over "ZLOSIJT" (ai.py:1441-1466)
VARIANTS = 4                                     # (active reflected, other reflected) = (0,0) (0,1) (1,0) (1,1)


def _policy_mirror_index():
    """cell -> mirrored cell of reflect_policy (ai.py:1468-1516): piece / rotation swap and column flip
    new_col = 14 - col - size (+ adjustment for the folded Z/S/I rotations); -1 where it leaves the grid."""
    from .const import MATRIX_SIZE
    swap = {"Z": "S", "S": "Z", "L": "J", "J": "L"}
    perm = np.full(POLICY_SIZE, -1, dtype=np.int64)
    n_rows, n_cols = POLICY_SHAPE[1], POLICY_SHAPE[2]
    for plane in range(POLICY_SHAPE[0]):
        piece, rot, tsi = policy_index_to_piece[plane]
        new_piece = swap.get(piece, piece)
        new_rot = {1: 3, 3: 1}.get(rot, rot)
        adj = 0
        if new_piece in ("Z", "S", "I") and new_rot == 3:
            adj, new_rot = -1, 1
        new_plane = policy_piece_to_index[new_piece][new_rot][tsi]
        for col in range(n_cols):
            new_col = 10 - (col - 2) - MATRIX_SIZE[piece] + 2 + adj
            if 0 <= new_col < n_cols:
                for row in range(n_rows):
                    perm[(plane * n_rows + row) * n_cols + col] = (new_plane * n_rows + row) * n_cols + new_col
    return perm


POLICY_MIRROR = _policy_mirror_index()


def round4(x):
    """Vectorised Python round(x, 4) on float64 (the rounding of search_statistics, ai.py:1353): nearest multiple of
    1e-4 to the EXACT value of the double, ties to even.  np.rint(x * 1e4) agrees except where the product lands on a
    half although x itself is not one (3/160 = 0.01875 is stored slightly below the tie: Python gives 0.0187); those few
    elements go through Python's round."""
    x = np.asarray(x, dtype=np.float64)
    y = x * 1e4
    out = np.rint(y) / 1e4
    sus = np.flatnonzero(np.abs(y - np.floor(y) - 0.5) < 1e-6)
    if sus.size:
        out[sus] = [round(float(v), 4) for v in x[sus]]
    return out


class CompactSet:
    """state GAME_DTYPE[S], value f32[S], offsets i64[S+1], moves u16[T], frac f64[T] (visit fraction of a
    visited root child, rounded to 4 decimals exactly like search_statistics), augment bool."""

    def __init__(self, state, value, offsets, moves, frac, augment=True):
        self.state = np.ascontiguousarray(state, dtype=GAME_DTYPE)
        self.value = np.asarray(value, dtype=np.float32)
        self.offsets = np.asarray(offsets, dtype=np.int64)
        self.moves = np.asarray(moves, dtype=np.uint16)
        self.frac = np.asarray(frac, dtype=np.float64)
        self.augment = bool(augment)
        assert len(self.offsets) == len(self.state) + 1 == len(self.value) + 1 and self.offsets[-1] == len(self.moves) == len(self.frac)

    @property
    def n_searches(self):
        return len(self.state)

    def __len__(self):
        """Number of samples = what len() of the JSON sample list would be."""
        return self.n_searches * (VARIANTS if self.augment else 1)

    @classmethod
    def empty(cls, augment=True):
        return cls(np.zeros(0, GAME_DTYPE), [], [0], [], [], augment)

    @classmethod
    def from_searches(cls, searches, augment=True):
        """searches: iterable of (state record, moves, visits, value) in sample order."""
        states, values, offs, moves, frac = [], [], [0], [], []
        for st, mv, vis, val in searches:
            vis = np.asarray(vis, dtype=np.int64)
            total = int(vis.sum())
            assert total != 0
            nz = np.flatnonzero(vis)
            moves.append(np.asarray(mv, dtype=np.uint16)[nz])
            frac.append(np.array([round(int(n) / total, 4) for n in vis[nz]], dtype=np.float64))   # Python round: ai.py:1353
            offs.append(offs[-1] + len(nz))
            states.append(np.asarray(st, dtype=GAME_DTYPE).reshape(1))
            values.append(val)
        if not states:
            return cls.empty(augment)
        return cls(np.concatenate(states), values, offs, np.concatenate(moves), np.concatenate(frac), augment)

    @classmethod
    def from_records(cls, samples, ends, value_of, num_games=None, save_all=False, augment=True):
        """Vectorised form of generate_games' bookkeeping: `samples` (selfplay.SAMPLE_DTYPE) and `ends`
        (GAME_END_DTYPE) drained from the engine -> the saved searches of the first `num_games` finished games
        (by game id), per game player 0's searches in order then player 1's, with the outcome from the mover's
        point of view: value_of = (value_min, value_mid, value_max).  Same set, same order as from_searches
        over the reference-style loop."""
        if len(ends) == 0 or len(samples) == 0:
            return cls.empty(augment)
        ids, first = np.unique(ends["game_id"], return_index=True)
        if num_games is not None:
            ids, first = ids[:num_games], first[:num_games]
        winners = ends["winner"][first].astype(np.int64)
        keep = np.isin(samples["game_id"], ids) & ((samples["saved"] != 0) | bool(save_all))
        smp = samples[keep]
        smp = smp[np.lexsort((smp["search_no"], smp["turn"], smp["game_id"]))]
        w = winners[np.searchsorted(ids, smp["game_id"])]
        vmin, vmid, vmax = value_of
        value = np.where(w == -1, vmid, np.where(w == smp["turn"].astype(np.int64), vmax, vmin)).astype(np.float32)
        C = smp["n_children"].astype(np.int64)
        vis = smp["visits"].astype(np.int64)
        live = np.arange(vis.shape[1])[None, :] < C[:, None]
        vis = np.where(live, vis, 0)
        total = vis.sum(axis=1)
        assert (total > 0).all()
        nz = vis > 0
        cnt = nz.sum(axis=1)
        offsets = np.concatenate([[0], np.cumsum(cnt)])
        frac = round4(vis[nz] / np.repeat(total, cnt))
        return cls(smp["state"], value, offsets, smp["moves"][nz], frac, augment)

    @classmethod
    def concatenate(cls, sets):
        sets = [s for s in sets if s.n_searches]
        if not sets:
            return cls.empty()
        assert len({s.augment for s in sets}) == 1
        offs = [np.zeros(1, np.int64)]
        base = 0
        for s in sets:
            offs.append(s.offsets[1:] + base)
            base += int(s.offsets[-1])
        return cls(np.concatenate([s.state for s in sets]), np.concatenate([s.value for s in sets]), np.concatenate(offs),
                   np.concatenate([s.moves for s in sets]), np.concatenate([s.frac for s in sets]), sets[0].augment)

    def save(self, path):
        with open(path, "wb") as f:     # an open file: numpy must not append ".npz" to the reference-style name
            np.savez_compressed(f, state=self.state.view(np.uint8).reshape(len(self.state), -1), value=self.value,
                                offsets=self.offsets, moves=self.moves, frac=self.frac, augment=np.array(self.augment))

    @classmethod
    def load(cls, path):
        z = np.load(path)
        return cls(z["state"].copy().view(GAME_DTYPE).reshape(-1), z["value"], z["offsets"], z["moves"], z["frac"], bool(z["augment"]))

    def to_json_samples(self):
        """The reference's sample list (13-element lists, x4 mirror variants when augmented): what make_training_set
        writes with data_format="json" for the same games, in the same order."""
        from .ai import _policy_to_lists, features_from_state, reflect_grid, reflect_pieces, reflect_policy_array
        from .const import POLICY_SHAPE as shape
        out = []
        for i in range(self.n_searches):
            lo, hi = int(self.offsets[i]), int(self.offsets[i + 1])
            target = np.zeros(int(np.prod(shape)), dtype=np.float64)
            target[self.moves[lo:hi].astype(np.int64)] = self.frac[lo:hi]
            target = target.reshape(shape)
            feats = features_from_state(self.state[i])
            listify = lambda f: f.tolist() if isinstance(f, np.ndarray) else f  # noqa: E731
            value = float(self.value[i])
            value = int(value) if value == int(value) else value
            variants = [(0, 0)] if not self.augment else [(0, 0), (0, 1), (1, 0), (1, 1)]
            plain = mirrored = None
            for a_ref, o_ref in variants:
                d = [f.copy() if isinstance(f, np.ndarray) else f for f in feats]
                if a_ref:
                    d[0], d[1] = reflect_grid(d[0]), reflect_pieces(d[1])
                if o_ref:
                    d[5], d[6] = reflect_grid(d[5]), reflect_pieces(d[6])
                d = [listify(f) for f in d]
                if a_ref:
                    mirrored = mirrored or _policy_to_lists(reflect_policy_array(target))
                else:
                    plain = plain or _policy_to_lists(target)
                d.append(value)
                d.append(mirrored if a_ref else plain)
                out.append(d)
        return out

    # ---- expansion into the tensors the JSON path yields ------------------------------------------
    def batch_tensors(self, sample_idx, device="cpu"):
        """Samples `sample_idx` (indices into the virtual JSON list: search * 4 + variant when augmented) ->
        [a_grid (B,1,40,10) f32, a_pieces (B,7,7) f32, a_b2b, a_combo, a_garbage (B,) i64, o_grid, o_pieces,
        o_b2b, o_combo, o_garbage, color (B,) i64, value (B,) f32, policy (B,27,39,11) f32] = training._to_tensors
        of the corresponding JSON samples."""
        idx = np.asarray(sample_idx, dtype=np.int64)
        if self.augment:
            s, var = idx // VARIANTS, idx % VARIANTS
        else:
            s, var = idx, np.zeros_like(idx)
        refl = [(var >> 1) & 1, var & 1]                      # active reflected, other reflected
        st = self.state[s]
        B = len(idx)
        turn = st["turn"].astype(np.int64) & 1
        ar = np.arange(B)
        dev = torch.device(device)
        out = []
        for side in range(2):
            p = st["players"][ar, turn if side == 0 else 1 - turn]
            r = torch.from_numpy(refl[side].astype(np.bool_)).to(dev)
            rows = torch.from_numpy(p["rows"].astype(np.int32)).to(dev)                              # (B, 40)
            grid = ((rows[:, :, None] >> torch.arange(10, device=dev)) & 1).to(torch.float32)        # (B, 40, 10)
            grid = torch.where(r[:, None, None], grid.flip(-1), grid)
            ids = np.full((B, 2 + PREVIEWS), 255, dtype=np.int64)
            ids[:, 0], ids[:, 1] = p["piece"], p["held"]
            q = p["queue"][:, :PREVIEWS].astype(np.int64)
            ids[:, 2:] = np.where(np.arange(PREVIEWS)[None, :] < np.minimum(p["qlen"], PREVIEWS)[:, None], q, 255)
            ids = torch.from_numpy(ids).to(dev)
            mirror = torch.from_numpy(np.concatenate([_PIECE_MIRROR, np.full(249, 255)])).to(dev)
            ids = torch.where(r[:, None], mirror[ids], ids)
            table = (ids[:, :, None] == torch.arange(len(MINOS), device=dev)).to(torch.float32)       # (B, 7, 7)
            out += [grid[:, None], table,
                    torch.from_numpy(p["b2b"].astype(np.int64)).to(dev), torch.from_numpy(p["combo"].astype(np.int64)).to(dev),
                    torch.from_numpy(p["n_recv"].astype(np.int64)).to(dev)]
        out.append(torch.from_numpy(turn).to(dev))
        out.append(torch.from_numpy(self.value[s]).to(dev))
        # policy target: scatter the visit fractions of every sample's root children (mirrored with the active side)
        lo, hi = self.offsets[s], self.offsets[s + 1]
        cnt = hi - lo
        flat = np.repeat(lo - np.concatenate([[0], np.cumsum(cnt)[:-1]]), cnt) + np.arange(int(cnt.sum()))
        cells = self.moves[flat].astype(np.int64)
        b_of = np.repeat(ar, cnt)
        cells = np.where(refl[0][b_of] == 1, POLICY_MIRROR[cells], cells)
        assert (cells >= 0).all(), "a visited placement mirrors outside the policy grid"
        pol = torch.zeros((B, POLICY_SIZE), dtype=torch.float32, device=dev)
        pol[torch.from_numpy(b_of).to(dev), torch.from_numpy(cells).to(dev)] = torch.from_numpy(self.frac[flat]).to(torch.float32).to(dev)
        out.append(pol.view(B, *POLICY_SHAPE))
        return out
