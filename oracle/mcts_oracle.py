"""CPU ORACLE for the search (TEST INFRASTRUCTURE): a struct-of-arrays restatement of the
reference's MCTS / PUCT (ai.py:299-659), small cases only (pure-Python loops).

It uses the C oracle for the env step and the legal placements, a pluggable evaluator with the
reference's `evaluate` contract (value: float, policy: float32 ndarray (27,39,11)) and a tape
RNG object, so that it can be pinned EXACTLY against the unmodified reference driven by the
same evaluator and the same tape (oracle/pin_mcts_against_reference.py; vectors in
tests/golden/mcts_golden.npz).  Numeric expressions are written with the same operand types
as the reference (numpy float32 scalars for non-root priors, Python floats elsewhere) so the
float semantics of this container's numpy (NEP 50) carry over.

Parity status: pinned against the reference (chosen move, pre- and post-prune visit counts,
save flag, node count) — see DESIGN.md.
"""
import math
from bisect import bisect

import numpy as np

from tetris_reinforcement_learning_b200.const import POLICY_SHAPE, PREVIEWS
from tetris_reinforcement_learning_b200.state import GAME_DTYPE

from . import oracle

PURPOSE_COIN, PURPOSE_CHOICE = 3, 4


class SearchTape:
    """Every random draw of one search, as Philox streams keyed by
    (seed, game_id, search_no)."""

    def __init__(self, seed, game_id, search_no):
        self.seed, self.game_id, self.search_no = int(seed), int(game_id), int(search_no)
        self.garbage_ctr = 0

    def coin(self):  # ai.py:324 random.random()
        return oracle.uniform(self.seed, self.game_id, self.search_no, PURPOSE_COIN)

    def choice_uniform(self):  # the single random() inside random.choices, ai.py:604
        return oracle.uniform(self.seed, self.game_id, self.search_no, PURPOSE_CHOICE)

    def gamma(self, alpha, n):  # ai.py:489 np.random.gamma(alpha, 1, n)
        return np.array([oracle.gamma(self.seed, self.game_id, self.search_no, i, alpha) for i in range(n)])

    @property
    def garbage_stream(self):
        return 1 + self.search_no


def truncate_previews(rec):
    """MCTS root: each queue cut to PREVIEWS pieces (ai.py:304-309)."""
    out = np.array(rec, dtype=GAME_DTYPE).reshape(1).copy()
    for pl in range(2):
        if int(out[0]["players"][pl]["qlen"]) > PREVIEWS:
            out[0]["players"][pl]["qlen"] = PREVIEWS
    return out


def is_terminal(rec):
    return bool(rec["players"][0]["game_over"]) or bool(rec["players"][1]["game_over"])


def winner(rec):  # game.py:217-225
    if rec["players"][0]["game_over"]:
        return 1
    if rec["players"][1]["game_over"]:
        return 0
    return -1


def no_move(rec):  # game.py:211-215
    p = rec["players"][int(rec["turn"])]
    return int(p["piece"]) == 255 and int(p["held"]) == 255


def legal_moves(rec):
    """Flat policy indices in np.argwhere order (ai.py:1016-1024)."""
    p = rec["players"][int(rec["turn"])]
    cur = int(p["piece"])
    alt = int(p["held"]) if int(p["held"]) != 255 else (int(p["queue"][0]) if int(p["qlen"]) > 0 else 255)
    mask = oracle.movegen_one(p["rows"], cur, alt)[0]
    return np.flatnonzero(mask.reshape(-1))


def search(cfg, game_rec, evaluate, tape):
    """One MCTS call.  game_rec: GAME_DTYPE array of shape (1,).  Returns a dict with the chosen
    move (flat index), root child moves, pre- and post-prune visits, priors, save flag, sizes."""
    vmin = -1 if cfg.use_tanh else 0
    vmid = 0 if cfg.use_tanh else 0.5
    vmax = 1
    negate = (lambda v: -v) if cfg.use_tanh else (lambda v: 1 - v)

    # ---- tree, struct of arrays; node 0 is the root ----
    parent = [-1]
    first_child = [-1]
    n_children = [0]
    move_of = [-1]
    prior = [0]
    visits = [0]
    value_sum = [0]
    value_avg = [0]
    state = [truncate_previews(game_rec)]   # materialised game per node (None until visited)
    max_depth = 0

    fast_iter = False
    if cfg.training and cfg.use_playout_cap_randomization:  # ai.py:323-330
        denom = cfg.playout_cap_chance * (cfg.playout_cap_mult - 1) + 1
        if tape.coin() < cfg.playout_cap_chance:
            max_iterations = math.ceil(cfg.playout_cap_mult * (cfg.MAX_ITER / denom))
        else:
            max_iterations = math.floor(cfg.MAX_ITER / denom)
            fast_iter = True
    else:
        max_iterations = cfg.MAX_ITER

    for _ in range(max_iterations):
        node, depth = 0, 0
        # ---- select (ai.py:346-393) ----
        while n_children[node] > 0:
            best_score, best = -1, None
            pv = visits[node]
            sqrt_parent = math.sqrt(pv)
            unvisited_scale = cfg.CPUCT * sqrt_parent / cfg.DPUCT
            check_forced = (cfg.use_forced_playouts_and_policy_target_pruning and cfg.training and node == 0
                            and not (cfg.use_playout_cap_randomization and fast_iter))
            for c in range(first_child[node], first_child[node] + n_children[node]):
                vc = visits[c]
                if vc == 0:
                    u = unvisited_scale * prior[c]
                else:
                    u = cfg.CPUCT * prior[c] * sqrt_parent / (cfg.DPUCT + vc)
                score = value_avg[c] + u
                if check_forced and vc >= 1:
                    if vc < math.sqrt(cfg.CForcedPlayout * prior[c] * pv):
                        score = float("inf")
                if score >= best_score:
                    best_score, best = score, c
            node = best
            depth += 1
        max_depth = max(max_depth, depth)
        leaf = node

        # ---- materialise (ai.py:398-403) ----
        if leaf != 0:
            g = state[parent[leaf]].copy()
            _, tape.garbage_ctr = oracle.env_step_rng(g, move_of[leaf], False, tape.seed, tape.garbage_stream,
                                                      tape.garbage_ctr)
            state[leaf] = g
        g = state[leaf]

        # ---- evaluate + expand, or terminal (ai.py:406-479) ----
        if not is_terminal(g[0]):
            value, policy = evaluate(g[0])
            policy = np.asarray(policy).reshape(POLICY_SHAPE)
            policy[policy <= 0] = 1e-25
            if not no_move(g[0]):
                moves = legal_moves(g[0])
                assert len(moves) > 0
                flat = policy.reshape(-1)
                policies = [flat[m] for m in moves]
                if leaf == 0 and cfg.use_root_softmax:  # ai.py:428-434
                    log_max = math.log(max(policies))
                    inv_temp = 1.0 / cfg.RootSoftmaxTemp
                    policies = [math.exp((math.log(p) - log_max) * inv_temp) for p in policies]
                policy_sum = sum(policies)
                first_child[leaf] = len(parent)
                n_children[leaf] = len(moves)
                if cfg.FpuStrategy == "absolute":
                    init_q = max(vmin, cfg.FpuValue)
                else:
                    init_q = max(vmin, negate(value))
                for p, m in zip(policies, moves):
                    parent.append(leaf); first_child.append(-1); n_children.append(0); move_of.append(int(m))
                    prior.append(p / policy_sum); visits.append(0); value_sum.append(0); value_avg.append(init_q)
                    state.append(None)
        else:
            w = winner(g[0])
            turn = int(g[0]["turn"])
            value = vmax if w == turn else (vmin if w == 1 - turn else vmid)

        # ---- root noise (ai.py:482-499) ----
        if cfg.training and not fast_iter and cfg.use_dirichlet_noise and leaf == 0:
            n = n_children[0]
            alpha = cfg.DIRICHLET_ALPHA
            if cfg.use_dirichlet_s:
                alpha *= cfg.DIRICHLET_S / n
            noise = tape.gamma(alpha, n)
            for i in range(n):
                c = first_child[0] + i
                prior[c] = prior[c] * (1 - cfg.DIRICHLET_EXPLORATION) + noise[i] * cfg.DIRICHLET_EXPLORATION

        # ---- backup (ai.py:511-533) ----
        value = negate(value)
        pos_value, neg_value = value, negate(value)
        leaf_turn = int(g[0]["turn"])
        n = leaf
        while True:
            visits[n] += 1
            value_sum[n] += pos_value if int(state[n][0]["turn"]) == leaf_turn else neg_value
            value_avg[n] = value_sum[n] / visits[n]
            if n == 0:
                break
            n = parent[n]

        # ---- sibling FPU refresh (ai.py:542-565) ----
        if leaf != 0 and cfg.FpuStrategy == "reduction":
            par = parent[leaf]
            if par != 0:
                explored = 0
                unvisited = []
                for c in range(first_child[par], first_child[par] + n_children[par]):
                    if visits[c] > 0:
                        explored += prior[c]
                    else:
                        unvisited.append(c)
                fpu = max(vmin, negate(value_avg[par]) - cfg.FpuValue * math.sqrt(explored))
                for c in unvisited:
                    value_avg[c] = fpu

    # ---- move choice on pre-prune visits (ai.py:571-614) ----
    kids = list(range(first_child[0], first_child[0] + n_children[0]))
    pre = [visits[c] for c in kids]
    max_n, max_i = 0, None
    for i, c in enumerate(kids):
        if visits[c] >= max_n:
            max_n, max_i = visits[c], i
    temp = cfg.temperature if cfg.training else 0
    counts = np.array(pre)
    if temp == 0:
        sel = int(np.argmax(counts))
    else:
        probs = counts ** (1 / temp)
        probs = probs / np.sum(probs)
        cum = _accumulate(probs)
        total = cum[-1] + 0.0
        sel = bisect(cum, tape.choice_uniform() * total, 0, len(cum) - 1)
    chosen = move_of[kids[sel]]

    # ---- policy-target pruning (ai.py:619-648) ----
    pruned = cfg.use_forced_playouts_and_policy_target_pruning and cfg.training and not fast_iter
    if pruned:
        b = kids[max_i]
        ref = value_avg[b] + cfg.CPUCT * prior[b] * math.sqrt(visits[0]) / (cfg.DPUCT + visits[b])
        for c in kids:
            if c == b or visits[c] <= 0:
                continue
            n_forced = math.sqrt(cfg.CForcedPlayout * prior[c] * visits[0])
            count = 0
            while True:
                if visits[c] == 1:
                    visits[c] = 0
                    break
                s = value_avg[c] + cfg.CPUCT * prior[c] * math.sqrt(visits[0]) / (cfg.DPUCT + visits[0])
                if count < n_forced and s < ref:
                    count += 1
                    visits[c] -= 1
                else:
                    break
    post = [visits[c] for c in kids]

    return {"move": int(chosen), "moves": [move_of[c] for c in kids], "visits_pre": pre, "visits_post": post,
            "priors": [float(prior[c]) for c in kids], "save": not fast_iter, "iterations": max_iterations,
            "n_nodes": len(parent), "max_depth": max_depth, "root_visits": visits[0],
            "root_value_avg": float(value_avg[0]), "garbage_draws": tape.garbage_ctr}


def _accumulate(weights):
    """itertools.accumulate as random.choices uses it (cumulative sums, left to right)."""
    out, tot = [], None
    for w in weights:
        tot = w if tot is None else tot + w
        out.append(tot)
    return out


def policy_target(result):
    """search_statistics (ai.py:1330-1361): round(n / total, 4) at each visited root child."""
    total = sum(result["visits_post"])
    assert total != 0
    return {m: round(n / total, 4) for m, n in zip(result["moves"], result["visits_post"]) if n != 0}
