"""Oracle: PUCT search (test infrastructure only -- see oracle/__init__.py).

Restates ``alphazero/mcts.py``.  Same arithmetic, same operation order, same
tie-breaks; unlike the reference it (a) keeps child statistics in per-node
numpy arrays instead of one Python object per child, (b) builds a child's
position only when the child is first visited instead of cloning a board for
every legal move at expansion time (``alphazero/mcts.py:58-64``; that clone is
~85 % of the reference's run time and changes no result), and (c) also returns
the raw root visit counts, which the reference computes (``mcts.py:144``) but
never exposes.
"""

from __future__ import annotations

import numpy as np

from . import gomoku


class _Node:
    """One expanded-or-not tree node.  alphazero/mcts.py:25-48: a node is a
    leaf until ``expand`` gives it children; N/W/Q start at 0/0.0/0.0."""

    __slots__ = ("state", "acts", "P", "N", "W", "Q", "kids", "n_self")

    def __init__(self, state):
        self.state = state
        self.acts = None      # list[(r,c)] in row-major order
        self.P = None         # float64[k]  (float32 priors widened, mcts.py:63)
        self.N = None         # int64[k]    child visit counts
        self.W = None         # float64[k]  child total values
        self.Q = None         # float64[k]  child mean values
        self.kids = None      # list[_Node|None]
        self.n_self = 0       # this node's own visit count (self.N in mcts.py:73)

    def expanded(self) -> bool:
        return self.acts is not None and len(self.acts) > 0

    def expand(self, policy: np.ndarray, acts) -> None:
        """alphazero/mcts.py:50-64 -- one child per legal action, in the given
        (row-major) order, prior = float(policy[r,c]); NOT renormalised."""
        k = len(acts)
        self.acts = list(acts)
        self.P = np.array([float(policy[r, c]) for r, c in acts], dtype=np.float64)
        self.N = np.zeros(k, dtype=np.int64)
        self.W = np.zeros(k, dtype=np.float64)
        self.Q = np.zeros(k, dtype=np.float64)
        self.kids = [None] * k

    def select(self, c_puct: float) -> int:
        """alphazero/mcts.py:66-74 -- index of the first child maximising
        ``Q + c_puct * P * sqrt(N_self + 1e-8) / (1 + N_child)`` in float64,
        evaluated left to right; ``max`` keeps the FIRST maximum, i.e. the
        lowest row-major cell."""
        root_term = np.sqrt(self.n_self + 1e-8)
        score = self.Q + ((c_puct * self.P) * root_term) / (1 + self.N)
        return int(np.argmax(score))


class SearchResult:
    __slots__ = ("pi", "action", "visits", "actions", "root_value", "stats")

    def __init__(self, pi, action, visits, actions, root_value, stats):
        self.pi = pi                # float32 (B,B)
        self.action = action        # (r,c) or None
        self.visits = visits        # int64 (B,B): root child visit counts scattered to cells
        self.actions = actions      # list[(r,c)] root children in order
        self.root_value = root_value
        self.stats = stats          # dict: evals, terminal_leaves, sum_depth, sum_scanned


def policy_from_visits(visit_counts: np.ndarray, temperature) -> np.ndarray:
    """alphazero/mcts.py:143-170 on a float32 vector of child visit counts:
    clip T at 1e-7, ``log(N + 1e-8) / T``, subtract max, exp, normalise;
    uniform if the sum is < 1e-8 or NaN.  Dtypes follow NumPy's promotion of
    whatever ``temperature`` is (np.float64 from the self-play schedule =>
    float64 intermediates under NumPy >= 2; a Python float => float32)."""
    if temperature <= 1e-7:
        temperature = 1e-7
    log_counts = np.log(visit_counts + 1e-8) / temperature
    log_counts -= np.max(log_counts)
    probs = np.exp(log_counts)
    total = probs.sum()
    if total < 1e-8 or np.isnan(total):
        probs = np.ones_like(probs) / len(probs)
    else:
        probs /= total
    return probs


class Search:
    """alphazero/mcts.py:86-97.  ``evaluator(state) -> (policy (B,B) float32 over
    ALL cells, value float from the side to move's view)``."""

    def __init__(self, evaluator, num_simulations, c_puct,
                 dirichlet_alpha=0.3, dirichlet_weight=0.25, rng=None):
        self.evaluator = evaluator
        self.num_simulations = num_simulations
        self.c_puct = c_puct
        self.dirichlet_alpha = dirichlet_alpha
        self.dirichlet_weight = dirichlet_weight
        self.rng = rng if rng is not None else np.random   # reference: global NumPy RNG

    def run(self, root_state: gomoku.State, temperature, add_root_noise=False,
            forced_action=None) -> SearchResult:
        n = root_state.size
        stats = {"evals": 0, "terminal_leaves": 0, "sum_depth": 0, "sum_scanned": 0}
        root = _Node(root_state.copy())

        # root evaluation; its value is discarded (mcts.py:109)
        policy, root_value = self.evaluator(root_state)
        stats["evals"] += 1
        policy = np.array(policy, dtype=np.float32, copy=True).reshape(n, n)
        acts = root_state.legal_actions()

        # Dirichlet noise on legal cells only, mixed in float32 storage (mcts.py:113-116)
        if add_root_noise:
            noise = self.rng.dirichlet([self.dirichlet_alpha] * len(acts))
            w = self.dirichlet_weight
            for (r, c), eta in zip(acts, noise):
                policy[r, c] = (1 - w) * policy[r, c] + w * eta

        root.expand(policy, acts)          # no backup: root.n_self stays 0 (mcts.py:120)

        for _ in range(self.num_simulations):          # mcts.py:123-141
            node = root
            path = []                                  # (parent, child index)
            while node.expanded() and not node.state.is_terminal():
                i = node.select(self.c_puct)
                stats["sum_scanned"] += len(node.acts)
                kid = node.kids[i]
                if kid is None:
                    kid = node.kids[i] = _Node(node.state.play(node.acts[i]))
                path.append((node, i))
                node = kid
            stats["sum_depth"] += len(path)

            result = node.state.result()
            if result:                                 # terminal leaf, mcts.py:132-134
                value = 0 if result == gomoku.DRAW else (1 if result == node.state.to_move else -1)
                stats["terminal_leaves"] += 1
            else:                                      # mcts.py:136-138
                leaf_policy, value = self.evaluator(node.state)
                stats["evals"] += 1
                node.expand(np.asarray(leaf_policy).reshape(n, n), node.state.legal_actions())

            # backup(-value) from the leaf to the root, sign alternating (mcts.py:76-82, 141)
            v = -value
            node.n_self += 1
            for parent, i in reversed(path):
                parent.N[i] += 1
                parent.W[i] += v
                parent.Q[i] = parent.W[i] / parent.N[i]
                v = -v
                parent.n_self += 1
            # invariant: a child's own count equals the edge count into it
            # (kid.n_self == parent.N[i]); both are kept because select() reads n_self.

        visits = np.zeros((n, n), dtype=np.int64)
        pi = np.zeros((n, n), dtype=np.float32)
        if not root.expanded():                        # mcts.py:152-153
            return SearchResult(pi, None, visits, [], root_value, stats)
        counts = root.N.astype(np.float32)             # mcts.py:144
        probs = policy_from_visits(counts, temperature)
        for (r, c), p, k in zip(root.acts, probs, root.N):
            pi[r, c] = p
            visits[r, c] = k
        if forced_action is not None:
            action = (int(forced_action[0]), int(forced_action[1]))
        else:                                          # mcts.py:177 -- always sampled
            action = root.acts[self.rng.choice(len(root.acts), p=probs)]
        return SearchResult(pi, action, visits, list(root.acts), root_value, stats)
