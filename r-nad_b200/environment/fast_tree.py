"""
Level-synchronous tree generator (SURVEY.md section 8 f#1).

The reference builds its stochastic matrix tree with a recursive Python DFS and one
third-party solver call per node (environment/tree.py:236-366; 10-12 ms per node), which
rules out the large BASELINE configurations (depth 6, A = 3, C = 3 is 14.9 million
nodes).  `Tree.generate_fast` builds the same seven tensors as batched tensor algebra on
the tree's device:

  top-down     per level: legal masks, Dirichlet(1/C) chance rows thresholded and
               renormalised like tree.py:182-197, child specifications, child ids by a
               prefix sum over the existing transitions, terminal payoffs;
  bottom-up    per level: child values gathered from the level below, expected payoffs
               `sum_k chance * value` (tree.py:280-282), every matrix game of the level
               solved at once by Shapley-Snow support enumeration (the algorithm of
               util/matrix_game.py, batched; smallest supports first, so pure equilibria
               win ties as in tree.py:227-231), root values `x^T M y`.

Node numbering is level order instead of the reference's DFS pre-order: node 0 is the
absorbing terminal, node 1 the root, children have larger ids than their parents and the
non-zero entries of `index_tensor` are a bijection onto [2, S) - everything
`assert_index_is_tree`, the kernels and the NashConv metric rely on.

The reference draws child shapes from arbitrary Python lambdas (one call per child).
Here the child specification is vectorised: by default children inherit the parent's
action counts and `depth_bound - 1` (a regular tree, like the reference's defaults);
`depth_jitter` reproduces main.py's `depth_bound - 1 - 2 * (random() < p)` thinning, and
`child_spec` accepts any callable on whole tensors.  Randomness comes from one
`torch.Generator` seeded by `seed`.

Host-side set-up, not part of the self-play hot path.
"""

from itertools import combinations

import torch


def _supports(a: int):
    """(I, J) support pairs of an a x a game with |I| == |J|, smallest first, lexicographic (as util.matrix_game._supports)."""
    out = []
    for k in range(1, a + 1):
        for rows in combinations(range(a), k):
            for cols in combinations(range(a), k):
                out.append((rows, cols))
    return out


def solve_zero_sum_batched(M: torch.Tensor, rows: torch.Tensor, cols: torch.Tensor, tol: float = 1e-9):
    """
    M (N, A, A) payoff matrices of the row (maximising) player, only [:rows[n], :cols[n]] meaningful.
    Returns x (N, A), y (N, A) float64 strategies (zero outside the legal actions), v (N,) game values and a
    bool mask of the games no support validated for (numerically degenerate; the caller falls back).
    Same algorithm and tie-break as util.matrix_game.solve_zero_sum, batched over N.
    """
    M = M.to(torch.float64)
    n, a, _ = M.shape
    dev = M.device
    ar = torch.arange(a, device=dev)
    legal_r = ar.view(1, a) < rows.view(n, 1)                      # (N, A)
    legal_c = ar.view(1, a) < cols.view(n, 1)
    both = legal_r.unsqueeze(2) & legal_c.unsqueeze(1)
    scale = torch.where(both, M.abs(), torch.zeros_like(M)).amax(dim=(1, 2)).clamp_min(1.0)
    eps = tol * scale
    x_out = torch.zeros((n, a), dtype=torch.float64, device=dev)
    y_out = torch.zeros((n, a), dtype=torch.float64, device=dev)
    v_out = torch.zeros(n, dtype=torch.float64, device=dev)
    todo = torch.ones(n, dtype=torch.bool, device=dev)
    big = torch.finfo(torch.float64).max

    for I, J in _supports(a):
        k = len(I)
        sel = torch.nonzero(todo & (rows > I[-1]) & (cols > J[-1])).flatten()
        if sel.numel() == 0:
            continue
        m = M[sel]
        lr, lc, e = legal_r[sel], legal_c[sel], eps[sel]
        if k == 1:
            i, j = I[0], J[0]
            v = m[:, i, j]
            col_max = torch.where(lr, m[:, :, j], torch.full_like(m[:, :, j], -big)).amax(1)
            row_min = torch.where(lc, m[:, i, :], torch.full_like(m[:, i, :], big)).amin(1)
            ok = (col_max <= v + e) & (row_min >= v - e)           # saddle point
            xs = torch.ones((sel.numel(), 1), dtype=torch.float64, device=dev)
            ys = xs
            vx = v
        else:
            Ii = torch.tensor(I, device=dev)
            Jj = torch.tensor(J, device=dev)
            sub = m[:, Ii][:, :, Jj]                                # (n_sel, k, k)
            lhs = torch.zeros((sel.numel(), k + 1, k + 1), dtype=torch.float64, device=dev)
            lhs[:, :k, k] = -1.0
            lhs[:, k, :k] = 1.0
            rhs = torch.zeros((sel.numel(), k + 1, 1), dtype=torch.float64, device=dev)
            rhs[:, k, 0] = 1.0
            lhs[:, :k, :k] = sub.transpose(1, 2)                    # sum_i x_i M[i, j] = v,  sum x = 1
            sol_x, info_x = torch.linalg.solve_ex(lhs, rhs)
            lhs[:, :k, :k] = sub                                    # sum_j M[i, j] y_j = v,  sum y = 1
            sol_y, info_y = torch.linalg.solve_ex(lhs, rhs)
            xs, vx = sol_x[:, :k, 0], sol_x[:, k, 0]
            ys, vy = sol_y[:, :k, 0], sol_y[:, k, 0]
            ok = (info_x == 0) & (info_y == 0) & torch.isfinite(sol_x[:, :, 0]).all(1) & torch.isfinite(sol_y[:, :, 0]).all(1)
            ok &= (xs.amin(1) >= -tol) & (ys.amin(1) >= -tol) & ((vx - vy).abs() <= e)
            xs = xs.clamp_min(0.0)
            ys = ys.clamp_min(0.0)
            xs = xs / xs.sum(1, keepdim=True)
            ys = ys / ys.sum(1, keepdim=True)
        xf = torch.zeros((sel.numel(), a), dtype=torch.float64, device=dev)
        yf = torch.zeros((sel.numel(), a), dtype=torch.float64, device=dev)
        xf[:, list(I)] = xs
        yf[:, list(J)] = ys
        if k > 1:
            # no profitable deviation outside the supports
            row_pay = torch.einsum("ni,nij->nj", xf, m)             # the column player minimises this
            col_pay = torch.einsum("nij,nj->ni", m, yf)             # the row player maximises this
            lo = torch.where(lc, row_pay, torch.full_like(row_pay, big)).amin(1)
            hi = torch.where(lr, col_pay, torch.full_like(col_pay, -big)).amax(1)
            ok &= (lo >= vx - 10 * e) & (hi <= vx + 10 * e)
        win = sel[ok]
        x_out[win] = xf[ok]
        y_out[win] = yf[ok]
        v_out[win] = vx[ok]
        todo[win] = False
        if not bool(todo.any()):
            break
    return x_out, y_out, v_out, todo


def default_child_spec(rows, cols, depth, gen):
    """Children inherit the parent's action counts; depth_bound - 1 (the reference's default lambdas, tree.py:141-151)."""
    return rows, cols, depth - 1


def depth_jitter(p: float):
    """main.py:31-39's thinning: depth_bound - 1 - 2 * (random() < p), action counts inherited."""

    def spec(rows, cols, depth, gen):
        drop = torch.rand(depth.shape, generator=gen, device=depth.device) < p
        return rows, cols, depth - 1 - 2 * drop.to(depth.dtype)

    return spec


def generate_fast(tree, seed: int = 0, child_spec=None, solve_chunk: int = 1 << 20, log=None):
    """
    Fills `tree`'s seven tensors (and `hash`) level by level on `tree.device`.  `child_spec(rows, cols, depth, gen)`
    maps the parents' (rows, cols, depth_bound) - one entry per existing transition - to the children's.
    """
    from util.matrix_game import solve_zero_sum

    dev = torch.device(tree.device)
    a, c = tree.max_actions, tree.max_transitions
    child_spec = child_spec or default_child_spec
    gen = torch.Generator(device=dev)
    gen.manual_seed(int(seed))
    terminal_values = torch.tensor(list(tree.terminal_values), dtype=torch.float32, device=dev)
    ar = torch.arange(a, device=dev)

    rows = torch.tensor([tree.row_actions], dtype=torch.int64, device=dev)
    cols = torch.tensor([tree.col_actions], dtype=torch.int64, device=dev)
    depth = torch.tensor([tree.depth_bound], dtype=torch.int64, device=dev)
    first_id = 1                                   # id of the first node of the current level (0 = absorbing node)
    levels = []
    while rows.numel() > 0:
        n = rows.numel()
        legal = ((ar.view(1, a, 1) < rows.view(n, 1, 1)) & (ar.view(1, 1, a) < cols.view(n, 1, 1)))   # (N, A, A)
        if c == 1:
            p = torch.ones((n, a, a, 1), dtype=torch.float32, device=dev)
        else:
            conc = torch.full((n, a, a, c), 1.0 / c, dtype=torch.float64, device=dev)
            gam = torch._standard_gamma(conc, generator=gen).clamp_min(1e-300)
            p = (gam / gam.sum(-1, keepdim=True)).to(torch.float32)          # Dirichlet(1/C, ..., 1/C), tree.py:187-190
            p = p - torch.where(p < tree.transition_threshold, p, torch.zeros_like(p))
            p = p / p.abs().sum(-1, keepdim=True).clamp_min(1e-12)            # F.normalize(p=1), tree.py:194-196
        p = p * legal.unsqueeze(-1)                                           # (N, A, A, C): (r, c, k) order
        exists = p > 0
        e_idx = torch.nonzero(exists.reshape(-1)).flatten()                   # existing transitions, (n, r, c, k) order
        parent = e_idx // (a * a * c)
        ch_rows, ch_cols, ch_depth = child_spec(rows[parent], cols[parent], depth[parent], gen)
        ch_rows = ch_rows.clamp(1, a)
        ch_cols = ch_cols.clamp(1, a)
        ch_depth = ch_depth.clamp_min(0)
        is_node = ch_depth > 0
        n_children = int(is_node.sum())
        next_first = first_id + n
        index = torch.zeros(n * a * a * c, dtype=torch.int64, device=dev)
        index[e_idx[is_node]] = next_first + torch.arange(n_children, device=dev)
        payoff = torch.zeros(n * a * a * c, dtype=torch.float32, device=dev)
        leaves = e_idx[~is_node]
        if leaves.numel():
            pick = torch.randint(0, terminal_values.numel(), (leaves.numel(),), generator=gen, device=dev)
            payoff[leaves] = terminal_values[pick]
        to_kraa = lambda t: t.view(n, a, a, c).permute(0, 3, 1, 2).contiguous()   # -> the reference's (S, C, A, A)
        levels.append({"first": first_id, "rows": rows, "cols": cols, "legal": legal.to(torch.float32),
                       "chance": to_kraa(p), "index": to_kraa(index), "payoff": to_kraa(payoff)})
        if log:
            log(f"level {len(levels) - 1}: {n} nodes, {n_children} children")
        rows, cols, depth = ch_rows[is_node], ch_cols[is_node], ch_depth[is_node]
        first_id = next_first

    size = first_id                                 # ids 0 .. size-1
    root_value = torch.zeros(size, dtype=torch.float32, device=dev)
    solution = torch.zeros((size, 2 * a), dtype=torch.float32, device=dev)
    for lvl in reversed(levels):
        n = lvl["rows"].numel()
        idx = lvl["index"]
        value = torch.where(idx > 0, root_value[idx], lvl["payoff"])          # tree.py:269-277
        ev = (value * lvl["chance"]).sum(1)                                   # (N, A, A), tree.py:280-282
        lvl["value"], lvl["ev"] = value, ev
        sl = slice(lvl["first"], lvl["first"] + n)
        for lo in range(0, n, solve_chunk):
            hi = min(n, lo + solve_chunk)
            x, y, _, failed = solve_zero_sum_batched(ev[lo:hi], lvl["rows"][lo:hi], lvl["cols"][lo:hi])
            for j in torch.nonzero(failed).flatten().tolist():                # numerically degenerate: host solver
                r_, c_ = int(lvl["rows"][lo + j]), int(lvl["cols"][lo + j])
                xs, ys, _ = solve_zero_sum(ev[lo + j, :r_, :c_].double().cpu().numpy())
                x[j].zero_()
                y[j].zero_()
                x[j, :r_] = torch.as_tensor(xs, device=dev)
                y[j, :c_] = torch.as_tensor(ys, device=dev)
            x32, y32 = x.to(torch.float32), y.to(torch.float32)
            solution[lvl["first"] + lo: lvl["first"] + hi, :a] = x32
            solution[lvl["first"] + lo: lvl["first"] + hi, a:] = y32
            # x^T M y in fp32, the two products of tree.py:306-308
            root_value[lvl["first"] + lo: lvl["first"] + hi] = torch.einsum(
                "nj,nj->n", torch.einsum("ni,nij->nj", x32, ev[lo:hi]), y32)
        del sl

    def with_absorbing(key, shape_tail, dtype):
        head = torch.zeros((1,) + shape_tail, dtype=dtype, device=dev)
        return torch.cat([head] + [lvl[key].reshape((-1,) + shape_tail) for lvl in levels])

    tree.index_tensor = with_absorbing("index", (c, a, a), torch.int64)
    tree.value_tensor = with_absorbing("value", (c, a, a), torch.float32)
    tree.chance_tensor = with_absorbing("chance", (c, a, a), torch.float32)
    tree.expected_value_tensor = with_absorbing("ev", (1, a, a), torch.float32)
    tree.legal_tensor = with_absorbing("legal", (1, a, a), torch.float32)
    tree.chance_tensor[0, 0, 0, 0] = 1.0            # the absorbing node: one legal cell, one certain self-transition
    tree.legal_tensor[0, 0, 0, 0] = 1.0             # (tree.py:338-349)
    tree.root_value_tensor = root_value.view(size, 1)
    tree.solution_tensor = solution
    tree.hash = int(torch.randint(-(2 ** 62), 2 ** 62, (1,), generator=torch.Generator().manual_seed(int(seed))).item())
    tree._packed = None
    tree._depth_hint = len(levels)
    return tree
