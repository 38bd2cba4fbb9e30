"""
Stochastic matrix-tree game description (host side).

API mirror of the reference `environment/tree.py` (class `Tree`,
tree.py:66-442): same constructor keywords, same seven per-node tensors with
the same shapes / dtypes / DFS pre-order node numbering (node 0 = absorbing
terminal, node 1 = root), same `generate / assert_index_is_tree / save / load
/ to` methods, same `saved_keys`.

What is different underneath:

* the tree is built with an explicit pre-order node list in numpy instead of
  recursive per-node `torch.cat`s, and the matrix games are solved with
  `util.matrix_game.solve_zero_sum` because the reference's solver
  (pygambit) is a third-party package that is absent here;
* the *order and kind of RNG draws* (numpy Dirichlet per constructed node,
  python `random.choice` per terminal payoff, user lambdas evaluated in
  row / col / depth order) follows reference tree.py:164-197,253-277, so the
  same `np.random.seed / random.seed / torch.manual_seed` give the same tree
  (index / chance / legal bit-exact; values up to solver rounding);
* `packed()` hands the kernels a compact device table (see
  `_b200.PackedTree`, DESIGN.md "Data layout").

Tree construction is one-off set-up, not part of the self-play hot path.
"""

import logging
import os
import random
import time
from typing import Dict

import numpy as np
import torch

from util.matrix_game import solve_zero_sum


class _NodeSpec:
    """What the user lambdas see for a node (reference passes the sub-`Tree`)."""

    __slots__ = ("is_root", "device", "max_actions", "max_transitions", "row_actions",
                 "col_actions", "depth_bound", "transition_threshold", "terminal_values")

    def __init__(self, tree, row_actions, col_actions, depth_bound, is_root=False):
        self.is_root = is_root
        self.device = tree.device
        self.max_actions = tree.max_actions
        self.max_transitions = tree.max_transitions
        self.row_actions = row_actions
        self.col_actions = col_actions
        self.depth_bound = depth_bound
        self.transition_threshold = tree.transition_threshold
        self.terminal_values = tree.terminal_values


def _draw_chance(max_actions, max_transitions, threshold):
    """
    One node's chance tensor (C, A, A) float32, before legality masking.
    Same arithmetic as reference tree.py:182-197: Dirichlet(1/C) in float64,
    cast to float32, entries below the threshold zeroed, L1-renormalised.
    """
    c = max_transitions
    raw = np.random.dirichlet((1 / c,) * c, (1, max_actions, max_actions))
    p = raw[0].astype(np.float32)                           # (A, A, C)
    p = p - np.where(p < threshold, p, np.float32(0))
    norm = np.abs(p[..., 0])
    for k in range(1, c):
        norm = norm + np.abs(p[..., k])
    p = p / np.maximum(norm, np.float32(1e-12))[..., None]
    return np.ascontiguousarray(np.moveaxis(p, 2, 0))       # (C, A, A)


class Tree:
    def __init__(
        self,
        is_root=True,
        device=torch.device("cpu"),
        max_actions=3,
        max_transitions=1,
        row_actions=None,
        col_actions=None,
        depth_bound=1,
        row_actions_lambda=None,
        col_actions_lambda=None,
        depth_bound_lambda=None,
        transition_threshold=0,
        terminal_values=(-1, 1),
        desc="",
    ):
        if max_transitions * transition_threshold > 1:
            # reference tree.py:194-195 would leave a legal (r, c) without any
            # transition, which later breaks the chance sampling.
            raise ValueError("transition_threshold must not exceed 1 / max_transitions")
        self.is_root = is_root
        self.device = device
        self.max_actions = max_actions
        self.max_transitions = max_transitions
        self.row_actions = row_actions if row_actions is not None else max_actions
        self.col_actions = col_actions if col_actions is not None else max_actions
        self.depth_bound = depth_bound
        self.transition_threshold = transition_threshold
        self.terminal_values = terminal_values

        a, c = max_actions, max_transitions
        legal = np.zeros((a, a), dtype=np.float32)
        legal[: self.row_actions, : self.col_actions] = 1.0
        chance = _draw_chance(a, c, transition_threshold) * legal

        self.index_tensor = torch.zeros((1, c, a, a), device=device, dtype=torch.long)
        self.value_tensor = torch.zeros((1, c, a, a), device=device, dtype=torch.float)
        self.expected_value_tensor = torch.zeros((1, 1, a, a), device=device, dtype=torch.float)
        self.legal_tensor = torch.from_numpy(legal).reshape(1, 1, a, a).to(device)
        self.chance_tensor = torch.from_numpy(chance).reshape(1, c, a, a).to(device)
        self.root_value_tensor = torch.zeros((1, 1), device=device, dtype=torch.float)
        self.solution_tensor = torch.zeros((1, 2 * a), device=device, dtype=torch.float)
        self.desc = desc
        self.hash = 0

        self.saved_keys = list(self.__dict__.keys())
        # only the attributes above are written by save() and restored by load()

        self.row_actions_lambda = row_actions_lambda or (lambda node: node.row_actions)
        self.col_actions_lambda = col_actions_lambda or (lambda node: node.col_actions)
        self.depth_bound_lambda = depth_bound_lambda or (lambda node: node.depth_bound - 1)
        self._packed = None
        self._depth_hint = None      # number of levels, when the generator knows it

    # ------------------------------------------------------------------ build

    def _child_spec(self, parent: _NodeSpec) -> _NodeSpec:
        rows = min(self.max_actions, max(1, self.row_actions_lambda(parent)))
        cols = min(self.max_actions, max(1, self.col_actions_lambda(parent)))
        depth = max(0, self.depth_bound_lambda(parent))
        return _NodeSpec(self, rows, cols, depth)

    def _solve(self, M: np.ndarray, max_actions=None):
        """
        M: (rows, cols) float32 expected-payoff matrix.  Returns the joint
        strategy (2 * max_actions,) float32: row strategy then column strategy,
        zero-padded (layout of reference tree.py:199-234).
        """
        a = self.max_actions if max_actions is None else max_actions
        rows, cols = M.shape
        x, y, _ = solve_zero_sum(M)
        joint = np.zeros(2 * a, dtype=np.float32)
        joint[:rows] = x
        joint[a: a + cols] = y
        return joint

    def _build(self, spec: _NodeSpec, chance: np.ndarray, nodes: list) -> int:
        """Appends `spec`'s subtree to `nodes` in pre-order; returns its slot."""
        a, c = self.max_actions, self.max_transitions
        slot = len(nodes)
        rec = {
            "index": np.zeros((c, a, a), dtype=np.int64),
            "value": np.zeros((c, a, a), dtype=np.float32),
            "chance": chance,
            "ev": np.zeros((a, a), dtype=np.float32),
            "legal": np.zeros((a, a), dtype=np.float32),
        }
        rec["legal"][: spec.row_actions, : spec.col_actions] = 1.0
        nodes.append(rec)
        for r in range(spec.row_actions):
            for col in range(spec.col_actions):
                for k in range(c):
                    if not chance[k, r, col] > 0:
                        continue
                    child = self._child_spec(spec)
                    child_legal = np.zeros((a, a), dtype=np.float32)
                    child_legal[: child.row_actions, : child.col_actions] = 1.0
                    child_chance = _draw_chance(a, c, self.transition_threshold) * child_legal
                    if child.depth_bound > 0:
                        child_slot = self._build(child, child_chance, nodes)
                        rec["index"][k, r, col] = child_slot + 1      # +1: absorbing node is id 0
                        rec["value"][k, r, col] = nodes[child_slot]["root_value"]
                    else:
                        rec["value"][k, r, col] = random.choice(self.terminal_values)
                rec["ev"][r, col] = (rec["value"][:, r, col] * chance[:, r, col]).sum(dtype=np.float32)

        M = rec["ev"][: spec.row_actions, : spec.col_actions]
        joint = self._solve(M)
        p_row = torch.from_numpy(joint[: spec.row_actions]).unsqueeze(0)
        p_col = torch.from_numpy(joint[a: a + spec.col_actions]).unsqueeze(1)
        # same two fp32 matmuls as reference tree.py:306-308
        rec["root_value"] = torch.matmul(torch.matmul(p_row, torch.from_numpy(M.copy())), p_col).item()
        rec["solution"] = joint
        return slot

    def generate(self):
        """Builds the whole tree below this root and fills the seven tensors."""
        a, c = self.max_actions, self.max_transitions
        root = _NodeSpec(self, self.row_actions, self.col_actions, self.depth_bound, is_root=True)
        nodes: list = []
        self._build(root, self.chance_tensor[0].cpu().numpy(), nodes)

        if self.is_root:
            _draw_chance(a, c, 0)   # the reference constructs its absorbing node here (tree.py:336-345)
            absorbing = {
                "index": np.zeros((c, a, a), dtype=np.int64),
                "value": np.zeros((c, a, a), dtype=np.float32),
                "chance": np.zeros((c, a, a), dtype=np.float32),
                "ev": np.zeros((a, a), dtype=np.float32),
                "legal": np.zeros((a, a), dtype=np.float32),
                "root_value": 0.0,
                "solution": np.zeros(2 * a, dtype=np.float32),
            }
            absorbing["chance"][0, 0, 0] = 1.0
            absorbing["legal"][0, 0] = 1.0
            nodes.insert(0, absorbing)
        else:
            for rec in nodes:
                rec["index"] -= rec["index"] > 0

        dev = self.device
        self.index_tensor = torch.from_numpy(np.stack([n["index"] for n in nodes])).to(dev)
        self.value_tensor = torch.from_numpy(np.stack([n["value"] for n in nodes])).to(dev)
        self.chance_tensor = torch.from_numpy(np.stack([n["chance"] for n in nodes])).to(dev)
        self.expected_value_tensor = torch.from_numpy(np.stack([n["ev"] for n in nodes])[:, None]).to(dev)
        self.legal_tensor = torch.from_numpy(np.stack([n["legal"] for n in nodes])[:, None]).to(dev)
        self.root_value_tensor = torch.tensor([[n["root_value"]] for n in nodes], dtype=torch.float, device=dev)
        self.solution_tensor = torch.from_numpy(np.stack([n["solution"] for n in nodes])).to(dev)
        if self.is_root:
            self.hash = torch.randint(-(2 ** 63), 2 ** 63 - 1, size=(1,)).item()
        self._packed = None

    def generate_fast(self, seed: int = 0, child_spec=None, **kwargs):
        """
        Level-synchronous, batched construction on `self.device` (environment/fast_tree.py): the way to build the
        large trees the recursive generator cannot (millions of nodes).  Level-order node ids; ignores the three
        per-node Python lambdas (`child_spec` is their vectorised counterpart).
        """
        from environment.fast_tree import generate_fast

        if not self.is_root:
            raise Exception("generate_fast builds whole trees (is_root=True)")
        return generate_fast(self, seed=seed, child_spec=child_spec, **kwargs)

    # --------------------------------------------------------------- checks

    def assert_index_is_tree(self):
        """
        Non-zero entries of `index_tensor` are a bijection onto
        [1 + is_root, size) and every child id exceeds its parent's id
        (reference tree.py:368-383).
        """
        idx = self.index_tensor
        nz = idx[idx != 0]
        expected = torch.arange(1 + self.is_root, 1 + self.is_root + nz.numel(), device=idx.device)
        assert torch.equal(torch.sort(nz).values, expected)
        parent = torch.arange(idx.shape[0], device=idx.device).view(-1, 1, 1, 1)
        assert bool(torch.all((idx == 0) | (idx > parent)))

    # ---------------------------------------------------------- persistence

    @staticmethod
    def _saved_trees_dir():
        return os.path.join(os.path.dirname(os.path.realpath(__file__)), "..", "saved_trees")

    def save(self, directory_name=None):
        """Writes saved_trees/<directory_name>/tree.tar and saved_trees/recent/tree.tar."""
        if not self.is_root:
            raise Exception("Attempting to save non-root tree")
        base = self._saved_trees_dir()
        if directory_name is None:
            directory_name = str(int(time.time()))
        path = os.path.join(base, directory_name)
        recent = os.path.join(base, "recent")
        for d in (base, path, recent):
            os.makedirs(d, exist_ok=True)
        payload = {key: self.__dict__[key] for key in self.saved_keys}
        torch.save(payload, os.path.join(recent, "tree.tar"))
        torch.save(payload, os.path.join(path, "tree.tar"))
        logging.info("saving trees to '{}' and 'recent'".format(path))

    def load(self, directory_name="recent"):
        """Overwrites this tree's data with saved_trees/<directory_name>/tree.tar."""
        path = os.path.join(self._saved_trees_dir(), directory_name, "tree.tar")
        logging.info("loading tree from '{}'".format(directory_name))
        payload: Dict = torch.load(path)
        for key, value in payload.items():
            self.__dict__[key] = value
        self._packed = None
        logging.info("loaded tree has hash {}".format(self.hash))

    def to(self, device):
        """Moves every member tensor to `device` and updates `self.device`."""
        self.device = device
        for key, value in self.__dict__.items():
            if torch.is_tensor(value):
                self.__dict__[key] = value.to(device)
        self._packed = None

    # ------------------------------------------------------- kernel tables

    def packed(self):
        """
        Compact device-resident node table the CUDA kernels read (built once
        per tree/device by `rnad_tree_pack`; see include/rnad_b200.h).
        """
        import _b200

        key = (self.index_tensor.data_ptr(), self.value_tensor.data_ptr(), self.chance_tensor.data_ptr(),
               self.expected_value_tensor.data_ptr(), self.legal_tensor.data_ptr())
        if self._packed is None or self._packed.key != key:
            self._packed = _b200.PackedTree(self, key)
        return self._packed
