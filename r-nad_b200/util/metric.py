"""
Exact exploitability ("NashConv") of a joint policy on a Tree - API mirror of the
reference `util/metric.py` (`NashConvData` :10-190, `kld` :193-210).

The reference walks the tree with a recursive Python DFS on the CPU (0.8 ms per
node, and it moves the whole tree to the CPU and back on every evaluation,
metric.py:84-88).  Here the same quantities are computed level-synchronously with
batched tensor ops on the tree's own device: children always have larger ids than
their parent, so one sweep from the deepest level up gives every node's best
responses, and one sweep down gives the reach probabilities.

Reference behaviour that is kept on purpose:
  * `get_nashconv(tree, joint_policy)` honours its `joint_policy` argument only at
    `state_index`; every node below uses `self.joint_policy` (metric.py:148-150).
    `get_nashconv_from_net` fills `self.joint_policy` first, so the production path
    is the true NashConv; the reference's own test relies on the quirk.
  * the reach probability uses pi_col[r]*pi_row[c] for joint action (r, c)
    (metric.py:130-132), which only matters for the informational `reach_probability`.
"""

from typing import Dict

import torch


class NashConvData:
    def __init__(self, tree):
        self.size = tree.value_tensor.shape[0]
        dev = tree.device
        self.joint_policy = torch.zeros((self.size, 2 * tree.max_actions), device=dev, dtype=torch.float)
        self.row_best = torch.zeros((self.size,), device=dev, dtype=torch.float)
        self.col_best = torch.zeros((self.size,), device=dev, dtype=torch.float)
        self.reach_probability = torch.zeros((self.size,), device=dev, dtype=torch.float)
        self.depth = torch.zeros((self.size,), device=dev, dtype=torch.int)

    def to(self, device):
        for key, value in self.__dict__.items():
            if torch.is_tensor(value):
                self.__dict__[key] = value.to(device)

    def get_nashconv_from_net(self, tree, net, inference_batch_size: int = 10 ** 5) -> None:
        """Infers the net's policy for both players at every node, then evaluates it (metric.py:51-90)."""
        net.eval()
        a = tree.max_actions
        inference_batch_size = int(inference_batch_size)
        for start in range(0, self.size, inference_batch_size):
            stop = min(start + inference_batch_size, self.size)
            value_slice = tree.expected_value_tensor[start:stop]
            legal_slice = tree.legal_tensor[start:stop]
            with torch.no_grad():
                self.joint_policy[start:stop, :a] = net.forward_policy(torch.cat([value_slice, legal_slice], dim=1))
                self.joint_policy[start:stop, a:] = net.forward_policy(
                    torch.cat([-value_slice, legal_slice], dim=1).swapaxes(2, 3).contiguous())
        self.get_nashconv(tree, self.joint_policy)
        net.train()

    def get_nashconv(self, tree, joint_policy: torch.Tensor, state_index: int = 1, reach_probablity: float = 1,
                     depth: int = 0) -> None:
        """
        Best-response values of both players against `joint_policy` (size, 2A) for the
        subtree under `state_index`; fills row_best, col_best, reach_probability, depth
        (metric.py:93-175).  NashConv of the game = row_best[1] + col_best[1].
        """
        a = tree.max_actions
        dev = tree.index_tensor.device
        policy = self.joint_policy.to(dev).clone()
        policy[state_index] = joint_policy[state_index].to(dev)

        index, chance, value = tree.index_tensor, tree.chance_tensor, tree.value_tensor
        legal = tree.legal_tensor

        # levels of the subtree, following the edges the reference follows (chance > 0, child != 0)
        levels = [torch.tensor([state_index], dtype=torch.long, device=dev)]
        while True:
            nodes = levels[-1]
            children = index[nodes][(chance[nodes] > 0) & (index[nodes] != 0)]
            if children.numel() == 0:
                break
            levels.append(children)

        row_best, col_best, node_depth = self.row_best.to(dev), self.col_best.to(dev), self.depth.to(dev)
        for nodes in reversed(levels):
            idx, ch, val = index[nodes], chance[nodes], value[nodes]          # (n, C, A, A)
            live = ch > 0
            leaf = idx == 0
            zero = torch.zeros_like(val)
            row_case = torch.where(live, torch.where(leaf, val, row_best[idx]) * ch, zero).sum(dim=1)   # (n, A, A)
            col_case = torch.where(live, torch.where(leaf, -val, col_best[idx]) * ch, zero).sum(dim=1)
            pi_row, pi_col = policy[nodes, :a], policy[nodes, a:]
            row_resp = torch.matmul(row_case, pi_col.unsqueeze(-1)).squeeze(-1)                          # (n, A) over rows
            col_resp = torch.matmul(pi_row.unsqueeze(1), col_case).squeeze(1)                            # (n, A) over cols
            neg_inf = torch.full_like(row_resp, float("-inf"))
            row_best[nodes] = torch.where(legal[nodes, 0, :, 0] != 0, row_resp, neg_inf).max(dim=-1).values
            col_best[nodes] = torch.where(legal[nodes, 0, 0, :] != 0, col_resp, neg_inf).max(dim=-1).values
            child_depth = torch.where(live & ~leaf, node_depth[idx], torch.zeros_like(node_depth[idx]))
            node_depth[nodes] = 1 + child_depth.reshape(nodes.numel(), -1).max(dim=-1).values

        reach = self.reach_probability.to(dev)
        reach[state_index] = reach_probablity
        for nodes in levels[:-1]:
            idx, ch = index[nodes], chance[nodes]
            pi_row, pi_col = policy[nodes, :a], policy[nodes, a:]
            joint = pi_col.unsqueeze(-1) * pi_row.unsqueeze(1)                 # [n, r, c] = pi_col[r] * pi_row[c]
            child_reach = (reach[nodes].view(-1, 1, 1, 1) * joint.unsqueeze(1)) * ch
            follow = (ch > 0) & (idx != 0)
            reach[idx[follow]] = child_reach[follow]

        self.row_best, self.col_best, self.depth, self.reach_probability = row_best, col_best, node_depth, reach

    def mean_nashconv_by_depth(self) -> Dict[int, float]:
        """Mean of row_best + col_best over the nodes whose longest path to a terminal has each length."""
        max_depth = int(self.depth[1].item())
        nashconv = self.row_best + self.col_best
        means: Dict[int, float] = {}
        for depth in range(1, max_depth + 1):
            means[depth] = torch.mean(nashconv[self.depth == depth]).item()
        return means


def kld(p: torch.Tensor, q: torch.Tensor, valid: torch.Tensor, legal_actions: torch.Tensor, valid_count: int = None):
    """Mean over valid steps of KL(p || q) restricted to legal actions (metric.py:193-210)."""
    if valid_count is None:
        valid_count = valid.sum().item()
    where = (valid.unsqueeze(-1) * legal_actions).to(torch.bool)
    terms = torch.where(where, p * (torch.log(p) - torch.log(q)), torch.zeros_like(p))
    return terms.sum().item() / valid_count
