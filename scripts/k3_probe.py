"""Device timing of K3 (pmcts_backup) on the bench's depth-4 tree with realistic leaf histograms."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from iboat_pmcts_b200.engine import Engine


def main():
    dev = torch.device("cuda:0")
    eng = Engine(0)
    eng.use_torch_stream()
    parent_h, arm_h, leaves_h, _ = bench.depth4_tree()
    n_nodes, n_leaf = parent_h.size, leaves_h.size
    rng = np.random.default_rng(0)
    for S in (8, 64):
        # 256 replicas per leaf spread over ~4 neighbouring bins, as the rollout kernels leave them
        bins = np.zeros((S, n_leaf, 12), np.uint32)
        c = rng.integers(2, 9, (S, n_leaf))
        for j in range(4):
            np.add.at(bins, (np.arange(S)[:, None], np.arange(n_leaf)[None, :], np.minimum(c + j, 11)), 64)
        parent = torch.tensor(parent_h, device=dev)
        arm = torch.tensor(arm_h, device=dev)
        leaf_node = torch.tensor(np.tile(leaves_h, S), device=dev)
        slots = torch.tensor(rng.integers(0, 8, S * n_leaf).astype(np.int8), device=dev)
        table = torch.zeros(S * n_nodes * 96, dtype=torch.int32, device=dev)
        b = torch.tensor(bins.view(np.int32).reshape(-1), device=dev)
        packed = torch.zeros(b.numel(), dtype=torch.uint8, device=dev)
        eng.pack_bins(b, packed)
        for name, src in (("uint32", b), ("uint8", packed)):
          ms = []
          table.zero_()
          for it in range(6):
              e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
              torch.cuda.synchronize()
              e0.record()
              eng.backup(table, n_nodes, parent, arm, leaf_node, slots, src, n_scen=S)
              e1.record()
              torch.cuda.synchronize()
              ms.append(e0.elapsed_time(e1))
          root = int(table.view(S, n_nodes, 96)[:, 0].sum().item())
          assert root == 6 * int(bins.sum()), (root, bins.sum())
          print("K3 %s %2d scenarios x %d leaves: %.1f us (min of %s)" % (name, S, n_leaf, min(ms) * 1e3, [round(m * 1e3, 1) for m in ms]))


if __name__ == "__main__":
    main()
