"""One rank of the peer-window exchange test (tests/test_peer_exchange_gpu.py starts WORLD_SIZE of these, gloo for the
handles, every rank on the CUDA device PEER_DEVICE or on device RANK when that is -1)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from iboat_pmcts_b200.engine import Engine  # noqa: E402
from iboat_pmcts_b200.peer import open_exchange  # noqa: E402


def counters(rank, rnd, count, top):
    return np.random.default_rng(1000 * rnd + rank).integers(0, top + 1, count).astype(np.uint32)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    device = int(os.environ.get("PEER_DEVICE", "0"))
    device = rank if device < 0 else device
    rounds = int(os.environ.get("PEER_ROUNDS", "6"))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.cuda.set_device(device)
    dev = torch.device("cuda", device)
    eng = Engine(device)
    eng.use_torch_stream()
    n_scen, n_leaf, n_nodes = 2, 501, 1 + 8 + 64  # per rank: 2 scenarios x 501 leaves (a ragged last vector) of a depth-2 tree
    count = n_scen * n_leaf * 12
    for width, top in ((1, 255), (2, 40000)):
        dtype = torch.uint8 if width == 1 else torch.int16
        block = (count * width + 15) // 16 * 16
        px, why = open_exchange(eng, dist, world, rank, block)
        assert px is not None, why
        parent = np.concatenate([[-1], np.zeros(8), 1 + np.arange(64) // 8]).astype(np.int32)
        arm = np.concatenate([[0], np.arange(8), np.arange(64) % 8]).astype(np.int8)
        leaf_node = np.random.default_rng(5).integers(9, n_nodes, world * n_scen * n_leaf).astype(np.int32)
        slots = np.random.default_rng(6).integers(0, 8, world * n_scen * n_leaf).astype(np.int8)
        t = lambda a: torch.from_numpy(a).to(dev)  # noqa: E731
        parent_d, arm_d, leaf_d, slots_d = t(parent), t(arm), t(leaf_node), t(slots)
        got_table = torch.zeros(world * n_scen * n_nodes * 96, dtype=torch.int32, device=dev)
        want_table = torch.zeros_like(got_table)
        for rnd in range(rounds):
            mine = t(counters(rank, rnd, count, top).view(np.int32))
            px.put_bins(mine, width)
            gathered = px.wait()[:world * block].view(world, block)[:, :count * width].contiguous().view(dtype)
            eng.backup(got_table, n_nodes, parent_d, arm_d, leaf_d, slots_d, gathered.reshape(-1), n_scen=world * n_scen)
            every = np.concatenate([counters(r, rnd, count, top) for r in range(world)])
            eng.backup(want_table, n_nodes, parent_d, arm_d, leaf_d, slots_d, t(every.view(np.int32)), n_scen=world * n_scen)
            if rnd % 2:  # (a rank may run one put ahead of the slowest reader: leave some rounds unsynchronised)
                torch.cuda.synchronize()
                wide = gathered.cpu().numpy().view(np.uint8 if width == 1 else np.uint16).astype(np.uint32)
                assert np.array_equal(wide.reshape(-1), every), "round %d" % rnd
        torch.cuda.synchronize()
        assert torch.equal(got_table, want_table) and int(got_table.sum()) > 0
        # a block sent as it is (the update block of the native search: paths + counts)
        raw = t(np.random.default_rng(77 + rank).integers(0, 256, block - 5).astype(np.uint8))
        px.put(raw)
        back = px.wait().view(world, block)
        torch.cuda.synchronize()
        for r in range(world):
            want = np.random.default_rng(77 + r).integers(0, 256, block - 5).astype(np.uint8)
            assert np.array_equal(back[r, :block - 5].cpu().numpy(), want)
        dist.barrier()
        px.close()
    print("rank %d ok" % rank, flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
