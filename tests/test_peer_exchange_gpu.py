"""The update exchange over peer-mapped windows (pmcts_peer_*, csrc/pmcts_peer.cu): several processes -- one per rank as
in a multi-GPU search -- put their narrowed histograms into each other's windows, wait on the sequence numbers with
stream memory operations and back the gathered block up; the master tables equal those of the unpacked blocks.  With
one GPU the ranks share it (the windows are then mapped between processes of one device); with two or more every rank
takes its own."""
import os
import socket
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def run_ranks(world, device):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    procs = []
    for rank in range(world):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   PEER_DEVICE=str(device))
        procs.append(subprocess.Popen([sys.executable, os.path.join(HERE, "peer_worker.py")], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = []
    try:
        for p in procs:
            outs.append(p.communicate(timeout=240)[0])
    finally:
        for p in procs:
            if p.poll() is None:
                p.kill()
    for rank, (p, out) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and "rank %d ok" % rank in out, out[-3000:]


@pytest.mark.gpu
def test_two_ranks_exchange_through_windows_on_one_device():
    run_ranks(2, 0)


@pytest.mark.gpu
def test_ranks_on_their_own_devices():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("one GPU")
    run_ranks(min(n, 4), -1)
