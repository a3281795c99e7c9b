"""The fused all-gather (rvlp_logprob_batch_peers + rvlp_peer_barrier, ravest_b200/dist.py: PeerGather).

One GPU: the kernel's peer-store epilogue and the flag barrier on local buffers.  Two or more GPUs (skipped on a
one-GPU box): two processes over NCCL, the gathered vector against a single-GPU evaluation, bit for bit, through the
fused path and through the NCCL all-gather.
"""
import ctypes as C
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_peer_store_epilogue_and_barrier_on_one_gpu(cuda):
    from ravest_b200 import _lib, fit, workloads
    lib = _lib.load()
    spec, theta = workloads.make_multiplanet(3, 300, 9001, seed=5, instruments=("A", "B"), invalid_frac=0.03)
    post = fit.from_spec(spec)
    th = cuda.as_tensor(theta, device="cuda")
    ref = post.ctx.logprob(th)
    S, lo = len(theta) + 100, 60                                    # a block inside a longer gathered vector
    bufs = [cuda.full((S,), 777.0, dtype=cuda.float64, device="cuda") for _ in range(3)]
    outs = (C.c_void_p * 3)(*[b.data_ptr() for b in bufs])
    _lib.check(lib.rvlp_logprob_batch_peers(post.ctx._h, th.data_ptr(), len(theta), outs, 3, lo, _lib.stream_ptr(0)))
    flags = cuda.zeros(128, dtype=cuda.int64, device="cuda")
    fl = (C.c_void_p * 1)(flags.data_ptr())
    for epoch in (1, 2, 3):
        _lib.check(lib.rvlp_peer_barrier(0, fl, 1, 0, epoch, 1000, _lib.stream_ptr(0)))
    cuda.cuda.synchronize()
    assert flags[0].item() == 3 and flags[8].item() == 0             # arrived at epoch 3, no timeout
    for b in bufs:
        assert cuda.equal(b[lo:lo + len(theta)].view(cuda.int64), ref.view(cuda.int64))
        assert bool((b[:lo] == 777.0).all()) and bool((b[lo + len(theta):] == 777.0).all())
    # a rank that never arrives: the wait gives up and says so instead of hanging the GPU
    two = (C.c_void_p * 2)(flags.data_ptr(), flags.data_ptr() + 512)
    _lib.check(lib.rvlp_peer_barrier(0, two, 2, 0, 9, 300, _lib.stream_ptr(0)))
    cuda.cuda.synchronize()
    assert flags[8].item() == 1


CHILD = r"""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, %r)
from ravest_b200 import dist as rdist, fit, workloads
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
S = 30_002                                                          # unequal shards
spec, theta = workloads.make_c3(n_samples=S)
post = fit.from_spec(spec)
lo, hi = rdist.shard_bounds(S, world, rank)
th = torch.as_tensor(theta[lo:hi], device="cuda")
ref = post.ctx.logprob(torch.as_tensor(theta, device="cuda"))
outs = []
for it in range(4):                                                 # both halves of the double buffer, twice
    g = rdist.sharded_logprob(lambda t: post.ctx.logprob(t), th, n_samples=S, theta_is_local=True, ctx=post.ctx, fused=True)
    outs.append(bool(torch.equal(g.view(torch.int64), ref.view(torch.int64))))
pgs = list(post.ctx._peer_gathers.values())
fused = bool(pgs) and pgs[0] is not None
timed_out = pgs[0].timed_out() if fused else False
n = rdist.sharded_logprob(lambda t: post.ctx.logprob(t), th, n_samples=S, theta_is_local=True, fused=False)   # NCCL
ok_nccl = bool(torch.equal(n.view(torch.int64), ref.view(torch.int64)))
res = [None] * world
dist.all_gather_object(res, (outs, fused, timed_out, ok_nccl))
if rank == 0:
    print("RESULT", res)
dist.destroy_process_group()
"""


def test_fused_gather_matches_single_gpu_on_two_gpus(cuda, tmp_path):
    if cuda.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    script = tmp_path / "fused_child.py"
    script.write_text(CHILD % ROOT)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(script)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("RESULT")][-1]
    res = eval(line[len("RESULT"):])
    for outs, fused, timed_out, ok_nccl in res:
        assert all(outs) and ok_nccl and not timed_out
        assert fused, "CUDA IPC refused: the fused path did not run"
