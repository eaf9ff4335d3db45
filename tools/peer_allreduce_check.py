"""Multi-GPU check of the peer-memory all-reduce (run under torchrun on >= 2 GPUs):
compares it with NCCL's all_reduce on random buffers, eagerly and through CUDA-graph replay."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from egt_b200 import dp

rank, local, world = int(os.environ['RANK']), int(os.environ['LOCAL_RANK']), int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
n = 17000
peer = dp._peer_allreduce_for(torch.zeros(n, device=dev))
assert peer is not None, 'peer all-reduce unavailable'
ok = True
for it in range(6):
    g = torch.Generator(device='cpu').manual_seed(100 * it + rank)
    x = torch.randn(n, generator=g).to(dev)
    ref = x.clone()
    dist.all_reduce(ref)
    y = x.clone()
    peer(y)
    torch.cuda.synchronize()
    err = float((y - ref).abs().max())
    ok &= err < 1e-5
    if rank == 0:
        print(f'eager call {it}: max |peer - nccl| = {err:.2e}', flush=True)
# graph replay: static buffer, new contents every replay
buf = torch.zeros(n, device=dev)
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    peer(buf)
torch.cuda.current_stream().wait_stream(s)
torch.cuda.synchronize()
gr = torch.cuda.CUDAGraph()
with torch.cuda.graph(gr):
    peer(buf)
for it in range(5):
    g = torch.Generator(device='cpu').manual_seed(7000 + 100 * it + rank)
    x = torch.randn(n, generator=g).to(dev)
    ref = x.clone()
    dist.all_reduce(ref)
    buf.copy_(x)
    gr.replay()
    torch.cuda.synchronize()
    err = float((buf - ref).abs().max())
    ok &= err < 1e-5
    if rank == 0:
        print(f'graph replay {it}: max |peer - nccl| = {err:.2e}', flush=True)
# latency: 50 back-to-back calls per graph replay (ranks stay within one call of each other), CUDA events, max over ranks
gl = torch.cuda.CUDAGraph()
with torch.cuda.graph(gl):
    for _ in range(50):
        peer(buf)
for _ in range(3):
    gl.replay()
torch.cuda.synchronize()
dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    gl.replay()
e1.record()
torch.cuda.synchronize()
us = torch.tensor([e0.elapsed_time(e1) / 1000 * 1e3], device=dev)
dist.all_reduce(us, op=dist.ReduceOp.MAX)
e0.record()
for _ in range(200):
    dist.all_reduce(buf)
e1.record()
torch.cuda.synchronize()
us_nccl = torch.tensor([e0.elapsed_time(e1) / 200 * 1e3], device=dev)
dist.all_reduce(us_nccl, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f'latency of one {n}-float all-reduce on {world} GPUs: peer kernel {float(us):.1f} us (graph replay), NCCL {float(us_nccl):.1f} us (eager)', flush=True)
t = torch.tensor([int(ok)], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print('PEER_ALLREDUCE_OK' if int(t.item()) else 'PEER_ALLREDUCE_MISMATCH', flush=True)
sys.stdout.flush()
torch.cuda.synchronize()
dist.barrier()
os._exit(0)
