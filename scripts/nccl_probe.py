"""Times torch.distributed all_gather_into_tensor at the bench's update-block size (diagnostic)."""
import os, sys, time
import torch, torch.distributed as dist
rank=int(os.environ["RANK"]); world=int(os.environ["WORLD_SIZE"]); local=int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
for nbytes in (4096, 65536, 1572864, 16<<20):
    n=nbytes//4
    a=torch.ones(n,dtype=torch.int32,device="cuda"); b=torch.empty(n*world,dtype=torch.int32,device="cuda")
    for _ in range(5): dist.all_gather_into_tensor(b,a)
    torch.cuda.synchronize(); dist.barrier()
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): dist.all_gather_into_tensor(b,a)
    e1.record(); torch.cuda.synchronize()
    if rank==0: print(f"all_gather {nbytes} B/rank x{world}: {e0.elapsed_time(e1)/20*1e3:.1f} us", flush=True)
# with compute in between (like the bench)
x=torch.randn(4096,4096,device="cuda")
a=torch.ones(393216,dtype=torch.int32,device="cuda"); b=torch.empty(393216*world,dtype=torch.int32,device="cuda")
for _ in range(3):
    y=x@x; dist.all_gather_into_tensor(b,a)
torch.cuda.synchronize(); dist.barrier()
e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    y=x@x
e1.record(); torch.cuda.synchronize(); t_mm=e0.elapsed_time(e1)/20
e0.record()
for _ in range(20):
    y=x@x; dist.all_gather_into_tensor(b,a)
e1.record(); torch.cuda.synchronize(); t_both=e0.elapsed_time(e1)/20
if rank==0: print(f"matmul {t_mm*1e3:.0f} us; matmul + all_gather {t_both*1e3:.0f} us", flush=True)
dist.destroy_process_group()
