import os, torch, torch.distributed as dist, time
rank=int(os.environ["RANK"]); world=int(os.environ["WORLD_SIZE"]); local=int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
n=16*1024*1024
a=torch.ones(n,dtype=torch.int32,device="cuda"); b=torch.empty_like(a)
def xfer():
    ops=[]
    if rank>0: ops.append(dist.P2POp(dist.isend,a,rank-1))
    if rank<world-1: ops.append(dist.P2POp(dist.irecv,b,rank+1))
    for w in dist.batch_isend_irecv(ops): w.wait()
for _ in range(3): xfer()
torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): xfer()
e1.record(); torch.cuda.synchronize()
ms=e0.elapsed_time(e1)/10
print(f"rank {rank}: 64 MiB neighbour send/recv {ms:.3f} ms = {n*4/1e9/(ms/1e3):.1f} GB/s", flush=True)
small=torch.ones(2*(1+32768),dtype=torch.int64,device="cuda"); allb=torch.empty(world*small.numel(),dtype=torch.int64,device="cuda")
for _ in range(3): dist.all_gather_into_tensor(allb, small)
torch.cuda.synchronize(); e0.record()
for _ in range(20): dist.all_gather_into_tensor(allb, small)
e1.record(); torch.cuda.synchronize()
print(f"rank {rank}: all_gather 512 KiB/rank {e0.elapsed_time(e1)/20*1e3:.1f} us", flush=True)
dist.barrier(); dist.destroy_process_group()
