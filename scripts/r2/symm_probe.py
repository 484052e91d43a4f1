"""2-GPU probe: does torch symmetric memory give usable peer pointers on this box?
torchrun --nproc-per-node 2 scripts/r2/symm_probe.py"""
import os, time, torch, torch.distributed as dist
import torch.distributed._symmetric_memory as sm
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
t0 = time.time()
t = sm.empty(1 << 20, dtype=torch.float32, device="cuda")
h = sm.rendezvous(t, dist.group.WORLD)
print(rank, "rendezvous s", time.time() - t0, "ptrs", [hex(p) for p in h.buffer_ptrs], "sig", [hex(p) for p in h.signal_pad_ptrs],
      "multicast", h.has_multicast_support, hex(h.multicast_ptr) if h.has_multicast_support else None, flush=True)
t.fill_(float(rank + 1))
h.barrier()
peer = h.get_buffer((rank + 1) % world, (1 << 20,), torch.float32)
print(rank, "peer value", float(peer[123]), flush=True)
# peer bandwidth: copy 256 MB from peer
big = sm.empty(64 << 20, dtype=torch.float32, device="cuda"); hb = sm.rendezvous(big, dist.group.WORLD)
big.fill_(rank); hb.barrier()
pb = hb.get_buffer((rank + 1) % world, (64 << 20,), torch.float32)
dst = torch.empty_like(big)
for _ in range(3): dst.copy_(pb)
torch.cuda.synchronize(); e0 = torch.cuda.Event(True); e1 = torch.cuda.Event(True)
e0.record()
for _ in range(10): dst.copy_(pb)
e1.record(); torch.cuda.synchronize()
print(rank, "peer read GB/s", 10 * 256e-3 * 1.048576 / (e0.elapsed_time(e1) * 1e-3) , flush=True)
print(rank, "can_access_peer", torch.cuda.can_device_access_peer(rank, (rank + 1) % world), flush=True)
hb.barrier()
dist.destroy_process_group()
