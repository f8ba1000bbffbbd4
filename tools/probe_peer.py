"""Probe of the peer-memory plumbing on a multi-GPU box (run under torchrun, one rank per GPU):
torch symmetric memory (rendezvous, peer pointers, multicast pointer, barrier) and plain peer copies.
Prints one JSON line per check; nothing here is on the product path."""
import json
import os
import time

import torch
import torch.distributed as dist


def main():
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
    dev = torch.device('cuda', int(os.environ['LOCAL_RANK']))
    dist.init_process_group('nccl', device_id=dev)
    out = {'rank': rank, 'world': world}
    try:
        import torch.distributed._symmetric_memory as symm
        t = symm.empty(64 << 20, dtype=torch.float32, device=dev)
        hdl = symm.rendezvous(t, dist.group.WORLD.group_name)
        out['symm'] = True
        out['buffer_ptrs'] = [hex(p) for p in hdl.buffer_ptrs]
        out['multicast_ptr'] = hex(getattr(hdl, 'multicast_ptr', 0) or 0)
        out['signal_pad_ptrs'] = len(hdl.signal_pad_ptrs)
        t.fill_(float(rank))
        hdl.barrier(channel=0)
        peer = hdl.get_buffer((rank + 1) % world, (1024,), torch.float32)
        out['peer_first'] = float(peer[0].item())
        hdl.barrier(channel=0)
        # peer write bandwidth with a plain copy kernel into the peer's buffer
        src = torch.ones(64 << 20, dtype=torch.float32, device=dev)
        dst = hdl.get_buffer((rank + 1) % world, (64 << 20,), torch.float32)
        for _ in range(3):
            dst.copy_(src)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            dst.copy_(src)
        e1.record()
        torch.cuda.synchronize()
        out['peer_copy_GBps'] = 10 * src.numel() * 4 / (e0.elapsed_time(e1) * 1e-3) / 1e9
        e0.record()
        for _ in range(50):
            hdl.barrier(channel=0)
        e1.record()
        torch.cuda.synchronize()
        out['barrier_us'] = e0.elapsed_time(e1) * 1e3 / 50
    except Exception as e:      # noqa: BLE001
        out['symm'] = False
        out['symm_error'] = repr(e)[:400]
    # NCCL all-gather bandwidth for comparison
    blk = torch.ones(32 << 20, dtype=torch.float32, device=dev)
    full = torch.empty(world * blk.numel(), dtype=torch.float32, device=dev)
    for _ in range(3):
        dist.all_gather_into_tensor(full, blk)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        dist.all_gather_into_tensor(full, blk)
    e1.record()
    torch.cuda.synchronize()
    out['nccl_allgather_recv_GBps'] = 10 * (world - 1) * blk.numel() * 4 / (e0.elapsed_time(e1) * 1e-3) / 1e9
    print(json.dumps(out), flush=True)
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
