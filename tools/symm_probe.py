"""Probe: torch symmetric memory (peer-mapped buffers over NVLink) on this box — rendezvous, peer pointers, barrier."""
import os
import time

import torch
import torch.distributed as dist


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import torch.distributed._symmetric_memory as symm_mem
    rank, world = dist.get_rank(), dist.get_world_size()
    t0 = time.perf_counter()
    buf = symm_mem.empty((64 << 20,), dtype=torch.float32, device="cuda")
    hdl = symm_mem.rendezvous(buf, dist.group.WORLD)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    print(rank, "rendezvous %.1f ms" % ((t1 - t0) * 1e3), "ptrs", [hex(p) for p in hdl.buffer_ptrs][:4], "size", hdl.buffer_size,
          "signal pads", len(hdl.signal_pad_ptrs), flush=True)
    buf.fill_(float(rank + 1))
    hdl.barrier(channel=0)
    peer = hdl.get_buffer((rank + 1) % world, (64 << 20,), torch.float32)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dst = torch.empty_like(buf)
    dst.copy_(peer)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        dst.copy_(peer)
    e1.record()
    torch.cuda.synchronize()
    print(rank, "peer value", dst[0].item(), "pull GB/s %.1f" % (5 * buf.numel() * 4 / (e0.elapsed_time(e1) * 1e-3) / 1e9), flush=True)
    hdl.barrier(channel=0)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
