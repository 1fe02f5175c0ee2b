"""Host <-> device copy ceiling of this box at 1 / 2 / 4 / 8 ranks: every rank moves the byte counts of
one bench.py e2e step (pinned memory, large copies, both directions at once on two streams) with
nothing else running, so the e2e leg of bench.py can be read as a fraction of what the host's PCIe
root ports and memory controllers deliver.  Launch like bench.py:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P tools/pcie_ceiling.py [--h2d-gb 1.46 --d2h-gb 37.2 --chunk-mb 512]
Prints one JSON line on rank 0.  Measurement tool, not product code."""
import argparse
import json
import os
import time

import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--h2d-gb', type=float, default=1.46)
    ap.add_argument('--d2h-gb', type=float, default=37.2)
    ap.add_argument('--chunk-mb', type=int, default=512)
    ap.add_argument('--reps', type=int, default=2)
    a = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=dev)
    chunk = a.chunk_mb << 20
    hp = [torch.empty(chunk, dtype=torch.uint8, pin_memory=True) for _ in range(2)]   # h2d source, d2h target
    dp = [torch.empty(chunk, dtype=torch.uint8, device=dev) for _ in range(2)]
    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def run(h2d_bytes, d2h_bytes):
        n_in, n_out = int(h2d_bytes // chunk), int(d2h_bytes // chunk)
        barrier()
        t0 = time.perf_counter()
        with torch.cuda.stream(s_in):
            for _ in range(n_in):
                dp[0].copy_(hp[0], non_blocking=True)
        with torch.cuda.stream(s_out):
            for _ in range(n_out):
                hp[1].copy_(dp[1], non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), (n_in + n_out) * chunk

    res = {}
    for name, hi, ho in (('h2d_only', a.d2h_gb / 4, 0.0), ('d2h_only', 0.0, a.d2h_gb / 4),
                         ('e2e_step_mix', a.h2d_gb, a.d2h_gb), ('balanced_duplex', a.d2h_gb / 4, a.d2h_gb / 4)):
        run(hi * 1e9 / 4, ho * 1e9 / 4)                       # warm-up
        best = None
        for _ in range(a.reps):
            dt, by = run(hi * 1e9, ho * 1e9)
            if best is None or dt < best[0]:
                best = (dt, by)
        res[name] = {'seconds': best[0], 'gb_per_rank': best[1] / 1e9, 'gbs_per_rank': best[1] / best[0] / 1e9,
                     'gbs_aggregate': world * best[1] / best[0] / 1e9}
    if rank == 0:
        print(json.dumps({'tool': 'pcie_ceiling', 'ranks': world, 'chunk_mb': a.chunk_mb, **res}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
