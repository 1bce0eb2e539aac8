"""Ceiling of the host <-> device links of this box with N ranks copying at once.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 \
        scripts/host_link_probe.py [out.json]

Every rank streams a pinned host buffer to its GPU (H2D) and a device buffer back to pinned host memory (D2H) on two
streams for a fixed number of rounds, first one direction at a time, then both at once (what the e2e leg of bench.py
does: 16-bit PCM in, float32 features out).  Timed with CUDA events between barriers; rank 0 prints per-rank and
aggregate GB/s.  The e2e leg moves `h2d_bytes_per_step + d2h_bytes_per_step` per rank per step; dividing by the
both-directions rate here gives the floor of its step time on this box.
"""
import json
import os
import sys

import torch
import torch.distributed as dist


def main():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n_in, n_out = 1 << 30, 1 << 30                 # 1 GiB each way per round
    h_in = torch.empty(n_in, dtype=torch.uint8, pin_memory=True)
    h_out = torch.empty(n_out, dtype=torch.uint8, pin_memory=True)
    h_in.fill_(1)
    d_in = torch.empty(n_in, dtype=torch.uint8, device=dev)
    d_out = torch.ones(n_out, dtype=torch.uint8, device=dev)
    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    rounds = 6

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run(do_in, do_out):
        barrier()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        main = torch.cuda.current_stream()
        e0.record(main)
        s_in.wait_event(e0); s_out.wait_event(e0)
        for _ in range(rounds):
            if do_in:
                with torch.cuda.stream(s_in):
                    d_in.copy_(h_in, non_blocking=True)
            if do_out:
                with torch.cuda.stream(s_out):
                    h_out.copy_(d_out, non_blocking=True)
        e1.record(s_in); e2.record(s_out)
        main.wait_event(e1); main.wait_event(e2)
        end = torch.cuda.Event(enable_timing=True)
        end.record(main)
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(end)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        barrier()
        nbytes = rounds * ((n_in if do_in else 0) + (n_out if do_out else 0))
        return nbytes / (float(ms.item()) * 1e-3) / 1e9        # GB/s per rank, slowest rank's time

    run(True, True)                                            # warm-up
    res = {"n_gpus": world, "per_rank_GBps": {"h2d_only": run(True, False), "d2h_only": run(False, True),
                                              "both_directions_sum": run(True, True)}}
    # the pattern of the e2e leg: 3.79 GB in, 2.49 GB out per step -> bytes-weighted both-direction run
    n_in_e2e, n_out_e2e = int(n_in * 0.6), int(n_out * 0.4)

    def run_ratio():
        barrier()
        e0 = torch.cuda.Event(enable_timing=True)
        main = torch.cuda.current_stream()
        e0.record(main)
        s_in.wait_event(e0); s_out.wait_event(e0)
        for _ in range(rounds):
            with torch.cuda.stream(s_in):
                d_in[:n_in_e2e].copy_(h_in[:n_in_e2e], non_blocking=True)
            with torch.cuda.stream(s_out):
                h_out[:n_out_e2e].copy_(d_out[:n_out_e2e], non_blocking=True)
        main.wait_stream(s_in); main.wait_stream(s_out)
        end = torch.cuda.Event(enable_timing=True)
        end.record(main)
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(end)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        barrier()
        return rounds * (n_in_e2e + n_out_e2e) / (float(ms.item()) * 1e-3) / 1e9
    res["per_rank_GBps"]["both_60_40_like_e2e"] = run_ratio()
    # write-combined pinned memory as the H2D source (cudaHostAllocWriteCombined through the runtime)
    try:
        import ctypes
        rt = None
        with open("/proc/self/maps") as f:
            for ln in f:
                if "libcudart" in ln:
                    rt = ctypes.CDLL(ln.split()[-1])
                    break
        if rt is not None:
            ptr = ctypes.c_void_p()
            rt.cudaHostAlloc.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_size_t, ctypes.c_uint]
            rt.cudaMemcpyAsync.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
            if rt.cudaHostAlloc(ctypes.byref(ptr), n_in, 0x04) == 0:       # cudaHostAllocWriteCombined
                ctypes.memset(ptr, 1, n_in)

                def run_wc(do_out):
                    barrier()
                    e0 = torch.cuda.Event(enable_timing=True)
                    main = torch.cuda.current_stream()
                    e0.record(main)
                    s_in.wait_event(e0); s_out.wait_event(e0)
                    for _ in range(rounds):
                        rt.cudaMemcpyAsync(ctypes.c_void_p(d_in.data_ptr()), ptr, n_in, 1, ctypes.c_void_p(s_in.cuda_stream))
                        if do_out:
                            with torch.cuda.stream(s_out):
                                h_out.copy_(d_out, non_blocking=True)
                    main.wait_stream(s_in); main.wait_stream(s_out)
                    end = torch.cuda.Event(enable_timing=True)
                    end.record(main)
                    torch.cuda.synchronize()
                    ms = torch.tensor([e0.elapsed_time(end)], device=dev, dtype=torch.float64)
                    if world > 1:
                        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
                    barrier()
                    return rounds * (n_in + (n_out if do_out else 0)) / (float(ms.item()) * 1e-3) / 1e9
                res["per_rank_GBps"]["h2d_only_write_combined_source"] = run_wc(False)
                res["per_rank_GBps"]["both_directions_sum_write_combined_source"] = run_wc(True)
                rt.cudaFreeHost(ptr)
    except Exception as e:      # diagnostics only
        res["write_combined_error"] = repr(e)
    res["aggregate_GBps"] = {k: v * world for k, v in res["per_rank_GBps"].items()}
    res["cpu_count"] = os.cpu_count()
    res["how"] = f"{rounds} rounds x 1 GiB per direction per rank, pinned host memory, max over ranks of the CUDA-event time"
    if rank == 0:
        print(json.dumps(res))
        if len(sys.argv) > 1:
            with open(sys.argv[1], "w") as f:
                json.dump(res, f, indent=1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
