"""The PCIe / host-memory ceiling of the end-to-end leg, on its own: the bytes one step of the headline workload moves
(H2D 506 MB, D2H 504 MB) between pinned host memory and the device on two streams, no kernels, every rank at once.
  python tools/pcie_probe.py                       (one GPU)
  torchrun --nproc-per-node N tools/pcie_probe.py  (N GPUs of one box: the aggregate is what the host side sustains)
`bench.py` runs the same probe after its e2e leg and reports it as `e2e.copy_only`."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench

H2D, D2H, STEPS = 506273792, 503802292, 20


def main():
    rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("LOCAL_RANK", "0"), ("WORLD_SIZE", "1")))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    secs = bench.copy_only_probe(H2D, D2H, STEPS, barrier)
    if world > 1:
        t = torch.tensor([secs], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        secs = float(t.item())
    if rank == 0:
        print(json.dumps({"n_gpus": world, "h2d_bytes_per_step": H2D, "d2h_bytes_per_step": D2H, "steps": STEPS,
                          "gbs_each_way_per_gpu": round(H2D * STEPS / secs / 1e9, 1),
                          "gbs_each_way_all_gpus": round(world * H2D * STEPS / secs / 1e9, 1),
                          "scans_per_s_ceiling": round(world * 256 * STEPS / secs)}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
