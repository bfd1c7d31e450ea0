"""Worker for the multi-GPU parity test (launched with torch.distributed.run, one process per GPU).

Every rank owns its z-blocks of a seeded volume, runs the sharded bit-plane engine (peer ghost planes over
NVLink), the pieces are gathered on rank 0 and compared bit for bit with (a) the single-GPU result of the same
library and (b) the CPU oracle when the volume is small enough.
usage: multi_gpu_worker.py d0 d1 d2 gens rule block backend
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    d0, d1, d2, gens, rule, block = (int(v) for v in sys.argv[1:7])
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    rank, world = dist.get_rank(), dist.get_world_size()

    import clap_b200
    from clap_b200.slab import ShardedVolume, torch_all_gather_bytes
    clap_b200.init(local)

    rng = np.random.default_rng(1234)
    full = (rng.integers(1, 6, (d2, d1, d0)) * (rng.random((d2, d1, d0)) < 0.3)).astype(np.uint8)
    vol = ShardedVolume(d0, d1, d2, rank, world, gens, int(full.max()), block, torch_all_gather_bytes(dist, dev))
    mine = np.ascontiguousarray(full[vol.zglobal]) if vol.n_local else np.zeros((1, d1, d0), np.uint8)
    ok = True
    for rep in range(2):                                   # the second pass re-uses the mapped halo regions
        vol.upload(mine)
        dist.barrier()
        vol.prepare(rule, gens)
        torch.cuda.synchronize()
        dist.barrier()
        pop = vol.run()
        out = np.empty_like(mine)
        vol.download(out)
        tp = torch.tensor([pop], dtype=torch.int64, device=dev)
        dist.all_reduce(tp)
        # gather the pieces on rank 0
        pieces = [None] * world
        dist.all_gather_object(pieces, (vol.zglobal, out[:vol.n_local]))
        if rank == 0:
            got = np.empty_like(full)
            for zs, arr in pieces:
                if len(zs):
                    got[zs] = arr
            want = full.copy()
            wpop = clap_b200.ca3d_run(want, rule, gens)
            same = np.array_equal(got, want) and int(tp[0]) == wpop
            msg = f"rep {rep}: sharded vs single-GPU: {'equal' if same else 'DIFFERENT'} (pop {int(tp[0])} vs {wpop})"
            if full.size <= 1 << 22:
                import oracle_lib
                ora = oracle_lib.port()
                chk = full.copy()
                s, b, n = ora.ca3d_rule(rule)
                opop = ora.ca3d_run(chk, s, b, n, gens)
                same = same and np.array_equal(got, chk) and opop == wpop
                msg += f"; vs oracle: {'equal' if np.array_equal(got, chk) else 'DIFFERENT'}"
            print(msg, flush=True)
            ok = ok and same
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    dist.barrier()
    vol.close()
    dist.destroy_process_group()
    if rank == 0:
        print("MULTI_GPU_OK" if ok else "MULTI_GPU_FAIL", flush=True)
    sys.exit(0 if int(flag[0]) else 1)


if __name__ == "__main__":
    main()
