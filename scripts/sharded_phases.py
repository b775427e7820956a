"""Profiling aid: per-phase device time of the item-sharded step (torchrun, one rank per GPU)."""
import argparse, ctypes as C, os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fvx import parallel, synth
from fvx._lib import call, ptr, stream_ptr
from fvx.dataset.dataset import DataLoader
from fvx.engine import Engine

ap = argparse.ArgumentParser(); ap.add_argument("--batch", type=int, default=65536); a = ap.parse_args()
world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", local); torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
U, I, K, d, D, B = 40000 * world, 100000, 64, 20, 2048, a.batch * world
inter = synth.make_interactions(U, I, seed=1234)
p = argparse.Namespace(dataset="synthetic", batch_size=B, epochs=10 ** 6, sampler="device", seed=0)
data = DataLoader(p, interactions=inter)
lo, cnt = parallel.shard_bounds(I, world, rank)
e = Engine(U, I, K, d=d, D=D, max_batch=B, device=str(dev), use_tensor_cores=True, item_lo=lo, item_cnt=cnt)
g = torch.Generator(device=dev).manual_seed(1)
e.set_features(torch.rand(cnt, D, device=dev, generator=g), keep_fp32=False)
min_len = int(np.diff(inter.row_ptr).min())
ss = parallel.ShardedStep([e], parallel.DistGroup(), max_runs=B // max(min_len, 1) + 2)
batches = data.next_triple_batch(str(dev))
for _ in range(5):
    ss.step(*next(batches))
torch.cuda.synchronize(); dist.barrier()
names = ["run_ids", "A", "ar_S", "B1", "ar_RU", "B2", "ar_dE", "C"]
acc = {n: 0.0 for n in names}
N = 10
for _ in range(N):
    user, pos, neg = next(batches)
    Bq = user.numel()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]
    S, RU, dE = ss.S[0], ss.RU[0], ss.dE[0]
    ev[0].record()
    rid = parallel.run_ids(user); ev[1].record()
    call("fvx_bpr_step_sharded_a", C.byref(e.struct()), ptr(user), ptr(pos), ptr(neg), Bq, ptr(S), stream_ptr()); ev[2].record()
    dist.all_reduce(S); ev[3].record()
    call("fvx_bpr_step_sharded_b1", C.byref(e.struct()), ptr(user), Bq, ptr(S), ptr(rid), ptr(RU), RU.shape[0], 0, stream_ptr()); ev[4].record()
    dist.all_reduce(RU); ev[5].record()
    call("fvx_bpr_step_sharded_b2", C.byref(e.struct()), Bq, ptr(dE), stream_ptr()); ev[6].record()
    dist.all_reduce(dE); ev[7].record()
    call("fvx_bpr_step_sharded_c", C.byref(e.struct()), ptr(user), Bq, ptr(rid), ptr(RU), RU.shape[0], ptr(dE), 0 if rank == 0 else -1, stream_ptr()); ev[8].record()
    torch.cuda.synchronize()
    for i, n in enumerate(names):
        acc[n] += ev[i].elapsed_time(ev[i + 1]) / N
if rank == 0:
    print("world", world, "global batch", B, "RU rows", ss.RU[0].shape, "phases ms:", {k: round(v, 4) for k, v in acc.items()},
          "sum", round(sum(acc.values()), 4))
dist.destroy_process_group()
