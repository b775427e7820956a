#!/usr/bin/env python
"""Parity of the sharded paths over REAL NCCL (one process per GPU, launched by torchrun):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        scripts/multi_parity.py

* fvx_bpr_step_sharded (one C call per rank per step, NCCL inside) against the fp64 oracle: per-step loss
  <= 1e-4 relative on every rank, parameters after the steps (user rows gathered from their owners);
* item-sharded top-k (per-shard sweep + all-to-all + merge) and user-sliced top-k against the oracle's masked top-k.
Exit code 0 = parity green on every rank.  tests/test_gpu_multi.py runs it when the box has >= 2 GPUs."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fvx import parallel                                           # noqa: E402
from oracle import bpr, evaluator as oe                             # noqa: E402

REL = 1e-4


def main():
    world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    grp = parallel.DistGroup()
    ok = True
    for (U, I, K, d, D, B, mode, tc) in ((900, 1201, 64, 20, 256, 1024, "deferred", True),
                                         (500, 701, 16, 0, 0, 256, "dense", False),
                                         (700, 900, 32, 20, 128, 600, "deferred", False)):
        rng = np.random.default_rng(U + K)                              # the same problem on every rank
        P = bpr.init_params(U, I, K, d, D, seed=3)
        P["Bi"] = (0.1 * rng.standard_normal(I)).astype(np.float32)
        F = bpr.normalise_features(np.maximum(rng.standard_normal((I, D)), 0)) if D else None
        e = parallel.sharded_engine(world, rank, U, I, K, d=d, D=D, lr=1e-3, reg=1e-3, adam_mode=mode, max_batch=B,
                                    device=str(dev), use_tensor_cores=tc)
        if D:
            e.set_features(F[e.item_lo:e.item_lo + e.Ic])
        e.load_params(P)
        step = parallel.ShardedStep([e], grp, max_runs=B // 6 + 3)
        Q = {k: v.astype(np.float64) for k, v in P.items()}
        S = bpr.init_adam(Q)
        F64 = F.astype(np.float64) if D else None
        for s in range(10):
            order = rng.permutation(U)[:B // 6 + 1]
            if s % 3 == 1:
                order[-1] = order[0]
            u = np.repeat(order, 6)[:B]
            b = (u, rng.integers(0, I, B), rng.integers(0, I, B))
            want = bpr.train_step(Q, S, b, 1e-3, 1e-3, F64)
            step.step(*(torch.as_tensor(x, dtype=torch.int32).to(dev) for x in b), loss_slot=s % 3)
            got = step.read_loss(s % 3)
            if not abs(got - want) <= REL * abs(want):
                ok = False
                print("rank %d: loss mismatch at step %d: %r vs %r" % (rank, s, got, want), flush=True)
        step.sync_users()
        R_ = e.params()
        for name, ref in Q.items():
            got = R_[name]
            if name in ("Gi", "Bi"):
                ref = ref[e.item_lo:e.item_lo + e.Ic]
            dlt = np.abs(got.reshape(ref.shape) - ref) / np.abs(ref).max()
            if not ((dlt > REL).mean() <= 2e-3 and dlt.max() <= 5e-3):
                ok = False
                print("rank %d: %s off: max %g, frac %g" % (rank, name, dlt.max(), (dlt > REL).mean()), flush=True)
        # evaluation: both decompositions against the oracle on the trained model
        k = 20
        tr = [sorted(rng.choice(I, int(rng.integers(1, 9)), replace=False).tolist()) for _ in range(U)]
        rp = torch.as_tensor(np.concatenate([[0], np.cumsum([len(t) for t in tr])]), dtype=torch.int64).to(dev)
        cs = torch.as_tensor(np.concatenate(tr), dtype=torch.int32).to(dev)
        Pfull = {}
        for name in Q:                                                # the model as the GPUs hold it, assembled
            t = torch.as_tensor(R_[name]).to(dev)
            if name in ("Gi", "Bi"):
                cnts = [parallel.shard_bounds(I, world, r)[1] for r in range(world)]
                pad = torch.zeros((max(cnts),) + tuple(t.shape[1:]), dtype=t.dtype, device=dev)
                pad[:t.shape[0]] = t
                out = [torch.zeros_like(pad) for _ in range(world)]
                dist.all_gather(out, pad)
                t = torch.cat([o[:c] for o, c in zip(out, cnts)])
            Pfull[name] = t.cpu().numpy()
        o_ids, o_sc = oe.masked_topk(bpr.predict_all(Pfull, F), tr, k)
        per, _ = parallel.user_slices(U, world)
        u0 = rank * per
        for fn in (parallel.sharded_topk, parallel.user_sliced_topk):
            ids, sc = fn([e], grp, rp, cs, k)[0]
            ids, sc = ids.cpu().numpy(), sc.cpu().numpy()
            for j in range(min(per, U - u0)):
                good, msg = oe.topk_matches(ids[j], sc[j], o_ids[u0 + j], o_sc[u0 + j])
                if not good:
                    ok = False
                    print("rank %d: %s user %d: %s" % (rank, fn.__name__, u0 + j, msg), flush=True)
                    break
    flag = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(flag)
    all_ok = float(flag.item()) == world
    if rank == 0:
        print("multi_parity: %s on %d ranks" % ("GREEN" if all_ok else "RED", world), flush=True)
    grp.close()
    dist.destroy_process_group()
    sys.exit(0 if all_ok else 1)


if __name__ == "__main__":
    main()
