"""Multi-process multi-GPU check + timing of the library's exchange transports (run under torchrun, one rank per GPU):
  NCCL all-gather called from C (kgv_shard_allgather), peer-memory publish/wait over CUDA IPC mappings (kgv_shard_publish_bitmap),
  and the sharded replay (kgv_set_sharding) against the unsharded result.
usage: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/mgpu_check.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist
import rusty_kaspa_b200 as rk
from rusty_kaspa_b200 import Params, simgen, workload as W
from rusty_kaspa_b200.comm import ShardComm
from rusty_kaspa_b200.replay import DagReplayer, REPLAY_BLOCK_DTYPE

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
ctx = rk.GpuContext(lr)
stream = torch.cuda.Stream(device=dev)
ctx.use_stream(stream.cuda_stream)
say = lambda *a: print(f"[rank {rank}]", *a, flush=True) if rank == 0 else None
with torch.cuda.stream(stream):
    comm = ShardComm.from_torch_distributed(ctx, slice_capacity=1 << 22)
    n = 1 << 18
    pk, msg, sig, kind = W.schnorr_triples(1 << 14, seed=7 + rank, n_keys=1024, n_nonces=1024)
    pk, msg, sig, kind = W.tile_triples(pk, msg, sig, kind, n)
    d = [torch.from_numpy(a).to(dev) for a in (pk, msg, sig)]
    st = torch.empty(n, dtype=torch.uint8, device=dev)
    bm = torch.empty(n // 8, dtype=torch.uint8, device=dev)
    all_nccl = torch.zeros(world * n // 8, dtype=torch.uint8, device=dev)
    all_p2p = torch.zeros(world * n // 8, dtype=torch.uint8, device=dev)
    all_torch = torch.zeros(world * n // 8, dtype=torch.uint8, device=dev)
    ctx.verify_schnorr_batch(d[0], d[1], d[2], n=n, status=st)
    ctx.status_to_bitmap(st, n=n, bitmap=bm)
    dist.all_gather_into_tensor(all_torch, bm)
    comm.allgather(bm.data_ptr(), n // 8, all_nccl.data_ptr())
    e = comm.publish_bitmap(st.data_ptr(), n)
    comm.wait(e, n // 8, all_p2p.data_ptr())
    stream.synchronize()
    assert torch.equal(all_torch, all_nccl) and torch.equal(all_torch, all_p2p), "transports disagree"
    say("transports agree on", world, "ranks; valid bits", int(sum(bin(x).count("1") for x in all_p2p.cpu().numpy()[:4096])))
    # timing: exchange only (status already computed), 50 rounds each
    def timed(fn, reps=50):
        for _ in range(5):
            fn()
        stream.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream); stream.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    def f_torch():
        ctx.status_to_bitmap(st, n=n, bitmap=bm); dist.all_gather_into_tensor(all_torch, bm)
    def f_nccl():
        ctx.status_to_bitmap(st, n=n, bitmap=bm); comm.allgather(bm.data_ptr(), n // 8, all_nccl.data_ptr())
    def f_p2p():
        ep = comm.publish_bitmap(st.data_ptr(), n); comm.wait(ep, n // 8, all_p2p.data_ptr())
    say(f"exchange of {n//8} B/rank: torch NCCL {timed(f_torch)*1e3:.1f} us | C-ABI NCCL {timed(f_nccl)*1e3:.1f} us | peer stores {timed(f_p2p)*1e3:.1f} us")
    # verify + exchange per step (the bench step), 10 rounds
    def s_p2p():
        ctx.verify_schnorr_batch(d[0], d[1], d[2], n=n, status=st); ep = comm.publish_bitmap(st.data_ptr(), n); comm.wait(ep, n // 8, all_p2p.data_ptr())
    def s_nccl():
        ctx.verify_schnorr_batch(d[0], d[1], d[2], n=n, status=st); ctx.status_to_bitmap(st, n=n, bitmap=bm); comm.allgather(bm.data_ptr(), n // 8, all_nccl.data_ptr())
    def s_none():
        ctx.verify_schnorr_batch(d[0], d[1], d[2], n=n, status=st)
    say(f"verify({n}) alone {timed(s_none, 10):.3f} ms | + NCCL {timed(s_nccl, 10):.3f} ms | + peer {timed(s_p2p, 10):.3f} ms")
    # sharded replay
    g = simgen.FastDag(seed=33, n_keys=256, n_nonces=1024, coinbase_maturity=20, frac_invalid=0.02, coinbase_outputs=16)
    wins = []
    for _ in range(4):
        g.generate(256, 150); wins.append(g.take())
    prm = Params(coinbase_maturity=20, storage_mass_parameter=g.C)
    def replay(sharded):
        comm.shard_validation(sharded)
        r = DagReplayer(ctx, prm, 1 << 20)
        stream.synchronize(); dist.barrier()
        t0 = time.perf_counter(); res = []
        for b, first, pov in wins:
            arr = np.zeros(len(pov), dtype=REPLAY_BLOCK_DTYPE)
            arr["first_tx"], arr["n_txs"], arr["pov_daa_score"], arr["flags"] = first[:-1], np.diff(first), pov, 1
            res.append(r.replay_window(b, arr)["status"].copy())
        dt = time.perf_counter() - t0
        dig = r.us.digest(); r.close()
        return dt, res, dig
    t1, r1, d1 = replay(False)
    t2, r2, d2 = replay(True)
    assert d1 == d2 and all((a == b).all() for a, b in zip(r1, r2)), "sharded replay differs"
    ntx = sum(len(w[0].txs) for w in wins)
    say(f"replay of {ntx} txs: unsharded {t1*1e3:.1f} ms, sharded over {world} ranks {t2*1e3:.1f} ms (identical verdicts and digest)")
comm.close()
dist.barrier()
dist.destroy_process_group()
