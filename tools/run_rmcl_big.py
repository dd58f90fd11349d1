"""rMCL at BASELINE scale (configs C4 / C5), one process per GPU (torchrun for N > 1):
   python tools/run_rmcl_big.py planted 4000000 1000 [--iters 60 --eps 1e-6] [--parity 3]
   python tools/run_rmcl_big.py rmat 22 32 --iters 5
Runs b200_rmcl_iter_sharded (bounded arena: every step in row tiles), prints per-iteration time,
products, nnz and row tiles, iterations / second, and for planted graphs how well the clusters
match the planted blocks.  --parity S: before the timed loop, replays the first S iterations
step by step and checks a seeded sample of row blocks of EVERY one of them against the checker
(oracle/oracle.c through tests/oracle_lib.py; rank 0 only): structure exact, values <= 1e-12."""
import argparse, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import sparse_matrix_with_flops_b200 as smf

ap = argparse.ArgumentParser()
ap.add_argument("kind"); ap.add_argument("size", type=int); ap.add_argument("param", type=int)
ap.add_argument("--iters", type=int, default=5); ap.add_argument("--eps", type=float, default=0.0)
ap.add_argument("--parity", type=int, default=0); ap.add_argument("--blocks", type=int, default=3)
ap.add_argument("--block-rows", type=int, default=400)
args = ap.parse_args()
rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
if world > 1:
    import torch, torch.distributed as dist
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
smf.init(local)
if world > 1:
    uid = [smf.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    smf.comm_init(rank, world, uid[0])
t0 = time.perf_counter()
labels = None
if args.kind == "planted":
    A, labels = smf.synth_planted(args.size, args.param, 16, 2, 12345, want_labels=True)
    desc = "planted partition, %d vertices, %d blocks" % (args.size, args.param)
else:
    A = smf.synth_rmat(args.size, args.param, 12345, True)
    desc = "R-MAT scale %d edge factor %d symmetrised" % (args.size, args.param)
if rank == 0:
    print("%s: n %d nnz %d (generated in %.1f s)" % (desc, A.rows, A.nnz, time.perf_counter() - t0), flush=True)


def row_block(M, lo, hi):
    s, e = int(M.rowPtr[lo]), int(M.rowPtr[hi])
    return smf.CSR(M.values[s:e].copy(), M.colInd[s:e].copy(), (M.rowPtr[lo:hi + 1] - M.rowPtr[lo]).astype(np.int32),
                   hi - lo, M.cols)


if args.parity and rank == 0:
    import oracle_lib as ol
    rng = np.random.default_rng(7)
    dG = A.toGpuCSR()
    dT = A.toGpuCSR()
    Mt = A
    for k in range(args.parity):
        dN, ch, st = smf.gpuRmclOneStep(dG, dT, want_stats=True)
        los = rng.integers(0, A.rows - args.block_rows, args.blocks)
        worst = 0.0
        for lo in los:
            lo = int(lo); hi = lo + args.block_rows
            want = ol.o_make_ordered(ol.o_rmcl_onestep(ol.from_csr(row_block(A, lo, hi)), ol.from_csr(Mt)))
            got = dN.toCpuCSR(lo, hi).makeOrdered()
            ol.assert_same(ol.from_csr(got), want, 1e-12, "iteration %d rows [%d,%d)" % (k, lo, hi))
        print("parity iteration %d: %d blocks of %d rows equal the checker (tiles %d, products %d, unpruned %d, kept %d)" % (
            k, args.blocks, args.block_rows, st["row_tiles"], st["products"], st["nnz_unpruned"], st["nnz_out"]), flush=True)
        dT.deviceDispose()
        dT = dN
        if k + 1 < args.parity:
            Mt = dT.toCpuCSR()
    dT.deviceDispose(); dG.deviceDispose()

dG, dT = A.toGpuCSR(), A.toGpuCSR()
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
done, hist, ms, counts = smf.gpuRmclIterSharded(args.iters, dG, dT, eps=args.eps, want_counts=True)
if world > 1:
    dist.barrier()
sec = time.perf_counter() - t0
if rank == 0:
    print("ranks %d: %d iterations in %.3f s = %.3f iter/s (device time %.3f s)" % (world, done, sec, done / sec, ms.sum() / 1e3))
    for k in range(done):
        print("  iter %2d  %9.2f ms  products %.4g  nnz %.4g  unpruned %.4g  tiles %d  chaos %.6g" % (
            k, ms[k], counts[k, 0], counts[k, 1], counts[k, 2], counts[k, 3], hist[k]))
    if labels is not None:
        lab = dT.row_argmax()
        # purity: every found cluster votes for its most common planted block
        order = np.argsort(lab, kind="stable")
        ls, bs = lab[order], labels[order]
        starts = np.flatnonzero(np.concatenate([[True], ls[1:] != ls[:-1]]))
        ends = np.concatenate([starts[1:], [len(ls)]])
        hit = sum(int(np.bincount(bs[s:e]).max()) for s, e in zip(starts, ends))
        print("clusters found %d (planted %d); purity %.4f" % (len(starts), args.param, hit / len(ls)))
    sys.stdout.flush()
if world > 1:
    smf.comm_destroy()
    dist.destroy_process_group()
