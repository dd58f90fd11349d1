"""Multi-GPU rMCL check and timing (run under torchrun, one rank per GPU):
   python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/run_sharded_rmcl.py [kind size iters]
Every rank builds the same synthetic graph, runs b200_rmcl_iter_sharded over NCCL and rank 0
compares the final Mt with the checker (small sizes) and prints iterations / second."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch, torch.distributed as dist
import sparse_matrix_with_flops_b200 as smf

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
kind = sys.argv[1] if len(sys.argv) > 1 else "planted"
size = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 6
torch.cuda.set_device(local)
dist.init_process_group("gloo")          # only to ship the NCCL id and for barriers
smf.init(local)
uid = [smf.comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(uid, src=0)
smf.comm_init(rank, world, uid[0])
if kind == "planted": A = smf.synth_planted(size, max(1, size // 1000), 16, 2, 12345)
elif kind == "rmat": A = smf.synth_rmat(size, 16, 12345, True)
else: A = smf.synth_stencil27(size, size, size)
dG, dT = A.toGpuCSR(), A.toGpuCSR()
dist.barrier()
t0 = time.perf_counter()
done, hist, ms = smf.gpuRmclIterSharded(iters, dG, dT)
dist.barrier()
sec = time.perf_counter() - t0
Mt = dT.toCpuCSR()
if rank == 0:
    print("ranks %d %s %d: %d iterations in %.3f s = %.2f iter/s; per-iteration ms %s; final nnz %d chaos %.6g" % (
        world, kind, size, done, sec, done / sec, np.round(ms, 2).tolist(), Mt.nnz, hist[-1]), flush=True)
    if A.rows <= 200000:
        import oracle_lib as ol
        want, _, hw = ol.o_rmcl_iter(ol.from_csr(A), ol.from_csr(A), iters)
        ol.o_make_ordered(want)
        ol.assert_same(ol.from_csr(Mt), want, 1e-12, "sharded rMCL vs checker")
        assert np.allclose(hist, hw, rtol=0, atol=1e-12)
        print("parity with the checker: OK (structure exact, values <= 1e-12 rel, chaos history equal)", flush=True)
smf.comm_destroy()
dist.destroy_process_group()
