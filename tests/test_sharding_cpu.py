"""Host-side logic of the sharded (N > 1) path, on CPU: world_size-2 `gloo` processes cut the
rows with the product's arrayEqualPartition64, each computes ITS row block with the checker
through the row-block convention (IA + lo, m = hi - lo), the blocks are gathered
and must concatenate to the whole product.  Also the bench.py helpers that feed the CPU arm."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle_lib as ol

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as ol_
    import sparse_matrix_with_flops_b200 as smf
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    A = smf.synth_rmat(9, 8, 77, True)
    M = ol_.from_csr(A)
    prefix = ol_.o_flops_prefix(M, M)
    ends = smf.arrayEqualPartition64(prefix, world)          # product code (host arithmetic)
    assert np.array_equal(ends, ol_.o_equal_partition64(prefix, world))
    lo, hi = int(ends[rank]), int(ends[rank + 1])
    blk = ol_.M(M.I[lo:hi + 1], M.J, M.V, hi - lo, M.cols)   # IA + lo: absolute offsets
    # oracle_spgemm indexes JA / A through IA[i], exactly like the reference kernels
    C = ol_.o_spgemm(blk, M)
    ol_.o_make_ordered(C)
    gathered = [None] * world
    dist.all_gather_object(gathered, (lo, hi, C.I, C.J, C.V))
    if rank == 0:
        whole = ol_.o_make_ordered(ol_.o_spgemm(M, M))
        I = [0]
        J, V = [], []
        for (l, h, bi, bj, bv) in sorted(gathered, key=lambda t: t[0]):
            I.extend((bi[1:] + I[-1]).tolist())
            J.append(bj)
            V.append(bv)
        ok = (np.array_equal(np.array(I, dtype=np.int32), whole.I)
              and np.array_equal(np.concatenate(J), whole.J)
              and np.array_equal(np.concatenate(V).view(np.int64), whole.V.view(np.int64)))
        # flops balance: no block above 1.6x the mean unless it is a single row
        per = np.diff(prefix[ends])
        balanced = all(p <= 1.6 * per.mean() or (ends[k + 1] - ends[k]) == 1 for k, p in enumerate(per))
        json.dump({"ok": bool(ok), "balanced": bool(balanced)}, open(os.path.join(out_dir, "result.json"), "w"))
    dist.barrier()
    dist.destroy_process_group()


def test_row_blocks_across_two_gloo_ranks(tmp_path):
    import torch.multiprocessing as mp
    port = 29600 + (os.getpid() % 200)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    res = json.load(open(tmp_path / "result.json"))
    assert res["ok"], "row blocks computed on 2 ranks do not concatenate to the whole product"
    assert res["balanced"]


def test_bench_helpers(smf):
    sys.path.insert(0, ROOT)
    import bench
    A = smf.synth_rmat(10, 8, 5, False)
    M = ol.from_csr(A)
    assert np.array_equal(bench.host_flops_prefix(A), ol.o_flops_prefix(M, M))
    I, J, V, m = bench.row_sample(A, 4)
    assert 0 < m < A.rows and I[0] == 0 and I[-1] == len(J) == len(V)
    # every sampled row is a row of A, verbatim and in order
    ids = np.arange(A.rows, dtype=np.uint64)
    rows = np.nonzero(((ids * np.uint64(2654435761)) % np.uint64(1 << 32)) < np.uint64((1 << 32) // 4))[0]
    for k in (0, m // 2, m - 1):
        r = rows[k]
        assert np.array_equal(J[I[k]:I[k + 1]], A.colInd[A.rowPtr[r]:A.rowPtr[r + 1]])
    # algorithmic bytes of a numeric launch
    assert bench.kernel_bytes("num", 10, 1000, 100, 500) == 12 * 100 + 40 + 12 * 1000 + 800 + 6000 + 40
    # ... and of one fused rMCL iteration: the same with the PRUNED nnz on the output side
    assert bench.rmcl_iter_bytes(9, 100, 1000, 50) == 12 * 100 + 40 + 12 * 1000 + 800 + 600 + 40
    A2, desc, small = bench.make_rmcl_workload(smf, "rmcl-rmat9e4")
    assert A2.rows == 512 and small == "rmcl-rmat9e4" and "symmetrised" in desc
    assert bench.make_rmcl_workload(smf, "rmcl-planted400000c100")[2] == "rmcl-planted100000c25"


@pytest.mark.parametrize("workload,metric", [("rmat10", "spgemm_gflops"), ("rmcl-rmat10", "rmcl_iters_per_s"),
                                             ("rmcl-planted3000c10", "rmcl_iters_per_s")])
def test_reference_arm_prints_the_contract_line(workload, metric):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                          "--workload", workload, "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert len(out.stdout.strip().splitlines()) == 1, "exactly one line on stdout"
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step",
                "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["metric"] == metric and line["value"] > 0
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
