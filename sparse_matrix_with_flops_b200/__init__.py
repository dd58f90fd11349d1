"""B200-native drop-in for the SpGEMM / rMCL hot path of ankur-maximos/Sparse_Matrix_with_Flops.

The product is the C-ABI library ``libb200spgemm.so`` (include/b200_spgemm.h, sources under
``csrc/``); this package is the host-side mirror of the reference's container / driver
interface on top of it.  Importing the package does not need a GPU; the first call that
computes does, and fails loudly without one (there is no CPU fallback).
"""
from . import _lib
from .csr import (COO_DEDUP, COO_NORMALISE, COO_SELF_LOOPS, CSR, DeviceCSR, DevicePCSR, RMCL,
                  arrayEqualPartition64, comm_destroy, comm_init, cooToGpuCSR, rmclInitDevice,
                  comm_unique_id, cost_prefix, flops_prefix, gpuRmclIter, gpuRmclIterSharded, gpuRmclOneStep,
                  gpuSpMMWrapper, init, reload_options, rmclInit, set_topk, synth_planted, synth_rmat, synth_stencil27)

__all__ = ["COO_DEDUP", "COO_NORMALISE", "COO_SELF_LOOPS", "cooToGpuCSR", "rmclInitDevice",
           "CSR", "DeviceCSR", "DevicePCSR", "RMCL", "arrayEqualPartition64", "comm_destroy", "comm_init",
           "comm_unique_id", "cost_prefix", "flops_prefix", "gpuRmclIter", "gpuRmclIterSharded",
           "gpuRmclOneStep", "gpuSpMMWrapper", "init", "reload_options", "set_topk", "rmclInit", "synth_planted", "synth_rmat",
           "synth_stencil27", "_lib"]
