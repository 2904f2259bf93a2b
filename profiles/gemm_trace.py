"""Per-CTA clock64 timeline of the tcgen05 GEMM pipeline (debug aid)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from interdiff_b200.engine import Engine  # noqa: E402

eng = Engine("cuda:0")
names = ["entry", "setup_done", "tma_first_issued", "tma_all_issued", "conv_first_full", "conv_first_done", "mma_first_ready",
         "mma_all_issued", "conv_all_done", "epi_acc_ready", "epi_done", "final_sync", "dealloc"]
for (M, N, K) in ((1920, 256, 256), (1920, 1024, 256), (1920, 256, 1024)):
    A = torch.randn(M, K, device="cuda")
    W = torch.randn(N, K, device="cuda")
    out = torch.empty(M, N, device="cuda")
    for rep in range(3):
        tr = torch.zeros(4096, 16, dtype=torch.int64, device="cuda")
        eng._chk(eng.lib.idb_debug_gemm_trace(eng._h, C.c_void_p(A.data_ptr()), C.c_void_p(W.data_ptr()), C.c_void_p(out.data_ptr()), M, N, K,
                                              C.c_void_p(tr.data_ptr()), eng._stream()))
        torch.cuda.synchronize()
    t = tr.cpu()
    ncta = int((t[:, 0] != 0).sum())
    t = t[:ncta]
    print("GEMM %dx%dx%d: %d CTAs; cycles since CTA entry" % (M, N, K, ncta))
    for i, n in enumerate(names):
        d = (t[:, i] - t[:, 0]).float()
        print("   %-18s median %8.0f   max %8.0f" % (n, d.median().item(), d.max().item()))
    print("   spread of CTA entry times: %d cycles" % int(t[:, 0].max() - t[:, 0].min()))
