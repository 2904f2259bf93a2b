// tcgen05 3xTF32 GEMM backend -- placeholder until the tensor-core kernel lands.
#include "common.cuh"
bool idb_gemm_tcgen05_supported(int, int, int, int, int, int) { return false; }
int idb_gemm_tcgen05(idb_handle* h, const float*, int, const float*, int, const float*, const float*, int, float*, int,
                     int, int, int, int, cudaStream_t) {
    return idb_fail(h, IDB_ERR_STATE, "tcgen05 backend not built");
}
