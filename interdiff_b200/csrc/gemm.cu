// GEMM front door: C[M,N] = epi(A[M,K] . W[N,K]^T)  -- nn.Linear layout, row-major, fp32 in/out.
//
// Two backends, both fp32-grade (the sampling chain is ill-conditioned, DESIGN.md "Conditioning":
// plain TF32/FP16/BF16 operand rounding diverges to O(1) over a 100-1000 step loop):
//   backend 0: fp32 SIMT register-tiled kernel (this file) -- debug / bisect path and fallback
//              for shapes the tensor-core kernel does not take.
//   backend 1: tcgen05 3xTF32 split-precision kernel (gemm_tcgen05.cu).
#include "common.cuh"

int idb_gemm_tcgen05(idb_handle* h, const float* A, int lda, const float* W, int ldw, const float* bias,
                     const float* res, int ldr, float* C, int ldc, int M, int N, int K, int epi, cudaStream_t st);
bool idb_gemm_tcgen05_supported(int M, int N, int K, int lda, int ldw, int ldc);

namespace {

constexpr int BK = 16;

template <int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
gemm_tn_simt(const float* __restrict__ A, int lda, const float* __restrict__ W, int ldw,
             const float* __restrict__ bias, const float* __restrict__ res, int ldr,
             float* __restrict__ C, int ldc, int M, int N, int K, int epi) {
    constexpr int NT = (BM / TM) * (BN / TN);
    constexpr int LA = BM * BK / 4 / NT;  // float4 loads per thread for the A tile
    constexpr int LB = BN * BK / 4 / NT;
    static_assert(LA >= 1 && LB >= 1, "tile too small for the thread count");
    __shared__ __align__(16) float As[2][BK][BM + 4];
    __shared__ __align__(16) float Bs[2][BK][BN + 4];

    const int tid = threadIdx.x;
    const int tx = tid % (BN / TN), ty = tid / (BN / TN);
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; i++)
#pragma unroll
        for (int j = 0; j < TN; j++) acc[i][j] = 0.f;

    float4 ra[LA], rb[LB];
    auto gload = [&](int k0) {
#pragma unroll
        for (int l = 0; l < LA; l++) {
            int f = tid + l * NT, r = f / (BK / 4), kq = f % (BK / 4);
            int gm = m0 + r, gk = k0 + kq * 4;
            ra[l] = (gm < M && gk < K) ? *reinterpret_cast<const float4*>(A + (size_t)gm * lda + gk)
                                       : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int l = 0; l < LB; l++) {
            int f = tid + l * NT, r = f / (BK / 4), kq = f % (BK / 4);
            int gn = n0 + r, gk = k0 + kq * 4;
            rb[l] = (gn < N && gk < K) ? *reinterpret_cast<const float4*>(W + (size_t)gn * ldw + gk)
                                       : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    auto sstore = [&](int buf) {
#pragma unroll
        for (int l = 0; l < LA; l++) {
            int f = tid + l * NT, r = f / (BK / 4), kq = f % (BK / 4);
            As[buf][kq * 4 + 0][r] = ra[l].x; As[buf][kq * 4 + 1][r] = ra[l].y;
            As[buf][kq * 4 + 2][r] = ra[l].z; As[buf][kq * 4 + 3][r] = ra[l].w;
        }
#pragma unroll
        for (int l = 0; l < LB; l++) {
            int f = tid + l * NT, r = f / (BK / 4), kq = f % (BK / 4);
            Bs[buf][kq * 4 + 0][r] = rb[l].x; Bs[buf][kq * 4 + 1][r] = rb[l].y;
            Bs[buf][kq * 4 + 2][r] = rb[l].z; Bs[buf][kq * 4 + 3][r] = rb[l].w;
        }
    };

    const int nk = (K + BK - 1) / BK;
    gload(0);
    sstore(0);
    __syncthreads();
    for (int kb = 0; kb < nk; kb++) {
        const int buf = kb & 1;
        if (kb + 1 < nk) gload((kb + 1) * BK);
#pragma unroll
        for (int k = 0; k < BK; k++) {
            float a[TM], b[TN];
#pragma unroll
            for (int i = 0; i < TM; i += 4) {
                float4 v = *reinterpret_cast<const float4*>(&As[buf][k][ty * TM + i]);
                a[i] = v.x; a[i + 1] = v.y; a[i + 2] = v.z; a[i + 3] = v.w;
            }
#pragma unroll
            for (int j = 0; j < TN; j += 4) {
                float4 v = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * TN + j]);
                b[j] = v.x; b[j + 1] = v.y; b[j + 2] = v.z; b[j + 3] = v.w;
            }
#pragma unroll
            for (int i = 0; i < TM; i++)
#pragma unroll
                for (int j = 0; j < TN; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (kb + 1 < nk) {
            sstore(buf ^ 1);
            __syncthreads();
        }
    }

#pragma unroll
    for (int i = 0; i < TM; i++) {
        const int gm = m0 + ty * TM + i;
        if (gm >= M) continue;
#pragma unroll
        for (int j = 0; j < TN; j++) {
            const int gn = n0 + tx * TN + j;
            if (gn >= N) continue;
            float v = acc[i][j];
            if (epi & EPI_BIAS) v += bias[gn];
            if (epi & EPI_GELU) v = gelu_erf(v);
            if (epi & EPI_SILU) v = silu(v);
            if (epi & EPI_RES) v += res[(size_t)gm * ldr + gn];
            C[(size_t)gm * ldc + gn] = v;
        }
    }
}

}  // namespace

int idb_gemm(idb_handle* h, const float* A, int lda, const float* W, int ldw, const float* bias,
             const float* res, int ldr, float* C, int ldc, int M, int N, int K, int epi, cudaStream_t st) {
    if (M <= 0 || N <= 0 || K <= 0) return IDB_OK;
    if ((K % 4) || (lda % 4) || (ldw % 4))
        return idb_fail(h, IDB_ERR_ARG, "idb_gemm: K, lda, ldw must be multiples of 4 (K=%d lda=%d ldw=%d)", K, lda, ldw);
    if (h->gemm_backend == 1 && idb_gemm_tcgen05_supported(M, N, K, lda, ldw, ldc))
        return idb_gemm_tcgen05(h, A, lda, W, ldw, bias, res, ldr, C, ldc, M, N, K, epi, st);
    // tile choice: fill >= ~1 wave of 148 SMs when possible
    long tiles_big = (long)((M + 127) / 128) * ((N + 63) / 64);
    if (tiles_big >= 140) {
        dim3 grid((N + 63) / 64, (M + 127) / 128);
        gemm_tn_simt<128, 64, 8, 4><<<grid, 256, 0, st>>>(A, lda, W, ldw, bias, res, ldr, C, ldc, M, N, K, epi);
    } else {
        dim3 grid((N + 63) / 64, (M + 63) / 64);
        gemm_tn_simt<64, 64, 4, 4><<<grid, 256, 0, st>>>(A, lda, W, ldw, bias, res, ldr, C, ldc, M, N, K, epi);
    }
    LAUNCH_CHECK(h);
    return IDB_OK;
}
