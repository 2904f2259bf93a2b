// tcgen05 "3xTF32" split-precision GEMM for sm_100a:  C[M,N] = epi(A[M,K] . W[N,K]^T), fp32 in/out.
//
// Why split precision: the sampling chain amplifies operand rounding (DESIGN.md "Conditioning"),
// so the tensor-core path has to be fp32-grade.  Every fp32 operand x is split on chip into
//     big = tf32_rn(x),  small = tf32_rn(x - big)        (x - big is exact in fp32)
// and the product is accumulated in TMEM (fp32) as  a_s*w_b + a_b*w_s + a_b*w_b  (small terms
// first); the dropped a_s*w_s term is <= 2^-22 relative.
//
// Pipeline per CTA (one 128 x BN output tile, 320 threads):
//   warp 0   : TMA producer.  cp.async.bulk.tensor 2D loads of the raw fp32 A (128 x 32) and W
//              (BN x 32) k-blocks into 128B-swizzled shared memory, mbarrier complete_tx.
//   warps 2-9: converter.  Split the raw tiles in place (raw -> big) and write `small` copies at the
//              same swizzled offsets, fence.proxy.async, arrive on the stage's "ready" barrier.
//              After the k loop the same warps run the epilogue: tcgen05.ld (32x32b.x32) from TMEM,
//              bias / GELU / SiLU / residual, vectorised global stores.
//   warp 1   : MMA issuer (one elected lane).  12 tcgen05.mma.kind::tf32 per k-block (4 k-steps of 8
//              x 3 split terms), tcgen05.commit -> "empty" barrier (frees the stage) and finally
//              -> "accumulator full" barrier.  Also owns the TMEM allocation.
//
// Descriptor encodings follow cute/arch/mma_sm100_desc.hpp (SmemDescriptor / InstrDescriptor).
#include "common.cuh"

#include <cuda.h>

namespace {

constexpr int BM = 128;
constexpr int BK = 32;                 // 32 fp32 = one 128-byte swizzle row
constexpr int UMMA_K = 8;              // kind::tf32
constexpr int NUM_THREADS = 320;       // warp 0 TMA, warp 1 MMA, warps 2..9 converter + epilogue
constexpr int CONV_THREADS = 256;

template <int BN> struct Cfg {
    static constexpr int RAW_BYTES = (BM + BN) * BK * 4;     // A then W, both 1024-byte multiples
    static constexpr int STAGE_BYTES = 2 * RAW_BYTES;        // [raw->big | small]
    static constexpr int STAGES = (BN == 128) ? 3 : 4;
    // TMEM accumulators.  The tensor core truncates (rounds toward zero) on every fp32 accumulate,
    // so the error grows with the number of sequential accumulations into one accumulator
    // (measured: 2e-6 relative at K=256 with a single accumulator).  The big x big products are
    // therefore spread round-robin over NACC_MAIN accumulators and the two small cross terms get
    // their own one; the epilogue adds them up with IEEE round-to-nearest adds.  512 columns total.
    static constexpr int NACC_MAX = (BN == 128) ? 3 : 7;     // runtime `nacc` <= NACC_MAX
    static constexpr int TMEM_COLS = 512;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
};

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {}
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// round-to-nearest (ties away) fp32 -> tf32, i.e. cvt.rna.tf32.f32 without its inf/nan special-casing
// (2 integer ops; the PTX cvt expands to 4): finite inputs only, which is all a GEMM operand can be.
__device__ __forceinline__ uint32_t f32_to_tf32(float x) { return (__float_as_uint(x) + 0x1000u) & 0xFFFFE000u; }

// K-major, SWIZZLE_128B canonical layout: rows of 128 B, 8-row groups 1024 B apart.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);     // start address  [0,14)
    d |= (uint64_t)1 << 16;                       // leading byte offset (16 B, unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;             // stride byte offset: 8 rows x 128 B
    d |= (uint64_t)1 << 46;                       // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                       // SWIZZLE_128B
    return d;
}

template <int BN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_3xtf32_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
                   const float* __restrict__ bias, const float* __restrict__ res, int ldr,
                   float* __restrict__ C, int ldc, int M, int N, int K, int epi, int nacc, long long* __restrict__ trace) {
    using cfg = Cfg<BN>;
    // optional per-CTA timeline (clock64 at named points) for debugging the pipeline
    long long* tr = trace ? trace + (size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 16 : nullptr;
#define TRACE(slot) do { if (tr) tr[slot] = clock64(); } while (0)
    if (threadIdx.x == 0) TRACE(0);
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment for the 128B swizzle atoms
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t bars = base + cfg::STAGES * cfg::STAGE_BYTES;
    // barrier layout (8 B each): full[S], ready[S], empty[S], acc_full, then tmem slot
    auto bar_full = [&](int s) { return bars + 8u * s; };
    auto bar_ready = [&](int s) { return bars + 8u * (cfg::STAGES + s); };
    auto bar_empty = [&](int s) { return bars + 8u * (2 * cfg::STAGES + s); };
    const uint32_t bar_acc = bars + 8u * (3 * cfg::STAGES);
    const uint32_t tmem_slot = bar_acc + 8u;
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(base_ptr + cfg::STAGES * cfg::STAGE_BYTES + 8 * (3 * cfg::STAGES) + 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int num_kb = (K + BK - 1) / BK;

    if (threadIdx.x == 0) {
        for (int s = 0; s < cfg::STAGES; s++) {
            mbar_init(bar_full(s), 1);
            mbar_init(bar_ready(s), CONV_THREADS);
            mbar_init(bar_empty(s), 1);
        }
        mbar_init(bar_acc, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)cfg::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    if (threadIdx.x == 0) TRACE(1);

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            for (int kb = 0; kb < num_kb; kb++) {
                const int s = kb % cfg::STAGES, round = kb / cfg::STAGES;
                if (round > 0) mbar_wait(bar_empty(s), (round - 1) & 1);
                const uint32_t dst = base + s * cfg::STAGE_BYTES;
                mbar_arrive_expect_tx(bar_full(s), cfg::RAW_BYTES);
                tma_load_2d(dst, &map_a, bar_full(s), kb * BK, m0);
                tma_load_2d(dst + BM * BK * 4, &map_w, bar_full(s), kb * BK, n0);
                if (kb == 0) TRACE(2);
            }
            TRACE(3);
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            // instruction descriptor: D fp32, A/B tf32, both K-major, N>>3 @17, M>>4 @24
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
            for (int kb = 0; kb < num_kb; kb++) {
                const int s = kb % cfg::STAGES, round = kb / cfg::STAGES;
                mbar_wait(bar_ready(s), round & 1);
                tc_fence_after();
                if (kb == 0) TRACE(6);
                const uint32_t a_big = base + s * cfg::STAGE_BYTES, w_big = a_big + BM * BK * 4;
                const uint32_t a_sml = a_big + cfg::RAW_BYTES, w_sml = w_big + cfg::RAW_BYTES;
#pragma unroll
                for (int kk = 0; kk < BK / UMMA_K; kk++) {
                    const uint32_t koff = kk * UMMA_K * 4;   // 32 bytes per k-step inside the swizzle row
                    const uint64_t dab = make_smem_desc(a_big + koff), das = make_smem_desc(a_sml + koff);
                    const uint64_t dwb = make_smem_desc(w_big + koff), dws = make_smem_desc(w_sml + koff);
                    const uint32_t acc_main = tmem_base + (uint32_t)((kb % nacc) * BN);
                    const uint32_t acc_small = tmem_base + (uint32_t)(nacc * BN);
                    umma_tf32(acc_small, das, dwb, idesc, (kb | kk) ? 1u : 0u);
                    umma_tf32(acc_small, dab, dws, idesc, 1u);
                    umma_tf32(acc_main, dab, dwb, idesc, (kb >= nacc || kk > 0) ? 1u : 0u);
                }
                umma_commit(bar_empty(s));      // stage reusable once these MMAs have read it
            }
            umma_commit(bar_acc);               // accumulator complete
            TRACE(7);
        }
    } else {
        // ===================== converter, then epilogue (warps 2..5) =====================
        const int ct = threadIdx.x - 64;        // 0..127
        for (int kb = 0; kb < num_kb; kb++) {
            const int s = kb % cfg::STAGES, round = kb / cfg::STAGES;
            mbar_wait(bar_full(s), round & 1);
            if (kb == 0 && ct == 0) TRACE(4);
            uint8_t* st = base_ptr + s * cfg::STAGE_BYTES;
            constexpr int NCHUNK = cfg::RAW_BYTES / 16;
            static_assert(NCHUNK % (2 * CONV_THREADS) == 0, "converter loop assumes an even chunk count per thread");
#pragma unroll 2
            for (int c = ct; c < NCHUNK; c += 2 * CONV_THREADS) {
                const float4 x0 = *reinterpret_cast<float4*>(st + c * 16);
                const float4 x1 = *reinterpret_cast<float4*>(st + (c + CONV_THREADS) * 16);
                uint4 b0, s0, b1, s1;
                b0.x = f32_to_tf32(x0.x); b0.y = f32_to_tf32(x0.y); b0.z = f32_to_tf32(x0.z); b0.w = f32_to_tf32(x0.w);
                b1.x = f32_to_tf32(x1.x); b1.y = f32_to_tf32(x1.y); b1.z = f32_to_tf32(x1.z); b1.w = f32_to_tf32(x1.w);
                s0.x = f32_to_tf32(x0.x - __uint_as_float(b0.x)); s0.y = f32_to_tf32(x0.y - __uint_as_float(b0.y));
                s0.z = f32_to_tf32(x0.z - __uint_as_float(b0.z)); s0.w = f32_to_tf32(x0.w - __uint_as_float(b0.w));
                s1.x = f32_to_tf32(x1.x - __uint_as_float(b1.x)); s1.y = f32_to_tf32(x1.y - __uint_as_float(b1.y));
                s1.z = f32_to_tf32(x1.z - __uint_as_float(b1.z)); s1.w = f32_to_tf32(x1.w - __uint_as_float(b1.w));
                *reinterpret_cast<uint4*>(st + c * 16) = b0;
                *reinterpret_cast<uint4*>(st + (c + CONV_THREADS) * 16) = b1;
                *reinterpret_cast<uint4*>(st + cfg::RAW_BYTES + c * 16) = s0;
                *reinterpret_cast<uint4*>(st + cfg::RAW_BYTES + (c + CONV_THREADS) * 16) = s1;
            }
            fence_proxy_async();                // generic-proxy writes -> visible to the tensor core (async proxy)
            mbar_arrive(bar_ready(s));
            if (kb == 0 && ct == 0) TRACE(5);
        }
        if (ct == 0) TRACE(8);
        // ---- epilogue: 8 warps; warp w reads TMEM lane quarter (w & 3) and every other 32-column chunk.
        // TMEM -> registers (sum of the accumulators, bias, activation) -> padded smem tile (row per
        // lane) -> read back 4 rows x 128 B per instruction so the residual reads and the C stores
        // are fully coalesced.
        mbar_wait(bar_acc, 0);
        tc_fence_after();
        if (ct == 0) TRACE(9);
        const int q = warp & 3;                 // TMEM lane quarter this warp may access
        const int ew = warp - 2;                // 0..7
        constexpr int LDT = 36;                 // padded tile row (floats): conflict-free for both phases
        float* tile = reinterpret_cast<float*>(base_ptr) + ew * 32 * LDT;   // stage buffers are free now
        const int nused = (num_kb < nacc ? num_kb : nacc) + 1;              // used main accumulators + the small one
#pragma unroll 1
        for (int c0 = (ew >> 2) * 32; c0 < BN; c0 += 64) {
            float v[32];
#pragma unroll 1
            for (int a = 0; a < nused; a++) {
                const int acc = (a == nused - 1) ? nacc : a;                // small-term accumulator last
                uint32_t u[32];
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c0);
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                    : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]),
                      "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]),
                      "=r"(u[16]), "=r"(u[17]), "=r"(u[18]), "=r"(u[19]), "=r"(u[20]), "=r"(u[21]), "=r"(u[22]), "=r"(u[23]),
                      "=r"(u[24]), "=r"(u[25]), "=r"(u[26]), "=r"(u[27]), "=r"(u[28]), "=r"(u[29]), "=r"(u[30]), "=r"(u[31])
                    : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int j = 0; j < 32; j++) v[j] = (a == 0) ? __uint_as_float(u[j]) : v[j] + __uint_as_float(u[j]);
            }
            // phase 1: lane = tile row; bias / activation need only the column
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                const int col = n0 + c0 + j;
                float4 o = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                if (epi & EPI_BIAS) {
                    if (col + 3 < N) {
                        const float4 bb = *reinterpret_cast<const float4*>(bias + col);
                        o.x += bb.x; o.y += bb.y; o.z += bb.z; o.w += bb.w;
                    } else {
                        if (col < N) o.x += bias[col];
                        if (col + 1 < N) o.y += bias[col + 1];
                        if (col + 2 < N) o.z += bias[col + 2];
                    }
                }
                if (epi & EPI_GELU) { o.x = gelu_erf(o.x); o.y = gelu_erf(o.y); o.z = gelu_erf(o.z); o.w = gelu_erf(o.w); }
                if (epi & EPI_SILU) { o.x = silu(o.x); o.y = silu(o.y); o.z = silu(o.z); o.w = silu(o.w); }
                *reinterpret_cast<float4*>(tile + lane * LDT + j) = o;
            }
            __syncwarp();
            // phase 2: 8 lanes cover one 128-byte row segment; 4 rows per instruction
            const int cj = (lane & 7) * 4, col = n0 + c0 + cj;
#pragma unroll
            for (int rr = 0; rr < 32; rr += 4) {
                const int r = rr + (lane >> 3), row = m0 + q * 32 + r;
                float4 o = *reinterpret_cast<const float4*>(tile + r * LDT + cj);
                if (row < M) {
                    if (col + 3 < N) {
                        if (epi & EPI_RES) {
                            const float4 r4 = *reinterpret_cast<const float4*>(res + (size_t)row * ldr + col);
                            o.x += r4.x; o.y += r4.y; o.z += r4.z; o.w += r4.w;
                        }
                        *reinterpret_cast<float4*>(C + (size_t)row * ldc + col) = o;
                    } else {
                        const float oo[4] = {o.x, o.y, o.z, o.w};
                        for (int e = 0; e < 4; e++)
                            if (col + e < N) C[(size_t)row * ldc + col + e] = oo[e] + ((epi & EPI_RES) ? res[(size_t)row * ldr + col + e] : 0.f);
                    }
                }
            }
            __syncwarp();
        }
        tc_fence_before();
        if (ct == 0) TRACE(10);
    }
    __syncthreads();
    if (threadIdx.x == 0) TRACE(11);
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)cfg::TMEM_COLS) : "memory");
        if (lane == 0) TRACE(12);
    }
#undef TRACE
}

// ---------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;
bool g_attr_set = false;

int make_map(idb_handle* h, CUtensorMap* map, const float* ptr, int rows, int cols, int ld, int box_rows) {
    if (!g_encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        CUDA_TRY(h, cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (!fn || qres != cudaDriverEntryPointSuccess) return idb_fail(h, IDB_ERR_CUDA, "cuTensorMapEncodeTiled not available");
        g_encode = (EncodeTiledFn)fn;
    }
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)ptr, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return idb_fail(h, IDB_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return IDB_OK;
}

}  // namespace

bool idb_gemm_tcgen05_supported(int M, int N, int K, int lda, int ldw, int ldc) {
    // TMA needs 16-byte aligned row strides; the epilogue vector path needs ldc % 4 == 0
    return M >= 1 && N >= 8 && K >= 4 && (lda % 4 == 0) && (ldw % 4 == 0) && (ldc % 4 == 0) && (N % 4 == 0);
}

long long* g_idb_gemm_trace = nullptr;   // set by idb_debug_gemm_trace for ONE following launch
int g_idb_gemm_nacc = 0;                 // 0 = default; test hook (idb_debug_set_gemm_accumulators)

int idb_gemm_tcgen05(idb_handle* h, const float* A, int lda, const float* W, int ldw, const float* bias,
                     const float* res, int ldr, float* C, int ldc, int M, int N, int K, int epi, cudaStream_t st) {
    long long* trace = g_idb_gemm_trace;
    g_idb_gemm_trace = nullptr;
    if (((uintptr_t)A | (uintptr_t)W | (uintptr_t)C) & 15) return idb_fail(h, IDB_ERR_ARG, "tcgen05 GEMM needs 16-byte aligned pointers");
    if ((epi & EPI_RES) && ((ldr % 4) || ((uintptr_t)res & 15))) return idb_fail(h, IDB_ERR_ARG, "residual must be 16-byte aligned");
    if ((epi & EPI_BIAS) && ((uintptr_t)bias & 15)) return idb_fail(h, IDB_ERR_ARG, "bias must be 16-byte aligned");
    if (!g_attr_set) {
        CUDA_TRY(h, cudaFuncSetAttribute(gemm_3xtf32_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<128>::SMEM_BYTES));
        CUDA_TRY(h, cudaFuncSetAttribute(gemm_3xtf32_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<64>::SMEM_BYTES));
        g_attr_set = true;
    }
    // wide outputs take 128-column tiles; narrow ones 64 so more SMs get a tile
    const bool wide = (long)((M + BM - 1) / BM) * ((N + 127) / 128) >= 74;
    CUtensorMap ma, mw;
    int rc;
    if ((rc = make_map(h, &ma, A, M, K, lda, BM))) return rc;
    if ((rc = make_map(h, &mw, W, N, K, ldw, wide ? 128 : 64))) return rc;
    int nacc = wide ? Cfg<128>::NACC_MAX : Cfg<64>::NACC_MAX;
    if (g_idb_gemm_nacc > 0 && g_idb_gemm_nacc < nacc) nacc = g_idb_gemm_nacc;
    if (wide) {
        dim3 grid((N + 127) / 128, (M + BM - 1) / BM);
        gemm_3xtf32_kernel<128><<<grid, NUM_THREADS, Cfg<128>::SMEM_BYTES, st>>>(ma, mw, bias, res, ldr, C, ldc, M, N, K, epi, nacc, trace);
    } else {
        dim3 grid((N + 63) / 64, (M + BM - 1) / BM);
        gemm_3xtf32_kernel<64><<<grid, NUM_THREADS, Cfg<64>::SMEM_BYTES, st>>>(ma, mw, bias, res, ldr, C, ldc, M, N, K, epi, nacc, trace);
    }
    LAUNCH_CHECK(h);
    return IDB_OK;
}
