// tcgen05 split-precision GEMM for sm_100a:  C[M,N] = epi(A[M,K] . W[N,K]^T), fp32-grade accuracy.
//
// Why split precision: the sampling chain amplifies operand rounding (DESIGN.md "Conditioning"),
// so the tensor-core path has to be fp32-grade.  Every fp32 operand x is carried as a pair of fp16
//     hi = fp16_rn(x),   lo = fp16_rn((x - hi) * 2^11)          (x - hi is exact in fp32)
// so that x = hi + lo * 2^-11 up to 2^-23 |x| (the 2^11 scale keeps `lo` a normal fp16 number), and
//     A W^T = A_hi W_hi^T  +  2^-11 (A_lo W_hi^T + A_hi W_lo^T)  (+ 2^-22 A_lo W_lo^T, dropped).
// The two sums accumulate in SEPARATE fp32 TMEM accumulators and are combined in the epilogue.
// Compared with the first version of this kernel (3xTF32 on fp32 operands, split in shared memory)
// the fp16 pairs halve the L2->SM bytes, the shared-memory operand reads and the tensor time
// (profiles/README.md has the timelines that led here).  Range: |x| < 65504 (activations and weights of
// this model are O(1e-3 .. 1e3)).
//
// Pipeline per CTA (one 128 x BN output tile, 576 threads):
//   warp 0   : TMA producer (warp-uniform loop, one elected lane issues): 4 x cp.async.bulk.tensor.2d per
//              stage (A_hi, W_hi, A_lo, W_lo k-blocks of 64 fp16 = one 128B-swizzle row), mbarrier tx.
//   warp 1   : MMA issuer (warp-uniform loop, elected lane): 12 tcgen05.mma.kind::f16 per k-block
//              (4 k-steps of 16 x {lo*hi, hi*lo -> acc_small; hi*hi -> acc_main}), tcgen05.commit to the
//              stage's "empty" barrier and finally to the "accumulator full" barrier; owns TMEM.
//   warps 2-17: epilogue: tcgen05.ld 32x32b.x32 of the accumulators, sum, bias / GELU / SiLU, staged
//              through shared memory so that residual reads and C stores are coalesced; optional
//              output already split into (hi, lo) fp16 pairs for the next GEMM.
// The tensor core truncates on every fp32 accumulate, so the hi*hi products are spread round-robin over
// several TMEM accumulators (error vs #accumulators measured in profiles/README.md).
//
// Descriptor encodings follow cute/arch/mma_sm100_desc.hpp (SmemDescriptor / InstrDescriptor).
#include "common.cuh"
#include "tc_mlp.cuh"

#include <cuda.h>

namespace {

using namespace tc;

constexpr int GEMM_EW = 16;                          // epilogue warps (4 per TMEM lane quarter)
constexpr int GEMM_THREADS = (2 + GEMM_EW) * 32;     // + TMA producer + MMA issuer

template <int BN> struct Cfg {
    static constexpr int A_BYTES = BM * BK * 2;               // 16 KB
    static constexpr int W_BYTES = BN * BK * 2;
    static constexpr int HALF_BYTES = A_BYTES + W_BYTES;      // [A_hi | W_hi], then the same for lo
    static constexpr int STAGE_BYTES = 2 * HALF_BYTES;
    static constexpr int STAGES = (BN >= 192) ? 2 : (BN == 128) ? 3 : 4;
    static constexpr int NACC_MAX = (BN >= 192) ? 1 : (BN == 128) ? 3 : 7;   // runtime `nacc` <= NACC_MAX main accumulators (+1 small)
    static constexpr int TMEM_COLS = 512;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
};

// MC (multicast pairs): launched with cluster dimensions (1, 2, 1) - the two CTAs of a cluster own two ROW tiles of the same
// column tile.  Each loads its own A tile and HALF of the shared W tile, multicast into both CTAs' shared memory
// (cp.async.bulk.tensor ... .multicast::cluster), which halves the L2 -> SM traffic of the W operand; a stage is refilled only
// when BOTH CTAs' MMAs have released it (tcgen05.commit ... .multicast::cluster onto the "empty" barrier of both, count 2).
template <int BN, bool MC = false>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_split_f16_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
                      const __grid_constant__ CUtensorMap map_al, const __grid_constant__ CUtensorMap map_wl,
                      const float* __restrict__ bias, const float* __restrict__ res, int ldr,
                      float* __restrict__ C, __half* __restrict__ Ch, __half* __restrict__ Cl, int ldc, int M, int N, int K, int epi,
                      int nacc, float* __restrict__ zero_ptr, int zero_ld, int zero_cols, long long* __restrict__ trace) {
    using cfg = Cfg<BN>;
    // optional per-CTA timeline (clock64 at named points) for debugging the pipeline
    long long* tr = trace ? trace + (size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 16 : nullptr;
#define TRACE(slot) do { if (tr) tr[slot] = clock64(); } while (0)
    if (threadIdx.x == 0) { TRACE(0); chain_mark(1, 0); }
    pdl_trigger();   // the next kernel may start its own prologue now (it blocks in pdl_wait before touching our output)
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment for the 128B swizzle atoms
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t bars = base + cfg::STAGES * cfg::STAGE_BYTES;
    // barrier layout (8 B each): full[S], empty[S], acc_full, then the tmem slot
    auto bar_full = [&](int s) { return bars + 8u * s; };
    auto bar_empty = [&](int s) { return bars + 8u * (cfg::STAGES + s); };
    const uint32_t bar_acc = bars + 8u * (2 * cfg::STAGES);
    const uint32_t tmem_slot = bar_acc + 8u;
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(base_ptr + cfg::STAGES * cfg::STAGE_BYTES + 8 * (2 * cfg::STAGES) + 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    // split-K: blockIdx.z owns a contiguous range of k-blocks; the partial tiles are combined with
    // fp32 reductions onto a zeroed C (two addends only, so the result does not depend on their order)
    const int total_kb = (K + BK - 1) / BK, per_kb = (total_kb + gridDim.z - 1) / gridDim.z;
    const int kb0 = blockIdx.z * per_kb;
    const int num_kb = min(per_kb, total_kb - kb0);

    const int npre = num_kb < cfg::STAGES ? num_kb : cfg::STAGES;
    if (threadIdx.x == 0) {
        for (int s = 0; s < cfg::STAGES; s++) {
            mbar_init(bar_full(s), 1);
            mbar_init(bar_empty(s), MC ? 2 : 1);
        }
        mbar_init(bar_acc, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (!MC) {
            // The FIRST stage is requested right here, by the thread that initialised the barriers, before the TMEM
            // allocation and the CTA barrier: its weight tile (it does not depend on the previous kernel) at once, its activation
            // tile after the dependency wait; the producer warp requests the other stages after the barrier, while the MMA
            // warp already waits for this one.  A CTA that only got its SM when the previous kernel's CTAs left has nothing to
            // overlap its prologue with, so the sooner the first bytes are on their way - and the shorter the path to the
            // barrier - the better (gemm_trace.py: the barrier used to sit behind the requests of all four stages, 1.9 K cycles).
            mbar_arrive_expect_tx(bar_full(0), cfg::STAGE_BYTES);
            tma_load_2d(base + cfg::A_BYTES, &map_w, bar_full(0), kb0 * BK, n0);
            tma_load_2d(base + cfg::HALF_BYTES + cfg::A_BYTES, &map_wl, bar_full(0), kb0 * BK, n0);
            pdl_wait();
            chain_mark(1, 1);
            tma_load_2d(base, &map_a, bar_full(0), kb0 * BK, m0);
            tma_load_2d(base + cfg::HALF_BYTES, &map_al, bar_full(0), kb0 * BK, m0);
        } else {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&map_al) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&map_wl) : "memory");
        }
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
    uint32_t crank = 0;
    if (MC) {
        asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
        cluster_sync_all();      // the peer's barriers exist before anything is multicast into its shared memory
    }
    constexpr int WH = cfg::W_BYTES / 2;     // bytes of half a W tile (MC: the part this CTA fetches for both)

    if (warp == 0) {
        // ===================== TMA producer =====================
        // (multicast pairs only; otherwise the first fill was requested by thread 0 above)  The weight tiles of the first
        // pipeline fill do not depend on the previous kernel: requested before the dependency wait, the activation tiles after.
        if (!MC) {
            if (elect_one()) {
                for (int kb = 1; kb < npre; kb++) {      // the rest of the first pipeline fill (stage 0: thread 0 above)
                    const uint32_t dst = base + kb * cfg::STAGE_BYTES;
                    mbar_arrive_expect_tx(bar_full(kb), cfg::STAGE_BYTES);
                    tma_load_2d(dst + cfg::A_BYTES, &map_w, bar_full(kb), (kb0 + kb) * BK, n0);
                    tma_load_2d(dst + cfg::HALF_BYTES + cfg::A_BYTES, &map_wl, bar_full(kb), (kb0 + kb) * BK, n0);
                    tma_load_2d(dst, &map_a, bar_full(kb), (kb0 + kb) * BK, m0);
                    tma_load_2d(dst + cfg::HALF_BYTES, &map_al, bar_full(kb), (kb0 + kb) * BK, m0);
                }
            }
            __syncwarp();
        }
        if (MC) {
            if (elect_one()) {
                for (int kb = 0; kb < npre; kb++) {
                    const uint32_t dst = base + kb * cfg::STAGE_BYTES;
                    mbar_arrive_expect_tx(bar_full(kb), cfg::STAGE_BYTES);
                    tma_load_2d_mc(dst + cfg::A_BYTES + crank * WH, &map_w, bar_full(kb), (kb0 + kb) * BK, n0 + crank * (BN / 2), 3);
                    tma_load_2d_mc(dst + cfg::HALF_BYTES + cfg::A_BYTES + crank * WH, &map_wl, bar_full(kb), (kb0 + kb) * BK, n0 + crank * (BN / 2), 3);
                }
            }
            __syncwarp();
            pdl_wait();
            if (lane == 0) chain_mark(1, 1);
            if (elect_one()) {
                for (int kb = 0; kb < npre; kb++) {
                    const uint32_t dst = base + kb * cfg::STAGE_BYTES;
                    tma_load_2d(dst, &map_a, bar_full(kb), (kb0 + kb) * BK, m0);
                    tma_load_2d(dst + cfg::HALF_BYTES, &map_al, bar_full(kb), (kb0 + kb) * BK, m0);
                }
            }
            __syncwarp();
        }
        if (lane == 0) TRACE(2);
        for (int kb = npre; kb < num_kb; kb++) {
            const int s = kb % cfg::STAGES, round = kb / cfg::STAGES;
            mbar_wait(bar_empty(s), (round - 1) & 1);
            const uint32_t dst = base + s * cfg::STAGE_BYTES;
            if (elect_one()) {
                mbar_arrive_expect_tx(bar_full(s), cfg::STAGE_BYTES);
                tma_load_2d(dst, &map_a, bar_full(s), (kb0 + kb) * BK, m0);
                tma_load_2d(dst + cfg::HALF_BYTES, &map_al, bar_full(s), (kb0 + kb) * BK, m0);
                if (MC) {
                    tma_load_2d_mc(dst + cfg::A_BYTES + crank * WH, &map_w, bar_full(s), (kb0 + kb) * BK, n0 + crank * (BN / 2), 3);
                    tma_load_2d_mc(dst + cfg::HALF_BYTES + cfg::A_BYTES + crank * WH, &map_wl, bar_full(s), (kb0 + kb) * BK, n0 + crank * (BN / 2), 3);
                } else {
                    tma_load_2d(dst + cfg::A_BYTES, &map_w, bar_full(s), (kb0 + kb) * BK, n0);
                    tma_load_2d(dst + cfg::HALF_BYTES + cfg::A_BYTES, &map_wl, bar_full(s), (kb0 + kb) * BK, n0);
                }
            }
            __syncwarp();
        }
        if (lane == 0) TRACE(3);
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        // instruction descriptor: D fp32 (bit 4), A/B fp16 (format 0), both K-major, N>>3 @17, M>>4 @24
        const uint32_t idesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
        const uint32_t acc_small = tmem_base + (uint32_t)(nacc * BN);
        for (int kb = 0; kb < num_kb; kb++) {
            const int s = kb % cfg::STAGES, round = kb / cfg::STAGES;
            mbar_wait(bar_full(s), round & 1);
            tc_fence_after();
            if (kb == 0 && lane == 0) TRACE(6);
            const uint32_t a_hi = base + s * cfg::STAGE_BYTES, w_hi = a_hi + cfg::A_BYTES;
            const uint64_t dah = make_smem_desc(a_hi), dal = make_smem_desc(a_hi + cfg::HALF_BYTES);
            const uint64_t dwh = make_smem_desc(w_hi), dwl = make_smem_desc(w_hi + cfg::HALF_BYTES);
            const uint32_t acc_main = tmem_base + (uint32_t)((kb % nacc) * BN);
            const uint32_t first_main = kb >= nacc ? 1u : 0u, first_small = kb ? 1u : 0u;
            if (elect_one()) {
#pragma unroll
                for (int kk = 0; kk < BK / UMMA_K; kk++) {
                    const uint64_t koff = (uint64_t)((kk * UMMA_K * 2) >> 4);   // +32 bytes per k-step in the start-address field
                    if (!(epi & 64)) {   // debug bit 64: hi x hi only (MMA throughput probe)
                        umma_f16(acc_small, dal + koff, dwh + koff, idesc, kk ? 1u : first_small);
                        umma_f16(acc_small, dah + koff, dwl + koff, idesc, 1u);
                    }
                    umma_f16(acc_main, dah + koff, dwh + koff, idesc, kk ? 1u : first_main);
                }
                if (MC) umma_commit_mc(bar_empty(s), 3);      // ... in BOTH CTAs of the pair (each multicasts into the other's stage)
                else umma_commit(bar_empty(s));                // stage reusable once these MMAs have read it
                if (kb == num_kb - 1) umma_commit(bar_acc);    // accumulators complete
            }
            __syncwarp();
        }
        if (lane == 0) TRACE(7);
    } else {
        // ===================== epilogue (warps 2..17) =====================
        // warp w reads TMEM lane quarter (w & 3) and every fourth 32-column chunk.  TMEM -> registers (sum of
        // the accumulators, bias, activation) -> padded smem tile (row per lane) -> read back 4 rows x 128 B
        // per instruction so the residual reads and the C stores are fully coalesced.
        const int ct = threadIdx.x - 64;
        pdl_wait();      // bias / residual reads and the C stores below touch buffers of earlier kernels
        if (zero_ptr) {
            // side job while the MMAs run: clear this CTA's patch of an auxiliary matrix (the zeroed C that
            // the NEXT split-K GEMM reduces into)
            const int zc4 = zero_cols / (int)gridDim.x / 4;
            for (int i = ct; i < BM * zc4; i += GEMM_EW * 32) {
                const int r = i / zc4, c4 = i % zc4;
                if (m0 + r < M)
                    *reinterpret_cast<float4*>(zero_ptr + (size_t)(m0 + r) * zero_ld + (blockIdx.x * zc4 + c4) * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        const bool splitk = gridDim.z > 1, lead = blockIdx.z == 0;
        mbar_wait(bar_acc, 0);
        tc_fence_after();
        if (ct == 0) TRACE(9);
        const int q = warp & 3;                 // TMEM lane quarter this warp may access
        const int ew = warp - 2;                // 0..7
        constexpr int LDT = 36;                 // padded tile row (floats): conflict-free for both phases
        float* tile = reinterpret_cast<float*>(base_ptr) + ew * 32 * LDT;   // stage buffers are free now
        const int nmain = num_kb < nacc ? num_kb : nacc;
#pragma unroll 1
        for (int c0 = (ew >> 2) * 32; c0 < BN; c0 += 8 * GEMM_EW) {
            float v[32];
#pragma unroll 1
            for (int a = 0; a <= nmain; a++) {
                // a == nmain is the small-term accumulator: D = sum(main) + 2^-11 * small
                const int acc = (a == nmain) ? nacc : a;
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
                if (a == 0) {
#pragma unroll
                    for (int j = 0; j < 32; j++) v[j] = __uint_as_float(u[j]);
                } else if (a < nmain) {
#pragma unroll
                    for (int j = 0; j < 32; j++) v[j] += __uint_as_float(u[j]);
                } else {
#pragma unroll
                    for (int j = 0; j < 32; j++) v[j] = fmaf(__uint_as_float(u[j]), 1.0f / 2048.0f, v[j]);
                }
            }
            // phase 1: lane = tile row; raw sums into the padded tile
#pragma unroll
            for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(tile + lane * LDT + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            __syncwarp();
            // phase 2: 8 lanes cover one 128-byte row segment, 4 rows per instruction; bias (loop invariant
            // for the lane), activation, residual, stores
            const int cj = (lane & 7) * 4, col = n0 + c0 + cj;
            float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if ((epi & EPI_BIAS) && lead) {
                if (col + 3 < N) b4 = *reinterpret_cast<const float4*>(bias + col);
                else { if (col < N) b4.x = bias[col]; if (col + 1 < N) b4.y = bias[col + 1]; if (col + 2 < N) b4.z = bias[col + 2]; }
            }
#pragma unroll 2
            for (int rr = 0; rr < 32; rr += 4) {
                const int r = rr + (lane >> 3), row = m0 + q * 32 + r;
                float4 o = *reinterpret_cast<const float4*>(tile + r * LDT + cj);
                o.x += b4.x; o.y += b4.y; o.z += b4.z; o.w += b4.w;
                if (epi & EPI_GELU) { o.x = gelu_erf(o.x); o.y = gelu_erf(o.y); o.z = gelu_erf(o.z); o.w = gelu_erf(o.w); }
                if (epi & EPI_SILU) { o.x = silu(o.x); o.y = silu(o.y); o.z = silu(o.z); o.w = silu(o.w); }
                if (row < M) {
                    if (col + 3 < N) {
                        if ((epi & EPI_RES) && lead) {
                            const float4 r4 = *reinterpret_cast<const float4*>(res + (size_t)row * ldr + col);
                            o.x += r4.x; o.y += r4.y; o.z += r4.z; o.w += r4.w;
                        }
                        if (splitk) {
                            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(C + (size_t)row * ldc + col), "f"(o.x), "f"(o.y),
                                         "f"(o.z), "f"(o.w) : "memory");
                        } else if (C) *reinterpret_cast<float4*>(C + (size_t)row * ldc + col) = o;
                        if (Ch) {
                            __half2 h01, h23, l01, l23;
                            split_f16x2(o.x, o.y, h01, l01);
                            split_f16x2(o.z, o.w, h23, l23);
                            *reinterpret_cast<uint2*>(Ch + (size_t)row * ldc + col) = make_uint2(*reinterpret_cast<uint32_t*>(&h01), *reinterpret_cast<uint32_t*>(&h23));
                            *reinterpret_cast<uint2*>(Cl + (size_t)row * ldc + col) = make_uint2(*reinterpret_cast<uint32_t*>(&l01), *reinterpret_cast<uint32_t*>(&l23));
                        }
                    } else {
                        const float oo[4] = {o.x, o.y, o.z, o.w};
                        for (int e = 0; e < 4; e++)
                            if (col + e < N) {
                                const float val = oo[e] + (((epi & EPI_RES) && lead) ? res[(size_t)row * ldr + col + e] : 0.f);
                                if (splitk) atomicAdd(C + (size_t)row * ldc + col + e, val);
                                else if (C) C[(size_t)row * ldc + col + e] = val;
                                if (Ch) split_f16(val, Ch[(size_t)row * ldc + col + e], Cl[(size_t)row * ldc + col + e]);
                            }
                    }
                }
            }
            __syncwarp();
        }
        tc_fence_before();
        if (ct == 0) TRACE(10);
    }
    __syncthreads();
    if (MC) cluster_sync_all();      // the peer may still arrive on this CTA's barriers (multicast commits) until it is done too
    if (threadIdx.x == 0) { TRACE(11); chain_mark(1, 2); }
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)cfg::TMEM_COLS) : "memory");
        if (lane == 0) TRACE(12);
    }
#undef TRACE
}

// stand-alone launch of the feed-forward block: grid (8, ceil(M / 128)), one 128-row tile per cluster
constexpr int MLP_EW = 16;                          // epilogue warps of the stand-alone kernel (4 per TMEM lane quarter)
constexpr int MLP_THREADS = (2 + MLP_EW) * 32;      // + TMA producer + MMA issuer
__global__ void __cluster_dims__(mlp::CLUSTER, 1, 1) __launch_bounds__(MLP_THREADS, 1)
mlp_fused_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w1,
                 const __grid_constant__ CUtensorMap map_xl, const __grid_constant__ CUtensorMap map_w1l,
                 const __grid_constant__ CUtensorMap map_w2, const __grid_constant__ CUtensorMap map_w2l,
                 const float* __restrict__ b1, const float* __restrict__ b2, const float* __restrict__ res, int ldr,
                 float* __restrict__ Z, int ldz, int M, const float* __restrict__ ln_w, const float* __restrict__ ln_b,
                 __half* __restrict__ Zh, __half* __restrict__ Zl, long long* __restrict__ trace) {
    long long* tr = trace ? trace + (size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 16 : nullptr;
    if (tr && threadIdx.x == 0) tr[0] = clock64();
    if (threadIdx.x == 0) chain_mark(2, 0);
    pdl_trigger();
    extern __shared__ uint8_t smem_raw[];
    MlpArgs a{&map_x, &map_w1, &map_xl, &map_w1l, &map_w2, &map_w2l, b1, b2, res, ldr, Z, ldz, ln_w, ln_b, Zh, Zl, trace};
    const uint32_t tmem_base = mlp_setup<MLP_EW, true>(smem_raw, a, blockIdx.x, blockIdx.y * BM);
    mlp_run<MLP_EW, true>(smem_raw, a, blockIdx.x, blockIdx.y * BM, M, tmem_base, true, tr);
    if (threadIdx.x == 0) chain_mark(2, 2);
    mlp_teardown(tmem_base);
}

// ---------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;

int make_map(idb_handle* h, CUtensorMap* map, const __half* ptr, int rows, int cols, int ld, int box_rows) {
    if (!g_encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        CUDA_TRY(h, cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (!fn || qres != cudaDriverEntryPointSuccess) return idb_fail(h, IDB_ERR_CUDA, "cuTensorMapEncodeTiled not available");
        g_encode = (EncodeTiledFn)fn;
    }
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void*)ptr, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return idb_fail(h, IDB_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return IDB_OK;
}

}  // namespace

CHAIN_SETTER(idb_chain_set_gemm)
long long* g_idb_gemm_trace = nullptr;   // set by the debug hooks for ONE following launch
int g_idb_gemm_bn192 = 1;                // 192-column tiles where they fill more SMs than 256-column ones (probe switch)
int g_idb_gemm_nacc = 0;                 // 0 = default; test hook (idb_debug_set_gemm_accumulators)

bool idb_gemm_tcgen05_supported(const GemmArgs& g) {
    // TMA needs 16-byte aligned row strides (8 fp16); the vector epilogue needs ldc % 4 == 0
    return g.A_lo && g.W_lo && g.M >= 1 && g.N >= 8 && g.K >= 8 && (g.lda % 8 == 0) && (g.ldw % 8 == 0) && (g.ldc % 4 == 0) &&
           (g.N % 4 == 0);
}

int idb_gemm_tcgen05(idb_handle* h, const GemmArgs& g, cudaStream_t st) {
    long long* trace = g_idb_gemm_trace;
    g_idb_gemm_trace = nullptr;
    const int M = g.M, N = g.N, K = g.K;
    auto mis = [](const void* p) { return ((uintptr_t)p & 15) != 0; };
    if (mis(g.A_hi) || mis(g.W_hi) || mis(g.A_lo) || mis(g.W_lo) || mis(g.C) || mis(g.C_hi) || mis(g.C_lo))
        return idb_fail(h, IDB_ERR_ARG, "tcgen05 GEMM needs 16-byte aligned pointers");
    if ((g.epi & EPI_RES) && ((g.ldr % 4) || mis(g.res))) return idb_fail(h, IDB_ERR_ARG, "residual must be 16-byte aligned");
    if ((g.epi & EPI_BIAS) && mis(g.bias)) return idb_fail(h, IDB_ERR_ARG, "bias must be 16-byte aligned");
    if (!(h->attr_mask & 1u)) {      // per handle = per device (function attributes live in the device's context)
        CUDA_TRY(h, cudaFuncSetAttribute(gemm_split_f16_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<128>::SMEM_BYTES));
        CUDA_TRY(h, cudaFuncSetAttribute(gemm_split_f16_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<64>::SMEM_BYTES));
        CUDA_TRY(h, cudaFuncSetAttribute(gemm_split_f16_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<256>::SMEM_BYTES));
        CUDA_TRY(h, cudaFuncSetAttribute(gemm_split_f16_kernel<192>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<192>::SMEM_BYTES));
        h->attr_mask |= 1u;
    }
    // wide outputs take 128-column tiles; narrow ones 64 so more SMs get a tile
    const long tiles128 = (long)((M + BM - 1) / BM) * ((N + 127) / 128);
    const bool wide = tiles128 >= 74;
    // more 128-column tiles than SMs would run as two waves: 256-column tiles (one main accumulator, so only for
    // short reductions) put e.g. the folded QKV projection (N = 1536) on 90 CTAs in a single wave
    const bool xwide = tiles128 > h->sm_count && (N % 256) == 0 && ((K + BK - 1) / BK <= 4 || g.single_acc) && g.ksplit != 2 && !g.zero &&
                       (g_idb_gemm_nacc <= 0 || g_idb_gemm_nacc == 1);
    // ... and 192-column tiles put it on 120: same single accumulator and k order per element (bit-identical), 25 % less work per CTA
    const bool x192 = xwide && g_idb_gemm_bn192 && !g.single_acc && (N % 192) == 0 && (long)((M + BM - 1) / BM) * (N / 192) <= h->sm_count;
    const int bn = x192 ? 192 : xwide ? 256 : wide ? 128 : 64;
    CUtensorMap ma, mw, mal, mwl;
    int rc;
    if ((rc = make_map(h, &ma, g.A_hi, M, K, g.lda, BM))) return rc;
    if ((rc = make_map(h, &mw, g.W_hi, N, K, g.ldw, bn))) return rc;
    if ((rc = make_map(h, &mal, g.A_lo, M, K, g.lda, BM))) return rc;
    if ((rc = make_map(h, &mwl, g.W_lo, N, K, g.ldw, bn))) return rc;
    // split-K (2 only: two addends commute, so the reduction order cannot change the result) needs a C that
    // an earlier kernel zeroed (GemmArgs::zero of the producer GEMM) and the plain fp32 output
    int ksplit = g.ksplit == 2 && (K + BK - 1) / BK >= 2 ? 2 : 1;
    if (ksplit == 2 && (!g.C || g.C_hi || (g.epi & (EPI_GELU | EPI_SILU))))
        return idb_fail(h, IDB_ERR_ARG, "split-K GEMM: fp32 output only, no activation");
    float* zero = g.zero;
    if (zero) {
        const int gx = wide ? (N + 127) / 128 : (N + 63) / 64;
        if (g.zero_cols % (4 * gx) || (g.zero_ld & 3)) return idb_fail(h, IDB_ERR_ARG, "aux zero: %d columns do not split over %d column tiles", g.zero_cols, gx);
    }
    // Round-robin main accumulators bound the tensor core's truncate-on-accumulate error; every one costs an
    // extra TMEM read pass in the epilogue.  Two already beat the fp32 SIMT kernel up to 8 k-blocks per CTA
    // (measured, profiles/README.md: 3.3e-7 at K=256, 4.1e-7 at K=1024 split in two, vs 5.5e-7 / 1.5e-6); longer
    // reductions get one accumulator per 4 k-blocks up to the TMEM budget.
    // The accumulator count is a function of the GEMM's (N, K) only - never of M - so that a sample's result does not
    // depend on how many samples share the batch (multi-GPU runs slice one global batch).  Shapes that the
    // 256-column configuration serves at large M (one main accumulator) use one accumulator at every M.
    const int nacc_max = wide ? Cfg<128>::NACC_MAX : Cfg<64>::NACC_MAX;
    const int kb_per_cta = ((K + BK - 1) / BK + ksplit - 1) / ksplit;
    const bool one_acc = g.single_acc || ((N % 256) == 0 && (K + BK - 1) / BK <= 4 && g.ksplit != 2 && !g.zero);
    int nacc = one_acc ? 1 : kb_per_cta <= 8 ? 2 : (kb_per_cta + 3) / 4;
    if (nacc > nacc_max) nacc = nacc_max;
    if (g_idb_gemm_nacc > 0) nacc = g_idb_gemm_nacc < nacc_max ? g_idb_gemm_nacc : nacc_max;
    if (xwide && g.single_acc && h->gemm_multicast && (M + BM - 1) / BM >= 2) {
        // long, wide GEMMs (the SMPL-H blend): pairs of row tiles share the W tile by TMA multicast (cluster 1 x 2 x 1)
        CUtensorMap mwh, mwlh;
        if ((rc = make_map(h, &mwh, g.W_hi, N, K, g.ldw, 128))) return rc;
        if ((rc = make_map(h, &mwlh, g.W_lo, N, K, g.ldw, 128))) return rc;
        if (!(h->attr_mask & 8u)) {
            CUDA_TRY(h, cudaFuncSetAttribute(gemm_split_f16_kernel<256, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<256>::SMEM_BYTES));
            h->attr_mask |= 8u;
        }
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(N / 256, (((M + BM - 1) / BM) + 1) & ~1, 1);      // an even number of row tiles (a tile past M loads zeros, stores nothing)
        cfg.blockDim = dim3(GEMM_THREADS); cfg.dynamicSmemBytes = Cfg<256>::SMEM_BYTES; cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 1; at[0].val.clusterDim.y = 2; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        CUDA_TRY(h, cudaLaunchKernelEx(&cfg, gemm_split_f16_kernel<256, true>, ma, mwh, mal, mwlh, g.bias, g.res, g.ldr, g.C, g.C_hi, g.C_lo, g.ldc,
                                       M, N, K, g.epi, 1, (float*)nullptr, 0, 0, trace));
    } else if (x192) {
        dim3 grid(N / 192, (M + BM - 1) / BM, 1);
        idb_launch(g.pdl != 0, gemm_split_f16_kernel<192>, grid, GEMM_THREADS, Cfg<192>::SMEM_BYTES, st, ma, mw, mal, mwl, g.bias, g.res, g.ldr,
                   g.C, g.C_hi, g.C_lo, g.ldc, M, N, K, g.epi, 1, nullptr, 0, 0, trace);
    } else if (xwide) {
        dim3 grid(N / 256, (M + BM - 1) / BM, 1);
        idb_launch(g.pdl != 0, gemm_split_f16_kernel<256>, grid, GEMM_THREADS, Cfg<256>::SMEM_BYTES, st, ma, mw, mal, mwl, g.bias, g.res, g.ldr,
                   g.C, g.C_hi, g.C_lo, g.ldc, M, N, K, g.epi, 1, nullptr, 0, 0, trace);
    } else if (wide) {
        dim3 grid((N + 127) / 128, (M + BM - 1) / BM, ksplit);
        idb_launch(g.pdl != 0, gemm_split_f16_kernel<128>, grid, GEMM_THREADS, Cfg<128>::SMEM_BYTES, st, ma, mw, mal, mwl, g.bias, g.res, g.ldr,
                   g.C, g.C_hi, g.C_lo, g.ldc, M, N, K, g.epi, nacc, zero, g.zero_ld, g.zero_cols, trace);
    } else {
        dim3 grid((N + 63) / 64, (M + BM - 1) / BM, ksplit);
        idb_launch(g.pdl != 0, gemm_split_f16_kernel<64>, grid, GEMM_THREADS, Cfg<64>::SMEM_BYTES, st, ma, mw, mal, mwl, g.bias, g.res, g.ldr,
                   g.C, g.C_hi, g.C_lo, g.ldc, M, N, K, g.epi, nacc, zero, g.zero_ld, g.zero_cols, trace);
    }
    LAUNCH_CHECK(h);
    return IDB_OK;
}

// Fused feed-forward block (see mlp_fused_kernel).  x / w1 / w2 as fp16 (hi, lo) pairs, row-major:
// x [M][256], w1 [F][256], w2 [256][F];  Z[M][256] = gelu(x w1^T + b1) w2^T + b2 + res, optionally followed by the layer's
// final LayerNorm (ln_w / ln_b) and an additional fp16 (hi, lo) copy of the output (Z_hi / Z_lo) for the next GEMM.
bool idb_mlp_tcgen05_supported(int d_model, int d_ff) { return d_model == mlp::DM && d_ff == mlp::FC * mlp::CLUSTER; }

// order: x_hi, w1_hi, x_lo, w1_lo, w2_hi, w2_lo (the kernels' parameter order)
int idb_mlp_make_maps(idb_handle* h, const __half* x_hi, const __half* x_lo, const __half* w1_hi, const __half* w1_lo, const __half* w2_hi,
                      const __half* w2_lo, int M, CUtensorMap* out6) {
    const int F = mlp::FC * mlp::CLUSTER;
    int rc;
    if ((rc = make_map(h, &out6[0], x_hi, M, mlp::DM, mlp::DM, BM))) return rc;
    if ((rc = make_map(h, &out6[1], w1_hi, F, mlp::DM, mlp::DM, mlp::FC))) return rc;
    if ((rc = make_map(h, &out6[2], x_lo, M, mlp::DM, mlp::DM, BM))) return rc;
    if ((rc = make_map(h, &out6[3], w1_lo, F, mlp::DM, mlp::DM, mlp::FC))) return rc;
    // W2: boxes of 128 rows (one CTA's half of the output columns) x 64 hidden columns
    if ((rc = make_map(h, &out6[4], w2_hi, mlp::DM, F, F, mlp::NH))) return rc;
    return make_map(h, &out6[5], w2_lo, mlp::DM, F, F, mlp::NH);
}

int idb_mlp_tcgen05(idb_handle* h, const __half* x_hi, const __half* x_lo, const __half* w1_hi, const __half* w1_lo, const float* b1,
                    const __half* w2_hi, const __half* w2_lo, const float* b2, const float* res, int ldr, float* Z, int ldz, int M,
                    int pdl, cudaStream_t st, const float* ln_w, const float* ln_b, __half* Z_hi, __half* Z_lo) {
    long long* trace = g_idb_gemm_trace;
    g_idb_gemm_trace = nullptr;
    if (!(h->attr_mask & 2u)) {
        CUDA_TRY(h, cudaFuncSetAttribute(mlp_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, mlp::SMEM_BYTES));
        h->attr_mask |= 2u;
    }
    const int F = mlp::FC * mlp::CLUSTER;
    CUtensorMap mx, mxl, mw1, mw1l, mw2, mw2l;
    int rc;
    if ((rc = make_map(h, &mx, x_hi, M, mlp::DM, mlp::DM, BM))) return rc;
    if ((rc = make_map(h, &mxl, x_lo, M, mlp::DM, mlp::DM, BM))) return rc;
    if ((rc = make_map(h, &mw1, w1_hi, F, mlp::DM, mlp::DM, mlp::FC))) return rc;
    if ((rc = make_map(h, &mw1l, w1_lo, F, mlp::DM, mlp::DM, mlp::FC))) return rc;
    if ((rc = make_map(h, &mw2, w2_hi, mlp::DM, F, F, mlp::NH))) return rc;
    if ((rc = make_map(h, &mw2l, w2_lo, mlp::DM, F, F, mlp::NH))) return rc;
    dim3 grid(mlp::CLUSTER, (M + BM - 1) / BM);
    idb_launch(pdl != 0, mlp_fused_kernel, grid, MLP_THREADS, mlp::SMEM_BYTES, st, mx, mw1, mxl, mw1l, mw2, mw2l, b1, b2, res, ldr, Z, ldz, M, ln_w, ln_b, Z_hi, Z_lo, trace);
    LAUNCH_CHECK(h);
    return IDB_OK;
}
