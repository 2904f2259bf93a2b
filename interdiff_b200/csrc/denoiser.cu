// Denoiser (MDM decoder path): per-timestep transformer over (batch, frames, channels).
// Replaces reference model/diffusion_smpl.py:226-246, model/diffusion_skeleton.py:218-257,
// model/layers.py:24-26,42-43,258-264, model/sublayers.py:18-35,295-375 and the
// torch.nn.TransformerDecoderLayer instances of layers 0/7.
//
// Token layout in HBM: row m = b*T + t (batch-major; a sample's frames are adjacent rows, which
// is what the 3-tap QaN filter and the per-sample attention want), features contiguous.
// All arithmetic fp32 (see gemm.cu header for why).
#include "common.cuh"

#include <cstdarg>

// ------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------
namespace {

constexpr int D = 256;  // d_model is fixed at 256 for both shipped models (checked at init)

// TimestepEmbedder.forward (model/layers.py:42-43): pe[t] -> Linear -> SiLU -> Linear. grid B, block 256
__global__ void k_temb(const float* __restrict__ pe, const long long* __restrict__ t, const float* __restrict__ w0T,
                       const float* __restrict__ b0, const float* __restrict__ w2T, const float* __restrict__ b2,
                       float* __restrict__ out, int pe_rows) {
    __shared__ float s_in[D], s_h[D];
    const int b = blockIdx.x, n = threadIdx.x;
    long long ti = t[b];
    if (ti < 0) ti = 0;
    if (ti >= pe_rows) ti = pe_rows - 1;
    s_in[n] = pe[(size_t)ti * D + n];
    __syncthreads();
    float a = b0[n];
#pragma unroll 8
    for (int k = 0; k < D; k++) a = fmaf(w0T[k * D + n], s_in[k], a);
    s_h[n] = silu(a);
    __syncthreads();
    a = b2[n];
#pragma unroll 8
    for (int k = 0; k < D; k++) a = fmaf(w2T[k * D + n], s_h[k], a);
    out[b * D + n] = a;
}

// Input embedding (model/diffusion_smpl.py:227-232): h[b,t,:] = W_in . x[b,:,t] + b_in + temb[b] + pe[t].
// x layout (B,1,C,T).  grid B, block 256 (one output feature per thread, all T frames in registers).
template <int TT>
__global__ void k_embed(const float* __restrict__ x, const float* __restrict__ w_inT, const float* __restrict__ b_in,
                        const float* __restrict__ temb, const float* __restrict__ pe, float* __restrict__ h,
                        int C, int T) {
    extern __shared__ float xs[];  // [C][T]
    const int b = blockIdx.x, n = threadIdx.x;
    for (int i = n; i < C * T; i += blockDim.x) xs[i] = x[(size_t)b * C * T + i];
    __syncthreads();
    for (int t0 = 0; t0 < T; t0 += TT) {
        float acc[TT];
#pragma unroll
        for (int i = 0; i < TT; i++) acc[i] = 0.f;
        for (int c = 0; c < C; c++) {
            const float w = w_inT[c * D + n];
            const float* xr = xs + c * T + t0;
#pragma unroll
            for (int i = 0; i < TT; i++)
                if (t0 + i < T) acc[i] = fmaf(w, xr[i], acc[i]);
        }
        const float base = b_in[n] + temb[b * D + n];
#pragma unroll
        for (int i = 0; i < TT; i++)
            if (t0 + i < T) h[((size_t)b * T + t0 + i) * D + n] = (acc[i] + base) + pe[(size_t)(t0 + i) * D + n];
    }
}

// out[m,:] = LayerNorm(a[m,:] (+ r[m,:])) * w + b.  warp per row, D = 256.
__global__ void k_add_ln(const float* __restrict__ a, const float* __restrict__ r, const float* __restrict__ w,
                         const float* __restrict__ bb, float* __restrict__ out, int M) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= M) return;
    const float4* pa = reinterpret_cast<const float4*>(a + (size_t)warp * D);
    float v[8];
    float4 u0 = pa[lane], u1 = pa[lane + 32];
    v[0] = u0.x; v[1] = u0.y; v[2] = u0.z; v[3] = u0.w; v[4] = u1.x; v[5] = u1.y; v[6] = u1.z; v[7] = u1.w;
    if (r) {
        const float4* pr = reinterpret_cast<const float4*>(r + (size_t)warp * D);
        float4 q0 = pr[lane], q1 = pr[lane + 32];
        v[0] += q0.x; v[1] += q0.y; v[2] += q0.z; v[3] += q0.w; v[4] += q1.x; v[5] += q1.y; v[6] += q1.z; v[7] += q1.w;
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) s += v[i];
    const float mean = warp_sum(s) * (1.0f / D);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) { const float d = v[i] - mean; q = fmaf(d, d, q); }
    const float rstd = 1.0f / sqrtf(warp_sum(q) * (1.0f / D) + 1e-5f);
    const float4* pw = reinterpret_cast<const float4*>(w);
    const float4* pb = reinterpret_cast<const float4*>(bb);
    float4 w0 = pw[lane], w1 = pw[lane + 32], b0 = pb[lane], b1 = pb[lane + 32];
    float4 o0, o1;
    o0.x = (v[0] - mean) * rstd * w0.x + b0.x; o0.y = (v[1] - mean) * rstd * w0.y + b0.y;
    o0.z = (v[2] - mean) * rstd * w0.z + b0.z; o0.w = (v[3] - mean) * rstd * w0.w + b0.w;
    o1.x = (v[4] - mean) * rstd * w1.x + b1.x; o1.y = (v[5] - mean) * rstd * w1.y + b1.y;
    o1.z = (v[6] - mean) * rstd * w1.z + b1.z; o1.w = (v[7] - mean) * rstd * w1.w + b1.w;
    float4* po = reinterpret_cast<float4*>(out + (size_t)warp * D);
    po[lane] = o0; po[lane + 32] = o1;
}

// Multi-head attention core softmax(q k^T / sqrt(hd)) v for one (sample, head) per block.
// q rows: (b*qsb + i*qst), k/v rows: (b*ksb + j*kst)  (strides in rows) so the same kernel reads
// batch-major tokens (self-attention) and the seq-first memory K/V (cross-attention).
__global__ void k_mha(const float* __restrict__ q, int ldq, int qsb, int qst,
                      const float* __restrict__ k, const float* __restrict__ v, int ldkv, int ksb, int kst,
                      float* __restrict__ out, int ldo, int Tq, int Tk, int H, float scale) {
    constexpr int HD = 64;
    extern __shared__ float sm[];
    float* sq = sm;                    // [Tq][HD]
    float* sk = sq + Tq * HD;          // [Tk][HD+1]
    float* sv = sk + Tk * (HD + 1);    // [Tk][HD]
    float* sp = sv + Tk * HD;          // [Tq][Tk]
    const int b = blockIdx.x / H, hh = blockIdx.x % H, tid = threadIdx.x, nt = blockDim.x;
    for (int i = tid; i < Tq * HD; i += nt) {
        int r = i / HD, c = i % HD;
        sq[i] = q[(size_t)(b * qsb + r * qst) * ldq + hh * HD + c] * scale;
    }
    for (int i = tid; i < Tk * HD; i += nt) {
        int r = i / HD, c = i % HD;
        size_t row = (size_t)(b * ksb + r * kst) * ldkv;
        sk[r * (HD + 1) + c] = k[row + hh * HD + c];
        sv[i] = v[row + hh * HD + c];
    }
    __syncthreads();
    for (int i = tid; i < Tq * Tk; i += nt) {
        int r = i / Tk, c = i % Tk;
        float a = 0.f;
#pragma unroll 16
        for (int d = 0; d < HD; d++) a = fmaf(sq[r * HD + d], sk[c * (HD + 1) + d], a);
        sp[i] = a;
    }
    __syncthreads();
    const int warp = tid >> 5, lane = tid & 31, nw = nt >> 5;
    for (int r = warp; r < Tq; r += nw) {
        float mx = -INFINITY;
        for (int c = lane; c < Tk; c += 32) mx = fmaxf(mx, sp[r * Tk + c]);
        mx = warp_max(mx);
        float s = 0.f;
        for (int c = lane; c < Tk; c += 32) { float e = expf(sp[r * Tk + c] - mx); sp[r * Tk + c] = e; s += e; }
        s = warp_sum(s);
        const float inv = 1.0f / s;
        for (int c = lane; c < Tk; c += 32) sp[r * Tk + c] *= inv;
    }
    __syncthreads();
    for (int i = tid; i < Tq * HD; i += nt) {
        int r = i / HD, c = i % HD;
        float a = 0.f;
        for (int j = 0; j < Tk; j++) a = fmaf(sp[r * Tk + j], sv[j * HD + c], a);
        out[(size_t)(b * qsb + r * qst) * ldo + hh * HD + c] = a;
    }
}

// QaN block + residual + LayerNorm1 for one sample per block (model/sublayers.py:343-352 + :332):
//   P[t', s*N+n] = h[t'] . Qt[s][n]         (Qt = rotary-folded, 1/16-scaled normalised queries)
//   a[t,n,:] = softmax over valid key slots s in {t-1,t,t+1} of P[t+s-1, s*N+n]
//   y[t] = sum_s (sum_n wk[n] a[t,n,s]) h[t+s-1];   out = LN1(h + y)
__global__ void k_qan_ln(const float* __restrict__ h, const float* __restrict__ qt, const float* __restrict__ wk,
                         const float* __restrict__ lnw, const float* __restrict__ lnb, float* __restrict__ out,
                         int T, int N) {
    constexpr int LD = D + 4;
    extern __shared__ __align__(16) float sm[];
    float* sh = sm;                 // [T][LD]
    float* sqt = sh + T * LD;       // [32][LD]  (3N <= 32 vectors)
    float* sP = sqt + 32 * LD;      // [T][32]
    float* sc = sP + T * 32;        // [T][4]
    const int b = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
    const int NQ = 3 * N;
    for (int i = tid; i < T * (D / 4); i += nt) {
        int r = i / (D / 4), c = i % (D / 4);
        *reinterpret_cast<float4*>(sh + r * LD + c * 4) = reinterpret_cast<const float4*>(h + ((size_t)b * T + r) * D)[c];
    }
    for (int i = tid; i < 32 * (D / 4); i += nt) {
        int r = i / (D / 4), c = i % (D / 4);
        float4 v = (r < NQ) ? reinterpret_cast<const float4*>(qt + (size_t)r * D)[c] : make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(sqt + r * LD + c * 4) = v;
    }
    __syncthreads();
    for (int i = tid; i < T * 32; i += nt) {
        int r = i >> 5, j = i & 31;
        float a = 0.f;
        const float4* ph = reinterpret_cast<const float4*>(sh + r * LD);
        const float4* pq = reinterpret_cast<const float4*>(sqt + j * LD);
#pragma unroll 8
        for (int c = 0; c < D / 4; c++) {
            float4 x = ph[c], y = pq[c];
            a = fmaf(x.x, y.x, a); a = fmaf(x.y, y.y, a); a = fmaf(x.z, y.z, a); a = fmaf(x.w, y.w, a);
        }
        sP[i] = a;
    }
    __syncthreads();
    for (int t = tid; t < T; t += nt) {
        float c0 = 0.f, c1 = 0.f, c2 = 0.f;
        for (int n = 0; n < N; n++) {
            const float l1 = sP[t * 32 + N + n];
            const bool v0 = t > 0, v2 = t < T - 1;
            const float l0 = v0 ? sP[(t - 1) * 32 + n] : -INFINITY;
            const float l2 = v2 ? sP[(t + 1) * 32 + 2 * N + n] : -INFINITY;
            const float mx = fmaxf(l1, fmaxf(l0, l2));
            const float e0 = v0 ? expf(l0 - mx) : 0.f, e1 = expf(l1 - mx), e2 = v2 ? expf(l2 - mx) : 0.f;
            const float inv = 1.0f / (e0 + e1 + e2);
            const float w = wk[n];
            c0 = fmaf(w, e0 * inv, c0); c1 = fmaf(w, e1 * inv, c1); c2 = fmaf(w, e2 * inv, c2);
        }
        sc[t * 4 + 0] = c0; sc[t * 4 + 1] = c1; sc[t * 4 + 2] = c2;
    }
    __syncthreads();
    const int warp = tid >> 5, lane = tid & 31, nw = nt >> 5;
    for (int t = warp; t < T; t += nw) {
        const float c0 = sc[t * 4], c1 = sc[t * 4 + 1], c2 = sc[t * 4 + 2];
        float v[8];
#pragma unroll
        for (int half = 0; half < 2; half++) {
            const int c = (lane + half * 32) * 4;
            float4 x1 = *reinterpret_cast<const float4*>(sh + t * LD + c);
            float4 y = make_float4(c1 * x1.x, c1 * x1.y, c1 * x1.z, c1 * x1.w);
            if (t > 0) {
                float4 x0 = *reinterpret_cast<const float4*>(sh + (t - 1) * LD + c);
                y.x = fmaf(c0, x0.x, y.x); y.y = fmaf(c0, x0.y, y.y); y.z = fmaf(c0, x0.z, y.z); y.w = fmaf(c0, x0.w, y.w);
            }
            if (t < T - 1) {
                float4 x2 = *reinterpret_cast<const float4*>(sh + (t + 1) * LD + c);
                y.x = fmaf(c2, x2.x, y.x); y.y = fmaf(c2, x2.y, y.y); y.z = fmaf(c2, x2.z, y.z); y.w = fmaf(c2, x2.w, y.w);
            }
            v[half * 4 + 0] = x1.x + y.x; v[half * 4 + 1] = x1.y + y.y; v[half * 4 + 2] = x1.z + y.z; v[half * 4 + 3] = x1.w + y.w;
        }
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 8; i++) s += v[i];
        const float mean = warp_sum(s) * (1.0f / D);
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < 8; i++) { const float d = v[i] - mean; q = fmaf(d, d, q); }
        const float rstd = 1.0f / sqrtf(warp_sum(q) * (1.0f / D) + 1e-5f);
#pragma unroll
        for (int half = 0; half < 2; half++) {
            const int c = (lane + half * 32) * 4;
            float4 w4 = *reinterpret_cast<const float4*>(lnw + c), b4 = *reinterpret_cast<const float4*>(lnb + c), o;
            o.x = (v[half * 4 + 0] - mean) * rstd * w4.x + b4.x; o.y = (v[half * 4 + 1] - mean) * rstd * w4.y + b4.y;
            o.z = (v[half * 4 + 2] - mean) * rstd * w4.z + b4.z; o.w = (v[half * 4 + 3] - mean) * rstd * w4.w + b4.w;
            *reinterpret_cast<float4*>(out + ((size_t)b * T + t) * D + c) = o;
        }
    }
}

// Output heads + layout back to (B,1,C,T) + optional inpainting blend
// (model/diffusion_smpl.py:234-237,245; diffusion/gaussian_diffusion.py:307-311).
// variant 0: C = Clin (body | obj heads).  variant 1 (skeleton, model/diffusion_skeleton.py:218-248):
// Clin = c_body + 7; output channels [body | R(quat) p + trans for P points | pose7].
// grid B, block 256; thread n < Clin computes output feature n for all frames.
template <int TT>
__global__ void k_heads(const float* __restrict__ h, const float* __restrict__ w_outT, const float* __restrict__ b_out,
                        const float* __restrict__ zero_pose, const float* __restrict__ gt,
                        const unsigned char* __restrict__ mask, float* __restrict__ out,
                        int T, int Clin, int C, int variant, int c_body, int n_points) {
    extern __shared__ __align__(16) float sm[];
    float* sh = sm;              // [T][D]
    float* so = sh + T * D;      // [Clin][T]  linear outputs
    const int b = blockIdx.x, n = threadIdx.x;
    for (int i = n; i < T * (D / 4); i += blockDim.x)
        reinterpret_cast<float4*>(sh)[i] = reinterpret_cast<const float4*>(h + (size_t)b * T * D)[i];
    __syncthreads();
    if (n < Clin) {
        for (int t0 = 0; t0 < T; t0 += TT) {
            float acc[TT];
#pragma unroll
            for (int i = 0; i < TT; i++) acc[i] = 0.f;
            for (int k = 0; k < D; k++) {
                const float w = w_outT[k * Clin + n];
#pragma unroll
                for (int i = 0; i < TT; i++)
                    if (t0 + i < T) acc[i] = fmaf(w, sh[(t0 + i) * D + k], acc[i]);
            }
#pragma unroll
            for (int i = 0; i < TT; i++)
                if (t0 + i < T) so[n * T + t0 + i] = acc[i] + b_out[n];
        }
    }
    __syncthreads();
    for (int i = n; i < C * T; i += blockDim.x) {
        const int c = i / T, t = i % T;
        float v;
        if (variant == 0 || c < c_body) {
            v = so[c * T + t];
        } else if (c >= c_body + 3 * n_points) {
            v = so[(c - 3 * n_points) * T + t];
        } else {
            // calc_obj_pred: pose = [trans3, quat xyzw]; quaternion_to_matrix on (w,x,y,z) un-normalised
            const int p = (c - c_body) / 3, ax = (c - c_body) % 3;
            const float* ps = so + c_body * T + t;  // pose component j at ps[j*T]
            const float tx = ps[0 * T], ty = ps[1 * T], tz = ps[2 * T];
            const float qi = ps[3 * T], qj = ps[4 * T], qk = ps[5 * T], qr = ps[6 * T];
            const float two_s = 2.0f / (qr * qr + qi * qi + qj * qj + qk * qk);
            float r0, r1, r2, tr;
            if (ax == 0) { r0 = 1 - two_s * (qj * qj + qk * qk); r1 = two_s * (qi * qj - qk * qr); r2 = two_s * (qi * qk + qj * qr); tr = tx; }
            else if (ax == 1) { r0 = two_s * (qi * qj + qk * qr); r1 = 1 - two_s * (qi * qi + qk * qk); r2 = two_s * (qj * qk - qi * qr); tr = ty; }
            else { r0 = two_s * (qi * qk - qj * qr); r1 = two_s * (qj * qk + qi * qr); r2 = 1 - two_s * (qi * qi + qj * qj); tr = tz; }
            const float* zp = zero_pose + ((size_t)b * n_points + p) * 3;
            v = (r0 * zp[0] + r1 * zp[1] + r2 * zp[2]) + tr;
        }
        const size_t o = (size_t)b * C * T + i;
        if (mask && mask[o]) v = gt[o];
        out[o] = v;
    }
}

__global__ void k_fill_t(long long* t_dev, const StepParams* tbl, const int* counter, int B) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B) t_dev[i] = tbl[*counter].t;
}

}  // namespace

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------

int idb_fail(idb_handle* h, int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (h) h->err = buf;
    return code;
}

int idb_dev_alloc(idb_handle* h, float** p, size_t n) {
    CUDA_TRY(h, cudaMalloc((void**)p, (n ? n : 1) * sizeof(float)));
    return IDB_OK;
}
int idb_upload(idb_handle* h, float** p, const float* host, size_t n) {
    int rc = idb_dev_alloc(h, p, n);
    if (rc) return rc;
    CUDA_TRY(h, cudaMemcpy(*p, host, n * sizeof(float), cudaMemcpyDefault));
    return IDB_OK;
}

static void denoiser_free_bound(Denoiser& d) {
    for (float* p : d.bound) cudaFree(p);
    d.bound.clear();
    if (d.t_dev) { cudaFree(d.t_dev); d.t_dev = nullptr; }
    d.B = d.T = d.M = 0;
}

void idb_denoiser_release(idb_handle* h) {
    Denoiser& d = h->den;
    denoiser_free_bound(d);
    for (auto& kv : d.raw) cudaFree(kv.second.p);
    d.raw.clear();
    for (float* p : d.owned) cudaFree(p);
    d.owned.clear();
    d.layers.clear();
    d.committed = false;
}

extern "C" int idb_denoiser_init(idb_handle* h, const idb_denoiser_config* cfg) {
    if (!h || !cfg) return IDB_ERR_ARG;
    if (cfg->d_model != D) return idb_fail(h, IDB_ERR_ARG, "d_model must be 256 (got %d)", cfg->d_model);
    if (cfg->n_heads * 64 != cfg->d_model) return idb_fail(h, IDB_ERR_ARG, "head_dim must be 64");
    if (cfg->n_queries * 3 > 32) return idb_fail(h, IDB_ERR_ARG, "n_queries must be <= 10");
    if (cfg->d_ff % 4 || cfg->n_layers < 1 || cfg->n_layers > 32) return idb_fail(h, IDB_ERR_ARG, "bad d_ff / n_layers");
    if (cfg->variant == 1 && (cfg->c_obj != 3 * cfg->n_points || cfg->c_extra != 7))
        return idb_fail(h, IDB_ERR_ARG, "skeleton variant needs c_obj == 3*n_points and c_extra == 7");
    idb_denoiser_release(h);
    h->den.cfg = *cfg;
    h->den.configured = true;
    return IDB_OK;
}

extern "C" int idb_denoiser_load(idb_handle* h, const char* name, const float* data, const int64_t* shape, int ndim) {
    if (!h || !name || !data) return IDB_ERR_ARG;
    Denoiser& d = h->den;
    if (!d.configured) return idb_fail(h, IDB_ERR_STATE, "idb_denoiser_init first");
    std::string s(name);
    // tensors on the hot path only (the conditioning encoder / PointNet++ are a later row)
    bool want = s.rfind("decoder.layers.", 0) == 0 || s.rfind("bodyEmbedding.", 0) == 0 || s.rfind("objEmbedding.", 0) == 0 ||
                s.rfind("bodyFinalLinear.", 0) == 0 || s.rfind("objFinalLinear.", 0) == 0 ||
                s.rfind("embedTimeStep.time_embed.", 0) == 0 || s == "PositionalEmbedding.pe";
    if (!want || s.find("inv_freq") != std::string::npos) return IDB_OK;
    DevTensor t;
    t.shape.assign(shape, shape + ndim);
    auto it = d.raw.find(s);
    if (it != d.raw.end()) { cudaFree(it->second.p); d.raw.erase(it); }
    int rc = idb_upload(h, &t.p, data, t.numel());
    if (rc) return rc;
    d.raw[s] = t;
    d.committed = false;
    return IDB_OK;
}

namespace {
struct Packer {
    idb_handle* h;
    Denoiser& d;
    int rc = IDB_OK;
    const DevTensor* get(const std::string& n, std::initializer_list<int64_t> shape) {
        auto it = d.raw.find(n);
        if (it == d.raw.end()) { rc = idb_fail(h, IDB_ERR_STATE, "missing weight '%s'", n.c_str()); return nullptr; }
        std::vector<int64_t> want(shape);
        if (it->second.shape != want) {
            rc = idb_fail(h, IDB_ERR_STATE, "weight '%s' has the wrong shape", n.c_str());
            return nullptr;
        }
        return &it->second;
    }
    std::vector<float> host(const DevTensor* t) {
        std::vector<float> v(t->numel());
        cudaMemcpy(v.data(), t->p, v.size() * sizeof(float), cudaMemcpyDeviceToHost);
        return v;
    }
    float* up(const std::vector<float>& v) {
        float* p = nullptr;
        if (idb_upload(h, &p, v.data(), v.size())) { rc = IDB_ERR_CUDA; return nullptr; }
        d.owned.push_back(p);
        return p;
    }
};
}  // namespace

extern "C" int idb_denoiser_commit(idb_handle* h) {
    if (!h) return IDB_ERR_ARG;
    Denoiser& d = h->den;
    if (!d.configured) return idb_fail(h, IDB_ERR_STATE, "idb_denoiser_init first");
    for (float* p : d.owned) cudaFree(p);
    d.owned.clear();
    d.layers.clear();
    const idb_denoiser_config& c = d.cfg;
    const int F = c.d_ff, N = c.n_queries, H = c.n_heads, HD = D / H;
    Packer P{h, d};
#define GET(var, name, ...) const DevTensor* var = P.get(name, {__VA_ARGS__}); if (!var) return P.rc;
    // input embedding: W_in^T [C][D] (zero rows for un-embedded channels), bias = b_body + b_obj
    {
        GET(wb, "bodyEmbedding.weight", D, c.c_body) GET(bb, "bodyEmbedding.bias", D)
        GET(wo, "objEmbedding.weight", D, c.c_obj) GET(bo, "objEmbedding.bias", D)
        const int C = c.c_body + c.c_obj + c.c_extra;
        auto hwb = P.host(wb), hbb = P.host(bb), hwo = P.host(wo), hbo = P.host(bo);
        std::vector<float> wT((size_t)C * D, 0.f), bsum(D);
        for (int n = 0; n < D; n++) {
            for (int k = 0; k < c.c_body; k++) wT[(size_t)k * D + n] = hwb[(size_t)n * c.c_body + k];
            for (int k = 0; k < c.c_obj; k++) wT[(size_t)(c.c_body + k) * D + n] = hwo[(size_t)n * c.c_obj + k];
            bsum[n] = hbb[n] + hbo[n];
        }
        d.w_inT = P.up(wT); d.b_in = P.up(bsum);
    }
    // output heads: [D][Clin] k-major; Clin = c_body + (variant 0 ? c_obj : 7)
    {
        const int c2 = c.variant == 0 ? c.c_obj : 7, Clin = c.c_body + c2;
        GET(wb, "bodyFinalLinear.weight", c.c_body, D) GET(bb, "bodyFinalLinear.bias", c.c_body)
        GET(wo, "objFinalLinear.weight", c2, D) GET(bo, "objFinalLinear.bias", c2)
        auto hwb = P.host(wb), hbb = P.host(bb), hwo = P.host(wo), hbo = P.host(bo);
        std::vector<float> wT((size_t)D * Clin), bcat(Clin);
        for (int n = 0; n < Clin; n++) {
            for (int k = 0; k < D; k++)
                wT[(size_t)k * Clin + n] = n < c.c_body ? hwb[(size_t)n * D + k] : hwo[(size_t)(n - c.c_body) * D + k];
            bcat[n] = n < c.c_body ? hbb[n] : hbo[n - c.c_body];
        }
        d.w_outT = P.up(wT); d.b_out = P.up(bcat);
    }
    // timestep MLP (k-major) + sinusoid table
    {
        GET(w0, "embedTimeStep.time_embed.0.weight", D, D) GET(b0, "embedTimeStep.time_embed.0.bias", D)
        GET(w2, "embedTimeStep.time_embed.2.weight", D, D) GET(b2, "embedTimeStep.time_embed.2.bias", D)
        auto tr = [&](const DevTensor* w) {
            auto hw = P.host(w);
            std::vector<float> t((size_t)D * D);
            for (int n = 0; n < D; n++) for (int k = 0; k < D; k++) t[(size_t)k * D + n] = hw[(size_t)n * D + k];
            return t;
        };
        d.te_w0T = P.up(tr(w0)); d.te_w2T = P.up(tr(w2));
        d.te_b0 = b0->p; d.te_b2 = b2->p;
        auto it = d.raw.find("PositionalEmbedding.pe");
        if (it == d.raw.end()) return idb_fail(h, IDB_ERR_STATE, "missing weight 'PositionalEmbedding.pe'");
        const DevTensor& pe = it->second;
        if (pe.shape.empty() || pe.shape.back() != D) return idb_fail(h, IDB_ERR_STATE, "pe has the wrong shape");
        d.pe = pe.p; d.pe_rows = (int)(pe.numel() / D);
    }
    for (int l = 0; l < c.n_layers; l++) {
        DenoiserLayer L;
        L.qan = (c.qan_mask >> l) & 1;
        const std::string p = "decoder.layers." + std::to_string(l) + ".";
        if (L.qan) {
            GET(q, p + "queries", N, D) GET(wk, p + "wk", N, 1)
            // fold: per-head unit norm (+1e-6), / sqrt(hd) (model/sublayers.py:18-35), * D^-0.5
            // (LocalAttention scale), rotate by o_s = q_pos - k_pos (rotary pairs (i, i+D/2))
            auto hq = P.host(q);
            std::vector<float> qt((size_t)3 * N * D);
            for (int n = 0; n < N; n++) {
                std::vector<double> qn(D);
                for (int hh = 0; hh < H; hh++) {
                    double nrm = 0;
                    for (int e = 0; e < HD; e++) { double v = hq[(size_t)n * D + hh * HD + e]; nrm += v * v; }
                    nrm = std::sqrt(nrm) + 1e-6;
                    for (int e = 0; e < HD; e++)
                        qn[hh * HD + e] = hq[(size_t)n * D + hh * HD + e] / nrm / std::sqrt((double)HD) * std::pow((double)D, -0.5);
                }
                for (int s = 0; s < 3; s++) {
                    const double o = c.rotary_offsets[s];
                    for (int i = 0; i < D / 2; i++) {
                        const double f = o * (double)(float)(1.0 / std::pow(10000.0, (double)(2 * i) / D));
                        const double cs = std::cos(f), sn = std::sin(f);
                        qt[((size_t)s * N + n) * D + i] = (float)(qn[i] * cs - qn[i + D / 2] * sn);
                        qt[((size_t)s * N + n) * D + i + D / 2] = (float)(qn[i + D / 2] * cs + qn[i] * sn);
                    }
                }
            }
            L.qt = P.up(qt); L.wk = wk->p;
        } else {
            GET(w, p + "self_attn.in_proj_weight", 3 * D, D) GET(b, p + "self_attn.in_proj_bias", 3 * D)
            GET(wo, p + "self_attn.out_proj.weight", D, D) GET(bo, p + "self_attn.out_proj.bias", D)
            L.w_qkv = w->p; L.b_qkv = b->p; L.w_o = wo->p; L.b_o = bo->p;
        }
        {
            GET(w, p + "multihead_attn.in_proj_weight", 3 * D, D) GET(b, p + "multihead_attn.in_proj_bias", 3 * D)
            GET(wo, p + "multihead_attn.out_proj.weight", D, D) GET(bo, p + "multihead_attn.out_proj.bias", D)
            L.w_qc = w->p; L.b_qc = b->p; L.w_kvc = w->p + (size_t)D * D; L.b_kvc = b->p + D; L.w_oc = wo->p; L.b_oc = bo->p;
            GET(w1, p + "linear1.weight", F, D) GET(b1, p + "linear1.bias", F)
            GET(w2, p + "linear2.weight", D, F) GET(b2, p + "linear2.bias", D)
            L.w1 = w1->p; L.b1 = b1->p; L.w2 = w2->p; L.b2 = b2->p;
            GET(n1w, p + "norm1.weight", D) GET(n1b, p + "norm1.bias", D) GET(n2w, p + "norm2.weight", D)
            GET(n2b, p + "norm2.bias", D) GET(n3w, p + "norm3.weight", D) GET(n3b, p + "norm3.bias", D)
            L.ln1w = n1w->p; L.ln1b = n1b->p; L.ln2w = n2w->p; L.ln2b = n2b->p; L.ln3w = n3w->p; L.ln3b = n3b->p;
        }
        d.layers.push_back(L);
    }
#undef GET
    if (P.rc) return P.rc;
    d.committed = true;
    denoiser_free_bound(d);
    return IDB_OK;
}

extern "C" int idb_denoiser_bind(idb_handle* h, int B, int T, int Tm, const float* cond, const float* zero_pose_obj, void* stream) {
    if (!h || !cond || B <= 0 || T <= 0 || Tm <= 0) return IDB_ERR_ARG;
    Denoiser& d = h->den;
    if (!d.committed) return idb_fail(h, IDB_ERR_STATE, "idb_denoiser_commit first");
    if (T > d.pe_rows) return idb_fail(h, IDB_ERR_ARG, "T exceeds the positional table");
    if (d.cfg.variant == 1 && !zero_pose_obj) return idb_fail(h, IDB_ERR_ARG, "skeleton variant needs zero_pose_obj");
    cudaStream_t st = (cudaStream_t)stream;
    const int M = B * T, F = d.cfg.d_ff;
    const int npts = d.cfg.n_points > 0 ? d.cfg.n_points : 1;
    if (B != d.B || T != d.T || Tm != d.Tm) {
        denoiser_free_bound(d);
        auto A = [&](float** p, size_t n) { int rc = idb_dev_alloc(h, p, n); if (!rc) d.bound.push_back(*p); return rc; };
        int rc = 0;
        rc |= A(&d.cond, (size_t)Tm * B * D); rc |= A(&d.temb, (size_t)B * D); rc |= A(&d.h, (size_t)M * D);
        rc |= A(&d.h2, (size_t)M * D); rc |= A(&d.qkv, (size_t)M * 3 * D); rc |= A(&d.att, (size_t)M * D);
        rc |= A(&d.ff, (size_t)M * F); rc |= A(&d.qc, (size_t)M * D);
        rc |= A(&d.zero_pose, (size_t)B * npts * 3);
        for (auto& L : d.layers) rc |= A(&L.kv_mem, (size_t)Tm * B * 2 * D);
        if (rc) return rc;
        CUDA_TRY(h, cudaMalloc((void**)&d.t_dev, sizeof(long long) * B));
        d.B = B; d.T = T; d.M = M; d.Tm = Tm;
    }
    CUDA_TRY(h, cudaMemcpyAsync(d.cond, cond, sizeof(float) * (size_t)Tm * B * D, cudaMemcpyDefault, st));
    if (zero_pose_obj && d.cfg.n_points > 0)
        CUDA_TRY(h, cudaMemcpyAsync(d.zero_pose, zero_pose_obj, sizeof(float) * (size_t)B * d.cfg.n_points * 3, cudaMemcpyDefault, st));
    // step-invariant cross-attention K|V of the memory (rows j*B + b, seq-first like the reference)
    for (auto& L : d.layers) {
        int rc = idb_gemm(h, d.cond, D, L.w_kvc, D, L.b_kvc, nullptr, 0, L.kv_mem, 2 * D, Tm * B, 2 * D, D, EPI_BIAS, st);
        if (rc) return rc;
    }
    return IDB_OK;
}

// Decoder body on the bound workspaces: d.h holds the embedded tokens on entry and the decoder
// output on exit.
static int denoiser_layers(idb_handle* h, cudaStream_t st) {
    Denoiser& d = h->den;
    const int B = d.B, T = d.T, M = d.M, Tm = d.Tm, F = d.cfg.d_ff, H = d.cfg.n_heads, N = d.cfg.n_queries;
    const float scale = 1.0f / sqrtf(64.0f);
    const int ln_blocks = (M * 32 + 255) / 256;
    const size_t smem_self = sizeof(float) * ((size_t)T * 64 + (size_t)T * 65 + (size_t)T * 64 + (size_t)T * T);
    const size_t smem_cross = sizeof(float) * ((size_t)T * 64 + (size_t)Tm * 65 + (size_t)Tm * 64 + (size_t)T * Tm);
    const size_t smem_qan = sizeof(float) * ((size_t)T * (D + 4) + 32 * (D + 4) + (size_t)T * 32 + (size_t)T * 4);
    float* x = d.h;     // residual stream
    float* y = d.h2;    // ping-pong
    int rc;
    for (auto& L : d.layers) {
        if (L.qan) {
            k_qan_ln<<<B, 256, smem_qan, st>>>(x, L.qt, L.wk, L.ln1w, L.ln1b, y, T, N);
            LAUNCH_CHECK(h);
        } else {
            if ((rc = idb_gemm(h, x, D, L.w_qkv, D, L.b_qkv, nullptr, 0, d.qkv, 3 * D, M, 3 * D, D, EPI_BIAS, st))) return rc;
            k_mha<<<B * H, 128, smem_self, st>>>(d.qkv, 3 * D, T, 1, d.qkv + D, d.qkv + 2 * D, 3 * D, T, 1, d.att, D, T, T, H, scale);
            LAUNCH_CHECK(h);
            if ((rc = idb_gemm(h, d.att, D, L.w_o, D, L.b_o, x, D, d.qc, D, M, D, D, EPI_BIAS | EPI_RES, st))) return rc;
            k_add_ln<<<ln_blocks, 256, 0, st>>>(d.qc, nullptr, L.ln1w, L.ln1b, y, M);
            LAUNCH_CHECK(h);
        }
        // cross attention on y -> x
        if ((rc = idb_gemm(h, y, D, L.w_qc, D, L.b_qc, nullptr, 0, d.qc, D, M, D, D, EPI_BIAS, st))) return rc;
        k_mha<<<B * H, 128, smem_cross, st>>>(d.qc, D, T, 1, L.kv_mem, L.kv_mem + D, 2 * D, 1, B, d.att, D, T, Tm, H, scale);
        LAUNCH_CHECK(h);
        if ((rc = idb_gemm(h, d.att, D, L.w_oc, D, L.b_oc, y, D, d.qc, D, M, D, D, EPI_BIAS | EPI_RES, st))) return rc;
        k_add_ln<<<ln_blocks, 256, 0, st>>>(d.qc, nullptr, L.ln2w, L.ln2b, x, M);
        LAUNCH_CHECK(h);
        // feed forward on x -> y -> x
        if ((rc = idb_gemm(h, x, D, L.w1, D, L.b1, nullptr, 0, d.ff, F, M, F, D, EPI_BIAS | EPI_GELU, st))) return rc;
        if ((rc = idb_gemm(h, d.ff, F, L.w2, F, L.b2, x, D, d.qc, D, M, D, F, EPI_BIAS | EPI_RES, st))) return rc;
        // QaN layers return tgt + (x - tgt) (model/sublayers.py:338-339); that differs from x by
        // <= 1 ulp of max(|x|,|tgt|) and is not reproduced (DESIGN.md "Deviations").
        k_add_ln<<<ln_blocks, 256, 0, st>>>(d.qc, nullptr, L.ln3w, L.ln3b, x, M);
        LAUNCH_CHECK(h);
    }
    return IDB_OK;
}

// x (B,1,C,T) -> x0 prediction (B,1,C,T); optional inpainting blend.  t_dev must hold the timesteps.
int idb_denoiser_run(idb_handle* h, const float* x, const long long* t_dev, const float* gt, const unsigned char* mask,
                     float* out, cudaStream_t st) {
    Denoiser& d = h->den;
    if (!d.B) return idb_fail(h, IDB_ERR_STATE, "idb_denoiser_bind first");
    const idb_denoiser_config& c = d.cfg;
    const int B = d.B, T = d.T, C = c.c_body + c.c_obj + c.c_extra;
    const int Clin = c.c_body + (c.variant == 0 ? c.c_obj : 7);
    k_temb<<<B, D, 0, st>>>(d.pe, t_dev, d.te_w0T, d.te_b0, d.te_w2T, d.te_b2, d.temb, d.pe_rows);
    LAUNCH_CHECK(h);
    const size_t smem_e = sizeof(float) * (size_t)C * T;
    k_embed<8><<<B, D, smem_e, st>>>(x, d.w_inT, d.b_in, d.temb, d.pe, d.h, C, T);
    LAUNCH_CHECK(h);
    int rc = denoiser_layers(h, st);
    if (rc) return rc;
    const size_t smem_h = sizeof(float) * ((size_t)T * D + (size_t)Clin * T);
    k_heads<8><<<B, 256, smem_h, st>>>(d.h, d.w_outT, d.b_out, d.zero_pose, gt, mask, out, T, Clin, C, c.variant, c.c_body, c.n_points);
    LAUNCH_CHECK(h);
    return IDB_OK;
}

int idb_denoiser_prepare_kernels(idb_handle* h) {
    // opt in to > 48 KB dynamic shared memory once (T <= 64 supported)
    CUDA_TRY(h, cudaFuncSetAttribute(k_qan_ln, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    CUDA_TRY(h, cudaFuncSetAttribute(k_mha, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    CUDA_TRY(h, cudaFuncSetAttribute(k_heads<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    CUDA_TRY(h, cudaFuncSetAttribute(k_embed<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    return IDB_OK;
}

int idb_denoiser_fill_t(idb_handle* h, cudaStream_t st) {
    Denoiser& d = h->den;
    k_fill_t<<<(d.B + 127) / 128, 128, 0, st>>>(d.t_dev, h->diff.tbl, h->diff.counter, d.B);
    LAUNCH_CHECK(h);
    return IDB_OK;
}

extern "C" int idb_denoiser_forward(idb_handle* h, const float* x, const int64_t* timesteps, float* out, void* stream) {
    if (!h || !x || !timesteps || !out) return IDB_ERR_ARG;
    return idb_denoiser_run(h, x, (const long long*)timesteps, nullptr, nullptr, out, (cudaStream_t)stream);
}
