// Geometry helpers around the loop (fp32).
//   vertex normals      : reference data/tools.py:4-39
//   signed nearest point: reference tools.py:11-76 (+ the chamfer_distance CUDA ext it calls,
//                         tools.py:45-47: first-minimum squared-L2 argmin)
//   rot6d -> axis-angle : pytorch3d.transforms 0.7.2 rotation_6d_to_matrix / matrix_to_quaternion
//                         / quaternion_to_axis_angle (reference eval_smpl_short.py:90-91,157-162)
#include "common.cuh"
#include "body.cuh"
#include "rot.cuh"

namespace {

// One thread per (frame, vertex); gathers the incident faces in the reference's accumulation
// order (no atomics => deterministic).
__global__ void k_vertex_normals(const float* __restrict__ verts, const int32_t* __restrict__ faces,
                                 const int32_t* __restrict__ vf_off, const int32_t* __restrict__ vf_ent,
                                 float* __restrict__ normals, int V) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    const int f = blockIdx.y;
    if (v >= V) return;
    const float* vb = verts + (size_t)f * V * 3;
    float nx = 0.f, ny = 0.f, nz = 0.f;
    for (int e = vf_off[v]; e < vf_off[v + 1]; e++) {
        const int ent = vf_ent[e], fi = ent >> 2, corner = ent & 3;
        const int i0 = faces[fi * 3], i1 = faces[fi * 3 + 1], i2 = faces[fi * 3 + 2];
        // corner 1: (v2-v1)x(v0-v1); corner 2: (v0-v2)x(v1-v2); corner 0: (v1-v0)x(v2-v0)
        const int ic = corner == 0 ? i0 : (corner == 1 ? i1 : i2);
        const int ia = corner == 0 ? i1 : (corner == 1 ? i2 : i0);
        const int ib = corner == 0 ? i2 : (corner == 1 ? i0 : i1);
        const float cx = vb[ic * 3], cy = vb[ic * 3 + 1], cz = vb[ic * 3 + 2];
        const float ax = vb[ia * 3] - cx, ay = vb[ia * 3 + 1] - cy, az = vb[ia * 3 + 2] - cz;
        const float bx = vb[ib * 3] - cx, by = vb[ib * 3 + 1] - cy, bz = vb[ib * 3 + 2] - cz;
        nx += ay * bz - az * by;
        ny += az * bx - ax * bz;
        nz += ax * by - ay * bx;
    }
    const float len = fmaxf(sqrtf(nx * nx + ny * ny + nz * nz), 1e-6f);  // F.normalize(eps=1e-6)
    float* o = normals + ((size_t)f * V + v) * 3;
    o[0] = nx / len; o[1] = ny / len; o[2] = nz / len;
}

// Brute-force nearest target per query, one frame per blockIdx.y, 256 queries per block, targets
// staged through shared memory in tiles.  d = (dx*dx + dy*dy) + dz*dz without FMA contraction and
// strict '<' so the FIRST minimum wins (same as the restated argmin in oracle/restate.py).
constexpr int NN_TILE = 1024;
__global__ void __launch_bounds__(256)
k_signed_nn(const float* __restrict__ query, const float* __restrict__ target, const float* __restrict__ tnormals,
            float* __restrict__ sdist, int32_t* __restrict__ idx_out, float* __restrict__ vec_out, int Pq, int Pt) {
    __shared__ float4 st[NN_TILE];
    const int f = blockIdx.y, q = blockIdx.x * 256 + threadIdx.x;
    const float* qb = query + (size_t)f * Pq * 3;
    const float* tb = target + (size_t)f * Pt * 3;
    float qx = 0.f, qy = 0.f, qz = 0.f;
    if (q < Pq) { qx = qb[q * 3]; qy = qb[q * 3 + 1]; qz = qb[q * 3 + 2]; }
    float best = INFINITY;
    int bi = 0;
    for (int t0 = 0; t0 < Pt; t0 += NN_TILE) {
        const int nt = min(NN_TILE, Pt - t0);
        __syncthreads();
        for (int i = threadIdx.x; i < nt; i += 256)
            st[i] = make_float4(tb[(size_t)(t0 + i) * 3], tb[(size_t)(t0 + i) * 3 + 1], tb[(size_t)(t0 + i) * 3 + 2], 0.f);
        __syncthreads();
#pragma unroll 4
        for (int i = 0; i < nt; i++) {
            const float4 t = st[i];
            const float dx = qx - t.x, dy = qy - t.y, dz = qz - t.z;
            const float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
            if (d < best) { best = d; bi = t0 + i; }
        }
    }
    if (q >= Pq) return;
    const float vx = qx - tb[(size_t)bi * 3], vy = qy - tb[(size_t)bi * 3 + 1], vz = qz - tb[(size_t)bi * 3 + 2];
    float d = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(vx, vx), __fmul_rn(vy, vy)), __fmul_rn(vz, vz)));
    if (tnormals) {
        const float* n = tnormals + ((size_t)f * Pt + bi) * 3;
        const float dot = __fadd_rn(__fadd_rn(__fmul_rn(n[0], vx), __fmul_rn(n[1], vy)), __fmul_rn(n[2], vz));
        d *= (dot > 0.f) ? 1.0f : ((dot < 0.f) ? -1.0f : 0.0f);
    }
    const size_t o = (size_t)f * Pq + q;
    if (sdist) sdist[o] = d;
    if (idx_out) idx_out[o] = bi;
    if (vec_out) { vec_out[o * 3] = vx; vec_out[o * 3 + 1] = vy; vec_out[o * 3 + 2] = vz; }
}

}  // namespace

namespace {
__global__ void k_rot6d_to_aa(const float* __restrict__ d6, float* __restrict__ aa, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float in[6], R[9], o[3];
    for (int k = 0; k < 6; k++) in[k] = d6[(size_t)i * 6 + k];
    idb_rot6d_to_matrix(in, R);
    idb_matrix_to_axis_angle(R, o);
    aa[(size_t)i * 3] = o[0]; aa[(size_t)i * 3 + 1] = o[1]; aa[(size_t)i * 3 + 2] = o[2];
}
}  // namespace

extern "C" int idb_vertex_normals(idb_handle* h, int F, const float* verts, float* normals, void* stream) {
    if (!h || !verts || !normals || F <= 0) return IDB_ERR_ARG;
    if (!h->body || !h->body->faces) return idb_fail(h, IDB_ERR_STATE, "idb_body_init with faces first");
    BodyModel& m = *h->body;
    dim3 grid((m.V + 127) / 128, F);
    k_vertex_normals<<<grid, 128, 0, (cudaStream_t)stream>>>(verts, m.faces, m.vf_off, m.vf_ent, normals, m.V);
    LAUNCH_CHECK(h);
    return IDB_OK;
}

extern "C" int idb_signed_nn(idb_handle* h, int F, int Pq, int Pt, const float* query, const float* target,
                             const float* target_normals, float* signed_dist, int32_t* idx, float* vec, void* stream) {
    if (!h || !query || !target || F <= 0 || Pq <= 0 || Pt <= 0) return IDB_ERR_ARG;
    dim3 grid((Pq + 255) / 256, F);
    k_signed_nn<<<grid, 256, 0, (cudaStream_t)stream>>>(query, target, target_normals, signed_dist, idx, vec, Pq, Pt);
    LAUNCH_CHECK(h);
    return IDB_OK;
}

extern "C" int idb_rot6d_to_axis_angle(idb_handle* h, int n, const float* rot6d, float* aa, void* stream) {
    if (!h || !rot6d || !aa || n <= 0) return IDB_ERR_ARG;
    k_rot6d_to_aa<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(rot6d, aa, n);
    LAUNCH_CHECK(h);
    return IDB_OK;
}
