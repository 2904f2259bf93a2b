// Body model state shared by lbs.cu / geometry.cu / correction.cu.
#pragma once
#include "common.cuh"

constexpr int NN_CLUSTERS = 256;

struct BodyModel {
    int V = 0, J = 0, NB = 0, Fc = 0, Kp = 0;
    float *v_templT = nullptr, *shapedirsT = nullptr, *posedirsT = nullptr, *weightsT = nullptr;
    float *J_templ = nullptr, *J_shape = nullptr;
    int32_t *parents = nullptr, *faces = nullptr, *vf_off = nullptr, *vf_ent = nullptr;
    int32_t* depth = nullptr; int max_depth = 0;     // tree level of every joint (root 0)
    // nearest-neighbour acceleration: vertices grouped into NN_CLUSTERS spatially compact clusters (k-means on the
    // template at init); nn_vid = vertex ids sorted by (cluster, id), nn_off = cluster offsets [NN_CLUSTERS + 1]
    uint16_t* nn_vid = nullptr; int32_t* nn_off = nullptr;
    // tensor-core pose blend (lbs.cu): posedirs * 2^8 as the W operand of the split-precision GEMM, fp16 (hi, lo) pairs
    // [Nb][Kld] with row n = v*3 + c (rows >= 3V and columns >= Kp zero), Nb = 3V rounded up to 256, Kld = Kp rounded up to 8
    __half *pd_hi = nullptr, *pd_lo = nullptr; int Nb = 0, Kld = 0;
    // skinning weights in ELL form: sk_n[v] non-zero bones of vertex v (<= SK_MAX), sk_j / sk_w [SK_MAX][V]; sk_dense = some
    // vertex has more than SK_MAX non-zero weights (then the kernel walks all J bones of weightsT instead)
    unsigned char *sk_n = nullptr, *sk_j = nullptr; float* sk_w = nullptr; bool sk_dense = false; int sk_max = 0;
    std::vector<void*> owned;
    // per-call workspace
    int capF = 0;
    float *A = nullptr, *pose_map = nullptr;
    __half *pm_hi = nullptr, *pm_lo = nullptr;   // pose_map as fp16 (hi, lo) pairs [F][Kld]
    float* blend = nullptr;                      // [F][Nb] pose-blend offsets * 2^8
};
constexpr int SK_MAX = 8;

int idb_body_workspace(idb_handle* h, int F);
