// Body model state shared by lbs.cu / geometry.cu / correction.cu.
#pragma once
#include "common.cuh"

constexpr int NN_CLUSTERS = 256;

struct BodyModel {
    int V = 0, J = 0, NB = 0, Fc = 0, Kp = 0;
    float *v_templT = nullptr, *shapedirsT = nullptr, *posedirsT = nullptr, *weightsT = nullptr;
    float *J_templ = nullptr, *J_shape = nullptr;
    int32_t *parents = nullptr, *faces = nullptr, *vf_off = nullptr, *vf_ent = nullptr;
    // nearest-neighbour acceleration: vertices grouped into NN_CLUSTERS spatially compact clusters (k-means on the
    // template at init); nn_vid = vertex ids sorted by (cluster, id), nn_off = cluster offsets [NN_CLUSTERS + 1]
    uint16_t* nn_vid = nullptr; int32_t* nn_off = nullptr;
    std::vector<void*> owned;
    // per-call workspace
    int capF = 0;
    float *A = nullptr, *pose_map = nullptr;
};

int idb_body_workspace(idb_handle* h, int F);
