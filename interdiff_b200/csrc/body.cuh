// Body model state shared by lbs.cu / geometry.cu / correction.cu.
#pragma once
#include "common.cuh"

struct BodyModel {
    int V = 0, J = 0, NB = 0, Fc = 0, Kp = 0;
    float *v_templT = nullptr, *shapedirsT = nullptr, *posedirsT = nullptr, *weightsT = nullptr;
    float *J_templ = nullptr, *J_shape = nullptr;
    int32_t *parents = nullptr, *faces = nullptr, *vf_off = nullptr, *vf_ent = nullptr;
    std::vector<void*> owned;
    // per-call workspace
    int capF = 0;
    float *A = nullptr, *pose_map = nullptr;
};

int idb_body_workspace(idb_handle* h, int F);
