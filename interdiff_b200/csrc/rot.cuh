// pytorch3d.transforms 0.7.2 rotation conversions as device functions (reference call sites:
// eval_smpl_short.py:90-91,157-162).  Quaternions are real-part-first.
#pragma once
#include <cuda_runtime.h>

// pytorch3d 0.7.2 chain for one 6-vector; shared with correction.cu
__device__ __forceinline__ void idb_rot6d_to_matrix(const float* d6, float* R) {
    const float a1x = d6[0], a1y = d6[1], a1z = d6[2], a2x = d6[3], a2y = d6[4], a2z = d6[5];
    const float n1 = fmaxf(sqrtf(a1x * a1x + a1y * a1y + a1z * a1z), 1e-12f);
    const float b1x = a1x / n1, b1y = a1y / n1, b1z = a1z / n1;
    const float dot = b1x * a2x + b1y * a2y + b1z * a2z;
    float b2x = a2x - dot * b1x, b2y = a2y - dot * b1y, b2z = a2z - dot * b1z;
    const float n2 = fmaxf(sqrtf(b2x * b2x + b2y * b2y + b2z * b2z), 1e-12f);
    b2x /= n2; b2y /= n2; b2z /= n2;
    R[0] = b1x; R[1] = b1y; R[2] = b1z;
    R[3] = b2x; R[4] = b2y; R[5] = b2z;
    R[6] = b1y * b2z - b1z * b2y; R[7] = b1z * b2x - b1x * b2z; R[8] = b1x * b2y - b1y * b2x;
}

__device__ __forceinline__ void idb_matrix_to_axis_angle(const float* m, float* aa) {
    const float m00 = m[0], m01 = m[1], m02 = m[2], m10 = m[3], m11 = m[4], m12 = m[5], m20 = m[6], m21 = m[7], m22 = m[8];
    float qa[4] = {1.0f + m00 + m11 + m22, 1.0f + m00 - m11 - m22, 1.0f - m00 + m11 - m22, 1.0f - m00 - m11 + m22};
    int best = 0;
    for (int i = 0; i < 4; i++) qa[i] = qa[i] > 0.f ? sqrtf(qa[i]) : 0.f;   // _sqrt_positive_part
    for (int i = 1; i < 4; i++) if (qa[i] > qa[best]) best = i;             // argmax, first maximum
    float c[4];
    if (best == 0) { c[0] = qa[0] * qa[0]; c[1] = m21 - m12; c[2] = m02 - m20; c[3] = m10 - m01; }
    else if (best == 1) { c[0] = m21 - m12; c[1] = qa[1] * qa[1]; c[2] = m10 + m01; c[3] = m02 + m20; }
    else if (best == 2) { c[0] = m02 - m20; c[1] = m10 + m01; c[2] = qa[2] * qa[2]; c[3] = m12 + m21; }
    else { c[0] = m10 - m01; c[1] = m20 + m02; c[2] = m21 + m12; c[3] = qa[3] * qa[3]; }
    const float den = 2.0f * fmaxf(qa[best], 0.1f);
    const float qw = c[0] / den, qx = c[1] / den, qy = c[2] / den, qz = c[3] / den;
    // quaternion_to_axis_angle
    const float nrm = sqrtf(qx * qx + qy * qy + qz * qz);
    const float half = atan2f(nrm, qw);
    const float angle = 2.0f * half;
    const float s = fabsf(angle) < 1e-6f ? 0.5f - (angle * angle) / 48.0f : sinf(half) / angle;
    aa[0] = qx / s; aa[1] = qy / s; aa[2] = qz / s;
}


// axis_angle_to_matrix = quaternion_to_matrix(axis_angle_to_quaternion(aa)) (pytorch3d 0.7.2; small-angle series below 1e-6)
__device__ __forceinline__ void idb_axis_angle_to_matrix(const float* aa, float* R) {
    const float x = aa[0], y = aa[1], z = aa[2];
    const float angle = sqrtf(x * x + y * y + z * z);
    const float half = 0.5f * angle;
    const float s = fabsf(angle) < 1e-6f ? 0.5f - (angle * angle) / 48.0f : sinf(half) / angle;
    const float r = cosf(half), i = x * s, j = y * s, k = z * s;
    const float two_s = 2.0f / (r * r + i * i + j * j + k * k);
    R[0] = 1 - two_s * (j * j + k * k); R[1] = two_s * (i * j - k * r);     R[2] = two_s * (i * k + j * r);
    R[3] = two_s * (i * j + k * r);     R[4] = 1 - two_s * (i * i + k * k); R[5] = two_s * (j * k - i * r);
    R[6] = two_s * (i * k - j * r);     R[7] = two_s * (j * k + i * r);     R[8] = 1 - two_s * (i * i + j * j);
}

// quaternion_to_matrix (pytorch3d 0.7.2), real part first, NOT normalised (the 2 / |q|^2 factor does it)
__device__ __forceinline__ void idb_quaternion_to_matrix(const float* q, float* R) {
    const float r = q[0], i = q[1], j = q[2], k = q[3];
    const float two_s = 2.0f / (r * r + i * i + j * j + k * k);
    R[0] = 1 - two_s * (j * j + k * k); R[1] = two_s * (i * j - k * r);     R[2] = two_s * (i * k + j * r);
    R[3] = two_s * (i * j + k * r);     R[4] = 1 - two_s * (i * i + k * k); R[5] = two_s * (j * k - i * r);
    R[6] = two_s * (i * k - j * r);     R[7] = two_s * (j * k + i * r);     R[8] = 1 - two_s * (i * i + j * j);
}
// matrix_to_quaternion (pytorch3d 0.7.2): 4 candidates, argmax of q_abs (first maximum), no sign standardisation; q = (w,x,y,z)
__device__ __forceinline__ void idb_matrix_to_quaternion(const float* m, float* q) {
    const float m00 = m[0], m01 = m[1], m02 = m[2], m10 = m[3], m11 = m[4], m12 = m[5], m20 = m[6], m21 = m[7], m22 = m[8];
    float qa[4] = {1.0f + m00 + m11 + m22, 1.0f + m00 - m11 - m22, 1.0f - m00 + m11 - m22, 1.0f - m00 - m11 + m22};
    int best = 0;
    for (int i = 0; i < 4; i++) qa[i] = qa[i] > 0.f ? sqrtf(qa[i]) : 0.f;
    for (int i = 1; i < 4; i++) if (qa[i] > qa[best]) best = i;
    float c[4];
    if (best == 0) { c[0] = qa[0] * qa[0]; c[1] = m21 - m12; c[2] = m02 - m20; c[3] = m10 - m01; }
    else if (best == 1) { c[0] = m21 - m12; c[1] = qa[1] * qa[1]; c[2] = m10 + m01; c[3] = m02 + m20; }
    else if (best == 2) { c[0] = m02 - m20; c[1] = m10 + m01; c[2] = qa[2] * qa[2]; c[3] = m12 + m21; }
    else { c[0] = m10 - m01; c[1] = m20 + m02; c[2] = m21 + m12; c[3] = qa[3] * qa[3]; }
    const float den = 2.0f * fmaxf(qa[best], 0.1f);
    q[0] = c[0] / den; q[1] = c[1] / den; q[2] = c[2] / den; q[3] = c[3] / den;
}
