// Rotation algebra used on the hot path, as device functions templated on float/double.
// Semantics follow the conventions the reference relies on:
//   * pytorch3d rotation conversions -- reference call sites lib/model/aggregation.py:50-56,224,232,257,265,
//     612,615; lib/model/VPHO.py:316,323; lib/model/head_object.py:57
//   * manopth batch_rodrigues / quat2mat -- behind lib/model/head_mano.py:84
//   * average_quaternion -- lib/utils/transform_fn.py:101-125 (top eigenvector of the weighted outer-product sum)
#pragma once
#include "vpho_common.cuh"

namespace vpho {

__device__ __forceinline__ float v_sqrt(float x) { return sqrtf(x); }
__device__ __forceinline__ double v_sqrt(double x) { return sqrt(x); }
__device__ __forceinline__ float v_sin(float x) { return sinf(x); }
__device__ __forceinline__ double v_sin(double x) { return sin(x); }
__device__ __forceinline__ float v_cos(float x) { return cosf(x); }
__device__ __forceinline__ double v_cos(double x) { return cos(x); }
__device__ __forceinline__ float v_atan2(float y, float x) { return atan2f(y, x); }
__device__ __forceinline__ double v_atan2(double y, double x) { return atan2(y, x); }
__device__ __forceinline__ float v_abs(float x) { return fabsf(x); }
__device__ __forceinline__ double v_abs(double x) { return fabs(x); }
__device__ __forceinline__ float v_max(float a, float b) { return fmaxf(a, b); }
__device__ __forceinline__ double v_max(double a, double b) { return fmax(a, b); }

template <typename T> struct Pi;
template <> struct Pi<float> { __device__ static __forceinline__ float v() { return 3.14159265358979323846f; } };
template <> struct Pi<double> { __device__ static __forceinline__ double v() { return 3.14159265358979323846; } };

// torch.sinc(x) = sin(pi x)/(pi x), 1 at 0   (ATen UnaryOps sinc)
template <typename T>
__device__ __forceinline__ T torch_sinc(T x) {
  if (x == T(0)) return T(1);
  T p = Pi<T>::v() * x;
  return v_sin(p) / p;
}

// pytorch3d.rotation_6d_to_matrix: rows (b1,b2,b3); F.normalize uses x / max(||x||, 1e-12)
template <typename T>
__device__ __forceinline__ void rot6d_to_matrix(const T* d, T* R) {
  T n1 = v_sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
  n1 = v_max(n1, T(1e-12));
  T b1x = d[0] / n1, b1y = d[1] / n1, b1z = d[2] / n1;
  T dot = b1x * d[3] + b1y * d[4] + b1z * d[5];
  T cx = d[3] - dot * b1x, cy = d[4] - dot * b1y, cz = d[5] - dot * b1z;
  T n2 = v_sqrt(cx * cx + cy * cy + cz * cz);
  n2 = v_max(n2, T(1e-12));
  T b2x = cx / n2, b2y = cy / n2, b2z = cz / n2;
  R[0] = b1x; R[1] = b1y; R[2] = b1z;
  R[3] = b2x; R[4] = b2y; R[5] = b2z;
  R[6] = b1y * b2z - b1z * b2y;
  R[7] = b1z * b2x - b1x * b2z;
  R[8] = b1x * b2y - b1y * b2x;
}

// pytorch3d.matrix_to_quaternion (0.7.x: best-conditioned candidate, standardised to w >= 0)
template <typename T>
__device__ __forceinline__ void matrix_to_quaternion(const T* m, T* q) {
  T m00 = m[0], m01 = m[1], m02 = m[2], m10 = m[3], m11 = m[4], m12 = m[5], m20 = m[6], m21 = m[7], m22 = m[8];
  T a[4] = {T(1) + m00 + m11 + m22, T(1) + m00 - m11 - m22, T(1) - m00 + m11 - m22, T(1) - m00 - m11 + m22};
  T qa[4];
  int best = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    qa[i] = a[i] > T(0) ? v_sqrt(a[i]) : T(0);
    if (qa[i] > qa[best]) best = i;   // torch.argmax returns the first maximal index
  }
  T c[4];
  if (best == 0) { c[0] = qa[0] * qa[0]; c[1] = m21 - m12; c[2] = m02 - m20; c[3] = m10 - m01; }
  else if (best == 1) { c[0] = m21 - m12; c[1] = qa[1] * qa[1]; c[2] = m10 + m01; c[3] = m02 + m20; }
  else if (best == 2) { c[0] = m02 - m20; c[1] = m10 + m01; c[2] = qa[2] * qa[2]; c[3] = m12 + m21; }
  else { c[0] = m10 - m01; c[1] = m20 + m02; c[2] = m21 + m12; c[3] = qa[3] * qa[3]; }
  T den = T(2) * v_max(qa[best], T(0.1));
  T s = (c[0] / den) < T(0) ? T(-1) : T(1);
#pragma unroll
  for (int i = 0; i < 4; ++i) q[i] = s * (c[i] / den);
}

// pytorch3d.quaternion_to_matrix
template <typename T>
__device__ __forceinline__ void quaternion_to_matrix(const T* q, T* R) {
  T r = q[0], i = q[1], j = q[2], k = q[3];
  T two_s = T(2) / (r * r + i * i + j * j + k * k);
  R[0] = T(1) - two_s * (j * j + k * k); R[1] = two_s * (i * j - k * r); R[2] = two_s * (i * k + j * r);
  R[3] = two_s * (i * j + k * r); R[4] = T(1) - two_s * (i * i + k * k); R[5] = two_s * (j * k - i * r);
  R[6] = two_s * (i * k - j * r); R[7] = two_s * (j * k + i * r); R[8] = T(1) - two_s * (i * i + j * j);
}

// pytorch3d.axis_angle_to_quaternion (sinc form)
template <typename T>
__device__ __forceinline__ void axis_angle_to_quaternion(const T* a, T* q) {
  T ang = v_sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
  T s = T(0.5) * torch_sinc(ang * T(0.5) / Pi<T>::v());
  q[0] = v_cos(ang * T(0.5));
  q[1] = a[0] * s; q[2] = a[1] * s; q[3] = a[2] * s;
}

// pytorch3d.quaternion_to_axis_angle (sinc form)
template <typename T>
__device__ __forceinline__ void quaternion_to_axis_angle(const T* q, T* a) {
  T n = v_sqrt(q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  T half = v_atan2(n, q[0]);
  T s = T(0.5) * torch_sinc(half / Pi<T>::v());
  a[0] = q[1] / s; a[1] = q[2] / s; a[2] = q[3] / s;
}

template <typename T>
__device__ __forceinline__ void matrix_to_axis_angle(const T* m, T* a) {
  T q[4];
  matrix_to_quaternion(m, q);
  quaternion_to_axis_angle(q, a);
}

// manopth batch_rodrigues: theta = ||a + 1e-8||, quat = (cos(theta/2), sin(theta/2) a/theta), normalised, quat2mat
__device__ __forceinline__ void manopth_rodrigues(const float* a, float* R) {
  float ax = a[0] + 1e-8f, ay = a[1] + 1e-8f, az = a[2] + 1e-8f;
  float ang = sqrtf(ax * ax + ay * ay + az * az);
  float nx = a[0] / ang, ny = a[1] / ang, nz = a[2] / ang;
  float h = ang * 0.5f;
  float c = cosf(h), s = sinf(h);
  float qw = c, qx = s * nx, qy = s * ny, qz = s * nz;
  float qn = sqrtf(qw * qw + qx * qx + qy * qy + qz * qz);
  float w = qw / qn, x = qx / qn, y = qy / qn, z = qz / qn;
  float w2 = w * w, x2 = x * x, y2 = y * y, z2 = z * z;
  float wx = w * x, wy = w * y, wz = w * z, xy = x * y, xz = x * z, yz = y * z;
  R[0] = w2 + x2 - y2 - z2; R[1] = 2 * xy - 2 * wz; R[2] = 2 * wy + 2 * xz;
  R[3] = 2 * wz + 2 * xy; R[4] = w2 - x2 + y2 - z2; R[5] = 2 * yz - 2 * wx;
  R[6] = 2 * xz - 2 * wy; R[7] = 2 * wx + 2 * yz; R[8] = w2 - x2 - y2 + z2;
}

// Top eigenvector of a symmetric 4x4 matrix by cyclic Jacobi sweeps (replaces torch.linalg.eigh at
// lib/utils/transform_fn.py:123); result sign-normalised to w > 0 like :124.  A is row-major, destroyed.
template <typename T>
__device__ __forceinline__ void sym4_top_eigvec(T* A, T* q) {
  T V[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) V[i] = (i % 5 == 0) ? T(1) : T(0);
  const T tiny = sizeof(T) == 8 ? T(1e-300) : T(1e-37);
  for (int sweep = 0; sweep < 30; ++sweep) {
    T off = T(0), diag = T(0);
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      diag += A[p * 4 + p] * A[p * 4 + p];
#pragma unroll
      for (int r = p + 1; r < 4; ++r) off += A[p * 4 + r] * A[p * 4 + r];
    }
    const T epsq = sizeof(T) == 8 ? T(1e-32) : T(1e-15);
    if (off <= epsq * diag || off < tiny) break;
    // p, r, k loops fully unrolled: A and V stay in registers (rolled, their dynamic indexing puts them in local memory)
#pragma unroll
    for (int p = 0; p < 3; ++p)
#pragma unroll
      for (int r = p + 1; r < 4; ++r) {
        T apq = A[p * 4 + r];
        if (v_abs(apq) < tiny) continue;
        T app = A[p * 4 + p], aqq = A[r * 4 + r];
        T tau = (aqq - app) / (T(2) * apq);
        T t = (tau >= T(0) ? T(1) : T(-1)) / (v_abs(tau) + v_sqrt(T(1) + tau * tau));
        T c = T(1) / v_sqrt(T(1) + t * t), s = t * c;
#pragma unroll
        for (int k = 0; k < 4; ++k) {   // columns p, r of A
          T akp = A[k * 4 + p], akq = A[k * 4 + r];
          A[k * 4 + p] = c * akp - s * akq;
          A[k * 4 + r] = s * akp + c * akq;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {   // rows p, r of A
          T apk = A[p * 4 + k], aqk = A[r * 4 + k];
          A[p * 4 + k] = c * apk - s * aqk;
          A[r * 4 + k] = s * apk + c * aqk;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          T vkp = V[k * 4 + p], vkq = V[k * 4 + r];
          V[k * 4 + p] = c * vkp - s * vkq;
          V[k * 4 + r] = s * vkp + c * vkq;
        }
      }
  }
  // arg-max of the diagonal (first maximum) selected without dynamic indexing
  T bv = A[0];
  T vb[4] = {V[0], V[4], V[8], V[12]};
#pragma unroll
  for (int i = 1; i < 4; ++i)
    if (A[i * 4 + i] > bv) {
      bv = A[i * 4 + i];
#pragma unroll
      for (int k = 0; k < 4; ++k) vb[k] = V[k * 4 + i];
    }
  T sgn = vb[0] > T(0) ? T(1) : T(-1);   // ((q_w > 0) - 0.5) * 2
#pragma unroll
  for (int k = 0; k < 4; ++k) q[k] = sgn * vb[k];
}

// average_quaternion over n quaternions (stride given) with optional weights (nullptr = ones).
// Mirrors transform_fn.py:115-124: orient to w>0, A = sum_n w_n q q^T / sum w, top eigenvector, orient.
template <typename T, typename TW>
__device__ __forceinline__ void average_quaternion(const T* Q, int n, const TW* W, T* out) {
  T A[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) A[i] = T(0);
  T wsum = T(0);
  for (int k = 0; k < n; ++k) {
    const T* qk = Q + 4 * k;
    T s = qk[0] > T(0) ? T(1) : T(-1);
    T w = W ? T(W[k]) : T(1);
    wsum += w;
    T o[4] = {s * qk[0], s * qk[1], s * qk[2], s * qk[3]};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) A[i * 4 + j] += (o[i] * o[j]) * w;
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) A[i] /= wsum;
  sym4_top_eigvec(A, out);
}

}  // namespace vpho
