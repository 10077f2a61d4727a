// inverse.cu -- the inverse map of the path: meshes -> per-triangle deformation gradients.
// Replaces TriangleDeformation::getDeformationGradients / getDeformationMatrix
// (reference deformation/cpp/src/deform_triangle_impl.hpp:144-213, :313-380) with _getTransform (:443-446),
// _getGradFromMat (:448-470: polar decomposition through an SVD) and rotation_log_exp::log
// (rotation/utils_rotation.cpp:71-175).  One thread per (frame, triangle), all arithmetic in float64 like the
// reference; embarrassingly parallel, so the only design point is coalesced index/vertex gathers.
//
// Polar decomposition without an SVD routine: C = T^T T = V diag(l) V^T by cyclic Jacobi, s_i = sqrt(l_i)
// sorted descending (JacobiSVD's order), d = sign(det T) = det(U V^T):
//     scale = V diag(s1, s2, d s3) V^T          (= V Temp S V^T, impl.hpp:457)
//     R     = T V diag(1/s1, 1/s2, d/s3) V^T    (= U Temp V^T,   impl.hpp:456)
#include "device_plan.hpp"

#include <algorithm>

namespace sdfa {

namespace {

struct M3 { double m[3][3]; };

__device__ inline M3 mul(const M3 &a, const M3 &b) {
    M3 r;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) r.m[i][j] = a.m[i][0] * b.m[0][j] + a.m[i][1] * b.m[1][j] + a.m[i][2] * b.m[2][j];
    return r;
}
__device__ inline M3 transpose(const M3 &a) {
    M3 r;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r.m[i][j] = a.m[j][i];
    return r;
}
__device__ inline double det3(const M3 &a) {
    return a.m[0][0] * (a.m[1][1] * a.m[2][2] - a.m[1][2] * a.m[2][1]) - a.m[0][1] * (a.m[1][0] * a.m[2][2] - a.m[1][2] * a.m[2][0]) +
           a.m[0][2] * (a.m[1][0] * a.m[2][1] - a.m[1][1] * a.m[2][0]);
}
__device__ inline M3 inverse3(const M3 &a) {          // cofactor formula, like Eigen's fixed-size 3x3 inverse
    const double d = 1.0 / det3(a);
    M3 r;
    r.m[0][0] = (a.m[1][1] * a.m[2][2] - a.m[1][2] * a.m[2][1]) * d;
    r.m[0][1] = (a.m[0][2] * a.m[2][1] - a.m[0][1] * a.m[2][2]) * d;
    r.m[0][2] = (a.m[0][1] * a.m[1][2] - a.m[0][2] * a.m[1][1]) * d;
    r.m[1][0] = (a.m[1][2] * a.m[2][0] - a.m[1][0] * a.m[2][2]) * d;
    r.m[1][1] = (a.m[0][0] * a.m[2][2] - a.m[0][2] * a.m[2][0]) * d;
    r.m[1][2] = (a.m[0][2] * a.m[1][0] - a.m[0][0] * a.m[1][2]) * d;
    r.m[2][0] = (a.m[1][0] * a.m[2][1] - a.m[1][1] * a.m[2][0]) * d;
    r.m[2][1] = (a.m[0][1] * a.m[2][0] - a.m[0][0] * a.m[2][1]) * d;
    r.m[2][2] = (a.m[0][0] * a.m[1][1] - a.m[0][1] * a.m[1][0]) * d;
    return r;
}

// _getEdge3 (impl.hpp:152-161): third edge = scaled normal; false for (nearly) collinear edges
__device__ inline bool edge3(const double *e1, const double *e2, double eps, double *e3) {
    e3[0] = e1[1] * e2[2] - e1[2] * e2[1];
    e3[1] = e1[2] * e2[0] - e1[0] * e2[2];
    e3[2] = e1[0] * e2[1] - e1[1] * e2[0];
    const double len1 = sqrt(e1[0] * e1[0] + e1[1] * e1[1] + e1[2] * e1[2]);
    const double len2 = sqrt(e2[0] * e2[0] + e2[1] * e2[1] + e2[2] * e2[2]);
    const double abs_cos = fabs((e1[0] * e2[0] + e1[1] * e2[1] + e1[2] * e2[2]) / (len1 * len2));
    if (abs_cos > 1.0 - eps) return false;          // NaN compares false: "good", as in the reference
    const double s = fmax(sqrt(sqrt(e3[0] * e3[0] + e3[1] * e3[1] + e3[2] * e3[2])), eps);
    e3[0] /= s; e3[1] /= s; e3[2] /= s;
    return true;
}

// symmetric 3x3 eigen-decomposition by cyclic Jacobi: c = v diag(l) v^T
__device__ inline void jacobi_eig3(M3 c, M3 &v, double *l) {
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) v.m[i][j] = (i == j) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 12; ++sweep) {
        const double off = c.m[0][1] * c.m[0][1] + c.m[0][2] * c.m[0][2] + c.m[1][2] * c.m[1][2];
        const double diag = c.m[0][0] * c.m[0][0] + c.m[1][1] * c.m[1][1] + c.m[2][2] * c.m[2][2];
        if (off <= 1e-32 * diag) break;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                if (c.m[p][q] == 0.0) continue;
                const double theta = (c.m[q][q] - c.m[p][p]) / (2.0 * c.m[p][q]);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double cs = 1.0 / sqrt(t * t + 1.0), sn = t * cs;
                for (int k = 0; k < 3; ++k) {               // c <- c * J
                    const double a = c.m[k][p], b = c.m[k][q];
                    c.m[k][p] = cs * a - sn * b; c.m[k][q] = sn * a + cs * b;
                }
                for (int k = 0; k < 3; ++k) {               // c <- J^T * c
                    const double a = c.m[p][k], b = c.m[q][k];
                    c.m[p][k] = cs * a - sn * b; c.m[q][k] = sn * a + cs * b;
                }
                for (int k = 0; k < 3; ++k) {               // v <- v * J
                    const double a = v.m[k][p], b = v.m[k][q];
                    v.m[k][p] = cs * a - sn * b; v.m[k][q] = sn * a + cs * b;
                }
            }
    }
    l[0] = c.m[0][0]; l[1] = c.m[1][1]; l[2] = c.m[2][2];
}

// rotation_log_exp::log(Matrix3d) (utils_rotation.cpp:71-175): entries (0,1), (0,2), (1,2) of angle * cross(axis)
__device__ inline void rotation_log(const M3 &R, double *l01, double *l02, double *l12) {
    const double tol = 1.0e-6;
    *l01 = *l02 = *l12 = 0.0;
    double nrm = 0.0;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            const double d = R.m[0][i] * R.m[0][j] + R.m[1][i] * R.m[1][j] + R.m[2][i] * R.m[2][j] - (i == j ? 1.0 : 0.0);
            nrm += d * d;
        }
    if (sqrt(nrm) > tol) return;                           // :73-77 (reference leaves the result undefined)
    double csin = (R.m[0][0] + R.m[1][1] + R.m[2][2] - 1.0) / 2.0;
    if (csin < -1.0 || csin > 1.0) {
        if (fabs(csin - 1.0) > tol && fabs(csin + 1.0) > tol) return;
        csin = fmax(fmin(1.0, csin), -1.0);
    }
    double angle = acos(csin);
    double ax[3];
    if (fabs(angle) < tol) return;
    const double PI = 3.14159265358979323846;
    if (fabs(angle - PI) < tol) {                          // :96-108
        const double b00 = (R.m[0][0] + 1.0) / 2.0, b11 = (R.m[1][1] + 1.0) / 2.0, b22 = (R.m[2][2] + 1.0) / 2.0;
        const double b01 = R.m[0][1] / 2.0, b02 = R.m[0][2] / 2.0;
        ax[0] = sqrt(b00);
        ax[1] = (ax[0] * b01 > 0.0) ? sqrt(b11) : -sqrt(b11);
        ax[2] = (ax[0] * b02 > 0.0) ? sqrt(b22) : -sqrt(b22);
        angle = PI;
    } else {
        const double tx[3] = {R.m[2][1] - R.m[1][2], R.m[0][2] - R.m[2][0], R.m[1][0] - R.m[0][1]};
        for (int attempt = 0; attempt < 2; ++attempt) {
            const double sinv = sin(angle);
            for (int k = 0; k < 3; ++k) ax[k] = tx[k] / (2.0 * sinv);
            const double k1 = 1.0 - csin;
            const double r01 = k1 * ax[0] * ax[1] - ax[2] * sinv, r02 = k1 * ax[0] * ax[2] + ax[1] * sinv;
            const double r10 = k1 * ax[0] * ax[1] + ax[2] * sinv, r12 = k1 * ax[1] * ax[2] - ax[0] * sinv;
            const double r20 = k1 * ax[0] * ax[2] - ax[1] * sinv, r21 = k1 * ax[1] * ax[2] + ax[0] * sinv;
            const double chk = (R.m[0][1] - r01) * (R.m[0][1] - r01) + (R.m[0][2] - r02) * (R.m[0][2] - r02) +
                               (R.m[1][0] - r10) * (R.m[1][0] - r10) + (R.m[1][2] - r12) * (R.m[1][2] - r12) +
                               (R.m[2][0] - r20) * (R.m[2][0] - r20) + (R.m[2][1] - r21) * (R.m[2][1] - r21);
            if (chk < tol || attempt == 1) break;
            angle = 2.0 * PI - angle;                      // :133-152
        }
    }
    // angle * (temp - temp^T), temp(2,1)=a0, temp(0,2)=a1, temp(1,0)=a2  (:161-166)
    *l01 = -angle * ax[2];
    *l02 = angle * ax[1];
    *l12 = -angle * ax[0];
}

template <typename OutT>
__global__ void k_deform_grad(const float *__restrict__ va, const float *__restrict__ vb, long long vb_stride,
                              const uint32_t *__restrict__ tris, int n_tris, int n_frames, double eps, int as_matrix,
                              OutT *__restrict__ out) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)n_tris * n_frames) return;
    const int j = (int)(idx % n_tris);
    const int f = (int)(idx / n_tris);
    const uint32_t i1 = tris[3 * j] * 3, i2 = tris[3 * j + 1] * 3, i3 = tris[3 * j + 2] * 3;
    const float *pb = vb + (long long)f * vb_stride;
    double ea1[3], ea2[3], ea3[3], eb1[3], eb2[3], eb3[3];
    for (int d = 0; d < 3; ++d) {
        ea1[d] = (double)va[i2 + d] - (double)va[i1 + d];
        ea2[d] = (double)va[i3 + d] - (double)va[i1 + d];
        eb1[d] = (double)pb[i2 + d] - (double)pb[i1 + d];
        eb2[d] = (double)pb[i3 + d] - (double)pb[i1 + d];
    }
    const bool good_a = edge3(ea1, ea2, eps, ea3), good_b = edge3(eb1, eb2, eps, eb3);
    OutT *o = out + idx * 9;
    if (!(good_a && good_b)) {                             // impl.hpp:198-201 / :366-369
        for (int k = 0; k < 9; ++k) o[k] = (OutT)((as_matrix && (k == 0 || k == 4 || k == 8)) ? 1.0 : 0.0);
        return;
    }
    M3 A, B;
    for (int d = 0; d < 3; ++d) {
        A.m[d][0] = ea1[d]; A.m[d][1] = ea2[d]; A.m[d][2] = ea3[d];
        B.m[d][0] = eb1[d]; B.m[d][1] = eb2[d]; B.m[d][2] = eb3[d];
    }
    const M3 T = mul(B, inverse3(A));                      // _getTransform, impl.hpp:443-446
    if (as_matrix) {
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) o[3 * r + c] = (OutT)T.m[r][c];
        return;
    }
    M3 V;
    double l[3];
    jacobi_eig3(mul(transpose(T), T), V, l);
    // sort descending (JacobiSVD's singular value order)
    int ord[3] = {0, 1, 2};
    if (l[ord[0]] < l[ord[1]]) { int t = ord[0]; ord[0] = ord[1]; ord[1] = t; }
    if (l[ord[1]] < l[ord[2]]) { int t = ord[1]; ord[1] = ord[2]; ord[2] = t; }
    if (l[ord[0]] < l[ord[1]]) { int t = ord[0]; ord[0] = ord[1]; ord[1] = t; }
    const double dsgn = det3(T) < 0.0 ? -1.0 : 1.0;
    double s[3], sinv[3];
    for (int k = 0; k < 3; ++k) {
        s[k] = sqrt(fmax(l[ord[k]], 0.0));
        sinv[k] = 1.0 / s[k];
    }
    s[2] *= dsgn; sinv[2] *= dsgn;
    M3 scale, sci;
    for (int i = 0; i < 3; ++i)
        for (int jj = 0; jj < 3; ++jj) {
            double a = 0.0, b = 0.0;
            for (int k = 0; k < 3; ++k) {
                const double vv = V.m[i][ord[k]] * V.m[jj][ord[k]];
                a += vv * s[k];
                b += vv * sinv[k];
            }
            scale.m[i][jj] = a;
            sci.m[i][jj] = b;
        }
    const M3 R = mul(T, sci);
    double l01, l02, l12;
    rotation_log(R, &l01, &l02, &l12);
    o[0] = (OutT)(scale.m[0][0] - 1.0); o[1] = (OutT)scale.m[0][1]; o[2] = (OutT)scale.m[0][2];
    o[3] = (OutT)(scale.m[1][1] - 1.0); o[4] = (OutT)scale.m[1][2]; o[5] = (OutT)(scale.m[2][2] - 1.0);
    o[6] = (OutT)l01; o[7] = (OutT)l02; o[8] = (OutT)l12;
}

}  // namespace

cudaError_t launch_deform_grad(const float *verts_a, const float *verts_b, long long vb_stride, const uint32_t *tris,
                               int n_tris, int n_frames, double eps, int as_matrix, void *out, bool out_f64,
                               cudaStream_t stream) {
    const long long total = (long long)n_tris * n_frames;
    if (total <= 0) return cudaSuccess;
    const unsigned grid = (unsigned)((total + 127) / 128);
    if (out_f64)
        k_deform_grad<double><<<grid, 128, 0, stream>>>(verts_a, verts_b, vb_stride, tris, n_tris, n_frames, eps, as_matrix, (double *)out);
    else
        k_deform_grad<float><<<grid, 128, 0, stream>>>(verts_a, verts_b, vb_stride, tris, n_tris, n_frames, eps, as_matrix, (float *)out);
    count_launch();
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// Batched saber.stream.seek (saber/data/stream/stream.py:20-46): the bracketing rows and the weight come from
// the host (api.cpp restates the reference's binary search); the blend is evaluated in float64 like numpy does
// with a float64 weight, then rounded to float32 (what frame_to_mesh casts to, viewer/frame.py:112).
__global__ void k_seek(const float *__restrict__ seq, long long width, const int2 *__restrict__ pairs,
                       const double *__restrict__ weights, float *__restrict__ out) {
    const int q = blockIdx.y;
    const int2 pr = pairs[q];
    const double a = weights[q];
    const float *lo = seq + (long long)pr.x * width, *hi = seq + (long long)pr.y * width;
    float *dst = out + (long long)q * width;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < width; i += (long long)gridDim.x * blockDim.x)
        dst[i] = pr.x == pr.y ? lo[i] : (float)__dadd_rn(__dmul_rn(a, (double)lo[i]), __dmul_rn(1.0 - a, (double)hi[i]));   // no FMA: numpy's rounding
}

cudaError_t launch_seek(const float *seq, long long width, const int2 *pairs, const double *weights, int n_query, float *out,
                        cudaStream_t stream) {
    if (n_query <= 0 || width <= 0) return cudaSuccess;
    const unsigned bx = (unsigned)std::min<long long>((width + 255) / 256, 64);
    k_seek<<<dim3(bx, (unsigned)n_query), 256, 0, stream>>>(seq, width, pairs, weights, out);
    count_launch();
    return cudaGetLastError();
}

}  // namespace sdfa
