// TEST INFRASTRUCTURE ONLY -- not part of the product.
//
// C-ABI shim around the UNMODIFIED reference solver so that the checker and the
// CPU baseline can (a) be loaded with ctypes and (b) run one independent
// deformation::TriangleDeformation instance per host thread.  (The reference's
// own pybind module keeps a single process-global instance --
// deformation/cpp/src/pybind.cpp:10 -- which forces process-level parallelism.)
//
// Nothing is copied from the reference: this file only #includes
// /root/reference/deformation/cpp/src/deform.hpp where it lies and forwards to
//   TriangleDeformation::setStaticTarget                 (deform_triangle_impl.hpp:7-142)
//   TriangleDeformation::getMeshFromDeformationGradients (deform_triangle_impl.hpp:215-310)
//   TriangleDeformation::getDeformationGradients         (deform_triangle_impl.hpp:144-213)
//   TriangleDeformation::getDeformationMatrix            (deform_triangle_impl.hpp:313-380)
//   TriangleDeformation::getMeshFromDeformationMatrix    (deform_triangle_impl.hpp:382-440)
// with the same template arguments the pybind layer instantiates
// (float verts, uint32 indices, double dgrad; pybind.cpp:13-117).
//
// Built by oracle/Makefile into oracle/_ref/libsdfa_ref.so (git-ignored).
#include <cstdint>
#include <cstring>
#include <chrono>
#include <thread>
#include <vector>
#include "deform.hpp"

using deformation::TriangleDeformation;

extern "C" {

void *sdfa_ref_create() { return new TriangleDeformation(); }
void sdfa_ref_destroy(void *h) { delete static_cast<TriangleDeformation *>(h); }

int sdfa_ref_set_target(void *h, const float *verts, int n_verts, const uint32_t *tris, int n_tris,
                        const uint32_t *cnsts, int n_cnsts, const uint32_t *corr_count, double reg) {
    auto *d = static_cast<TriangleDeformation *>(h);
    return d->setStaticTarget<float, uint32_t>(verts, (size_t)n_verts, tris, (size_t)n_tris, cnsts,
                                               (size_t)n_cnsts, corr_count, reg)
               ? 1 : 0;
}

int sdfa_ref_get_mesh(void *h, float *out_verts, const double *dgrad, const float *cnst_verts,
                      const uint32_t *corr_count, const uint32_t *corr_faces) {
    auto *d = static_cast<TriangleDeformation *>(h);
    return d->getMeshFromDeformationGradients<double, float, uint32_t>(out_verts, dgrad, cnst_verts,
                                                                       corr_count, corr_faces)
               ? 1 : 0;
}

int sdfa_ref_get_mesh_from_dm(void *h, float *out_verts, const double *dmat, const float *cnst_verts) {
    auto *d = static_cast<TriangleDeformation *>(h);
    return d->getMeshFromDeformationMatrix<double, float, uint32_t>(out_verts, dmat, cnst_verts,
                                                                    nullptr, nullptr)
               ? 1 : 0;
}

int sdfa_ref_get_deform_grad(void *h, double *dgrad, const float *verts_a, const float *verts_b,
                             int n_verts, const uint32_t *tris, int n_tris, double eps) {
    auto *d = static_cast<TriangleDeformation *>(h);
    return d->getDeformationGradients<double, float, uint32_t>(dgrad, verts_a, verts_b, (size_t)n_verts,
                                                               tris, (size_t)n_tris, eps)
               ? 1 : 0;
}

int sdfa_ref_get_deform_mat(void *h, double *dmat, const float *verts_a, const float *verts_b,
                            int n_verts, const uint32_t *tris, int n_tris, double eps) {
    auto *d = static_cast<TriangleDeformation *>(h);
    return d->getDeformationMatrix<double, float, uint32_t>(dmat, verts_a, verts_b, (size_t)n_verts,
                                                            tris, (size_t)n_tris, eps)
               ? 1 : 0;
}

// Frame loop of the reference path over a batch, one solver instance per thread.
// Per frame it does what viewer/frame.py:113,134 + pybind.cpp:101-117 do: widen the
// float32 dgrad row to float64, then call getMeshFromDeformationGradients.
// handles[t] must already have had set_target applied.  Returns wall seconds of
// the slowest thread (max over threads), frames are split in contiguous blocks.
double sdfa_ref_get_mesh_batch(void **handles, int n_threads, const float *dgrad_f32, long n_frames,
                               long dgrad_len, const float *cnst_verts, float *out_verts,
                               long out_len) {
    std::vector<double> secs((size_t)n_threads, 0.0);
    auto work = [&](int t) {
        auto *d = static_cast<TriangleDeformation *>(handles[t]);
        long lo = n_frames * t / n_threads, hi = n_frames * (t + 1) / n_threads;
        std::vector<double> row((size_t)dgrad_len);
        auto t0 = std::chrono::steady_clock::now();
        for (long f = lo; f < hi; ++f) {
            const float *src = dgrad_f32 + f * dgrad_len;
            for (long i = 0; i < dgrad_len; ++i) row[(size_t)i] = (double)src[i];
            d->getMeshFromDeformationGradients<double, float, uint32_t>(
                out_verts + f * out_len, row.data(), cnst_verts, nullptr, nullptr);
        }
        secs[(size_t)t] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    };
    if (n_threads <= 1) {
        work(0);
    } else {
        std::vector<std::thread> pool;
        for (int t = 0; t < n_threads; ++t) pool.emplace_back(work, t);
        for (auto &th : pool) th.join();
    }
    double mx = 0.0;
    for (double s : secs) mx = s > mx ? s : mx;
    return mx;
}

}  // extern "C"
