/* sdfa_b200.h -- C ABI of the B200-native dgrad -> mesh reconstruction path.
 *
 * Drop-in boundary for the ONE hot path of chaiyujin/sdfa-2019 that this library replaces:
 * `deformation.get_mesh` and friends (reference: deformation/cpp/src/pybind.cpp:129-153, which
 * forward to deformation::TriangleDeformation, deformation/cpp/src/deform_triangle.hpp:12-85).
 * Plain pointers and sizes only; no torch / pybind / Eigen types.  All compute entry points
 * run hand-written sm_100a CUDA kernels; there is no CPU fallback -- if no CUDA device is
 * usable they return SDFA_ERR_CUDA and sdfa_last_error() says why.
 *
 * Conventions
 *   - every function returns an int status (SDFA_OK == 0) unless stated otherwise;
 *     sdfa_last_error() returns a thread-local, human-readable message for the last failure
 *     (the reference instead logs and calls exit(1): deformation/cpp/src/log.hpp:32-33)
 *   - "dev" pointers are CUDA device pointers on the handle's device, "host" pointers are
 *     ordinary host memory; `stream` is a cudaStream_t passed as void* (NULL = default stream)
 *   - dgrad layout: [N, n_src_tris, 9] row-major, per triangle
 *     [s00,s01,s02,s11,s12,s22,r01,r02,r12]  (deform_triangle_impl.hpp:232-240;
 *     speech_anime/model/model.py:246-257)
 *   - vertex layout: [N, n_verts, 3] float32 row-major (pybind.cpp:108)
 */
#ifndef SDFA_B200_H_
#define SDFA_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SDFA_OK            0
#define SDFA_ERR_ARG       1   /* bad argument (shape, index out of range, NULL, duplicate constraint) */
#define SDFA_ERR_FACTOR    2   /* A^T A + reg*I not positive definite (reference: setStaticTarget returns false) */
#define SDFA_ERR_CUDA      3   /* CUDA runtime failure or no usable device */
#define SDFA_ERR_STATE     4   /* call order (e.g. decode before sdfa_set_pca) */
#define SDFA_ERR_UNSUPPORTED 5

typedef struct sdfa_handle sdfa_handle;   /* opaque: one template + factor + device plan */

/* ---- lifetime -------------------------------------------------------------------------- */

/* Replaces TriangleDeformation::setStaticTarget (deform_triangle_impl.hpp:7-142; marshalled by
 * SetTarget, pybind.cpp:13-33).  Builds A, A^T A + reg*I, its fp64 Cholesky factor and the device
 * schedules, uploads them to CUDA device `device`.
 *   verts      [n_verts,3] float32        tris  [n_tris,3] uint32
 *   cnsts      [n_cnsts] uint32 or NULL   corr_count [n_tris] uint32 or NULL (deform_triangle_impl.hpp:18-22)
 * `device` < 0 builds the host-side analysis only (no CUDA call is made; compute entry points then
 * fail with SDFA_ERR_CUDA) -- used by CPU-only tests of the host logic. */
int sdfa_create(sdfa_handle **out, const float *verts, int n_verts, const uint32_t *tris, int n_tris,
                const uint32_t *cnsts, int n_cnsts, const uint32_t *corr_count, double reg, int device);
/* Same, with construction options as "key=value;key=value" (NULL or "" = defaults):
 *   solver=auto|simt|tensor   which solve kernel (auto: the tcgen05 block solve when the template fits, else SIMT sweeps)
 *   pipe_chunk=N              frames per chunk of large batches (0 = never chunk); see sdfa_set_option
 *   frames_per_tile, asm_rows, ts_leaf   tuning knobs of the planners
 *   asm_gather=1|2|3         assembly kernel for reference-layout dgrad (3: TMA tensor-map boxes; 2: per-warp cp.async rings;
 *                            1: first generation; default: 3 for rows up to 1 MB, else 2)
 *   decode=f16|tf32          decode kernel: FP16 hi/lo split (default when the basis widths fit) or the TF32 split
 *   output=1|2, output_frames=8|16|32   output kernel generation and frames per CTA (default 2, 8)
 * An unknown key is SDFA_ERR_ARG.  (The SDFA_* environment variables of the same names are read at creation as
 * defaults for A/B runs; an option given here wins.) */
int sdfa_create_with(sdfa_handle **out, const float *verts, int n_verts, const uint32_t *tris, int n_tris,
                     const uint32_t *cnsts, int n_cnsts, const uint32_t *corr_count, double reg, int device,
                     const char *options);
void sdfa_destroy(sdfa_handle *h);

/* Threading and streams (SURVEY 8b "Threading"; the reference's singleton is neither thread-safe nor re-entrant):
 *   - a handle may be shared by several host threads: every entry point holds the handle's lock while it enqueues;
 *   - *_dev entry points are stream-ordered and asynchronous; calls on different streams are ordered on the device in
 *     call order (they share the handle's workspaces), and every setter / *_host / legacy call first waits for
 *     stream-ordered work still in flight;
 *   - the caller's current CUDA device is left unchanged. */

/* Run-time options: "pipe_chunk" = frames per chunk of large batches (-1 automatic, 0 never chunk). */
int sdfa_set_option(sdfa_handle *h, const char *name, long long value);

/* Counts; any out pointer may be NULL.  n_eq = number of equation blocks, n_active = blocks that touch
 * a free vertex, nnz_l = nonzeros of the Cholesky factor.  (IsSame, pybind.cpp:119-126, compares the
 * first three.) */
int sdfa_info(const sdfa_handle *h, int *n_verts, int *n_tris, int *n_cnsts, int *n_free, int *n_eq,
              int *n_active, long long *nnz_l);

const char *sdfa_last_error(void);

/* ---- constraint positions --------------------------------------------------------------- */

/* Positions of the constrained vertices used by the following reconstruct calls
 * (`_cnst_verts`, deform_triangle_impl.hpp:272-283, :302-308).  [n_cnsts,3] float32 host, or NULL to
 * use the template's own positions (the way viewer/frame.py:130-132 calls it).  Recomputes the fp64
 * base solution on the host. */
int sdfa_set_constraint_positions(sdfa_handle *h, const float *cnst_verts_host);

/* Correspondence mode of getMeshFromDeformationGradients (deform_triangle_impl.hpp:254-268):
 * corr_count [n_tris], corr_faces [n_eq] uint32 host arrays; corr_count == NULL switches back to
 * "block k reads source triangle k".  n_src_tris = triangles per source dgrad row. */
int sdfa_set_correspondences(sdfa_handle *h, const uint32_t *corr_count, const uint32_t *corr_faces,
                             int n_src_tris);

/* ---- reconstruction: dgrad -> vertices (getMeshFromDeformationGradients, impl.hpp:215-310) ---- */

/* Batched, device buffers, stream-ordered.  dgrad_dev [n_frames, n_src_tris*9] float32 with row stride
 * `dgrad_stride` floats (0 = dense); out_dev [n_frames, n_verts, 3] float32. */
int sdfa_reconstruct_dev(sdfa_handle *h, const float *dgrad_dev, long long dgrad_stride, int n_frames,
                         float *out_dev, void *stream);

/* Batched, HOST buffers (float32 in, float32 out): copies in, runs the kernels, copies out, synchronises. */
int sdfa_reconstruct_host(sdfa_handle *h, const float *dgrad_host, int n_frames, float *out_host);

/* The legacy single-frame call: float64 dgrad in, float32 vertices out (GetMeshFromGrad,
 * pybind.cpp:101-117).  cnst_verts / corr_* as in the reference call; NULL = not given. */
int sdfa_get_mesh_f64(sdfa_handle *h, const double *dgrad_host, long long dgrad_len,
                      const float *cnst_verts_host, const uint32_t *corr_count, const uint32_t *corr_faces,
                      long long corr_faces_len, float *out_host);

/* Raw-matrix variant (getMeshFromDeformationMatrix, impl.hpp:382-440; GetMeshFromMat pybind.cpp:60-74):
 * dmat [n_tris,9] float64 row-major T per triangle. */
int sdfa_get_mesh_from_dm_f64(sdfa_handle *h, const double *dmat_host, long long dmat_len,
                              const float *cnst_verts_host, float *out_host);

/* ---- PCA decode (PcaInversion.forward x2 + interleave: output_module.py:115-116, model.py:246-257) ---- */

/* compT_scale [n_tris*6, k_scale], means_scale [n_tris*6], compT_rotat [n_tris*3, k_rotat],
 * means_rotat [n_tris*3]; float32 host arrays, copied (only active triangles' rows are kept on device). */
int sdfa_set_pca(sdfa_handle *h, const float *compT_scale, const float *means_scale, int k_scale,
                 const float *compT_rotat, const float *means_rotat, int k_rotat);

/* coefficients -> vertices, device buffers: coeff_scale_dev [n_frames,k_scale], coeff_rotat_dev
 * [n_frames,k_rotat] float32; out_dev [n_frames,n_verts,3]. */
int sdfa_decode_reconstruct_dev(sdfa_handle *h, const float *coeff_scale_dev, const float *coeff_rotat_dev,
                                int n_frames, float *out_dev, void *stream);
int sdfa_decode_reconstruct_host(sdfa_handle *h, const float *coeff_scale_host, const float *coeff_rotat_host,
                                 int n_frames, float *out_host);

/* decode only: writes the full reference-layout dgrad [n_frames, n_tris*9] (inactive triangles included)
 * -- the tensor data_to_anime_feat returns; n_tris here is the source triangle count sdfa_set_pca was called with
 * (SDFA_ERR_STATE if correspondences with another count were set since). */
int sdfa_decode_dgrad_dev(sdfa_handle *h, const float *coeff_scale_dev, const float *coeff_rotat_dev,
                          int n_frames, float *dgrad_dev, void *stream);

/* decode only, into the internal frame-tiled compact layout the assembly kernel consumes (what
 * sdfa_decode_reconstruct_* produces with the tcgen05 kernel): [ceil(n_frames/64), slots, 64] float32
 * (tile of 64 frames, slot, frame inside the tile).
 * sdfa_compact_layout returns `slots` and copies the map slot -> source_triangle*9 + component
 * (-1 = nothing decoded there); map may be NULL. */
int sdfa_decode_compact_dev(sdfa_handle *h, const float *coeff_scale_dev, const float *coeff_rotat_dev,
                            int n_frames, float *dgrad_compact_dev, void *stream);
int sdfa_compact_layout(const sdfa_handle *h, int32_t *map, int cap);

/* ---- free rows only (opt-in; not in the reference) ----------------------------------------------------------
 * The constrained vertices of every reconstructed frame are constants (the positions given to
 * sdfa_set_constraint_positions, copied through by impl.hpp:302-308) -- 3762 of FLAME's 5023 rows with the default
 * mask.  The *_free variants write only the free vertices, [n_frames, n_free, 3] float32 in ascending vertex order
 * (sdfa_free_vertices gives the vertex index of each row), so a device->host copy or an NCCL gather moves 15 KB
 * instead of 60 KB per FLAME frame; sdfa_expand_free_dev rebuilds the reference layout on the receiving device. */
int sdfa_free_vertices(const sdfa_handle *h, int32_t *ids, int cap);     /* returns n_free; ids may be NULL */
int sdfa_reconstruct_free_dev(sdfa_handle *h, const float *dgrad_dev, long long dgrad_stride, int n_frames,
                              float *out_free_dev, void *stream);
int sdfa_reconstruct_free_host(sdfa_handle *h, const float *dgrad_host, int n_frames, float *out_free_host);
int sdfa_decode_reconstruct_free_dev(sdfa_handle *h, const float *coeff_scale_dev, const float *coeff_rotat_dev,
                                     int n_frames, float *out_free_dev, void *stream);
int sdfa_decode_reconstruct_free_host(sdfa_handle *h, const float *coeff_scale_host, const float *coeff_rotat_host,
                                      int n_frames, float *out_free_host);
/* free_dev [n_frames, n_free, 3] -> out_dev [n_frames, n_verts, 3] with the handle's current constraint positions. */
int sdfa_expand_free_dev(sdfa_handle *h, const float *free_dev, int n_frames, float *out_dev, void *stream);

/* ---- inverse path: meshes -> dgrad (getDeformationGradients, impl.hpp:144-213) ----------- */

/* Stateless like the reference (GetDeformGrad, pybind.cpp:78-99): verts_a/verts_b [n_verts,3] float32,
 * tris [n_tris,3] uint32 host; out [n_tris*9] float64 host.  as_matrix != 0 returns row-major T
 * instead (getDeformationMatrix, impl.hpp:313-380).  Runs on CUDA device `device`. */
int sdfa_get_deform_grad_host(const float *verts_a, const float *verts_b, int n_verts, const uint32_t *tris,
                              int n_tris, double eps, int as_matrix, int device, double *out_host);

/* Batched, device buffers: verts_a_dev [n_verts,3] (the template), verts_b_dev [n_frames,n_verts,3] float32,
 * tris_dev [n_tris,3] uint32; out_dev [n_frames,n_tris,9] float32 -- the dtype datasets store
 * (generate_dgrad, speech_anime/datasets/vocaset/preload.py:765-835).  Indices are not range-checked. */
int sdfa_deform_grad_batch_dev(const float *verts_a_dev, const float *verts_b_dev, int n_verts, const uint32_t *tris_dev,
                               int n_tris, int n_frames, double eps, int as_matrix, float *out_dev, void *stream);

/* ---- the step before the path: time interpolation of a sequence (saber.stream.seek) ------- */

/* Batched saber.stream.seek (saber/data/stream/stream.py:20-46): for every query time the two bracketing rows of
 * seq_dev [n_src, width] float32 are blended, out = a*seq[m] + (1-a)*seq[m+1] with a = (t[m+1]-ts)/(t[m+1]-t[m])
 * evaluated in float64 like the reference and rounded to float32; a query outside [t[0], t[-1]] copies the row the
 * reference's binary search stops at.  timestamps_host [n_src] (ascending) and query_host [n_query] are host
 * float64; out_dev [n_query, width].  The decode is affine, so seeking PCA coefficients and decoding equals
 * decoding and seeking the dgrad (model.py:201-212 does the latter on the host, one frame at a time). */
int sdfa_seek_dev(const float *seq_dev, int n_src, long long width, const double *timestamps_host,
                  const double *query_host, int n_query, float *out_dev, void *stream);

/* ---- measurement / introspection -------------------------------------------------------- */

/* Kernel launches issued by this library since process start (for bench.py's gpu_launches). */
long long sdfa_launch_count(void);

/* Per-stage device time of the most recent *_dev / *_host reconstruct call when timing is enabled
 * with sdfa_set_timing(h, 1): ms[0]=decode, ms[1]=assembly, ms[2]=solve, ms[3]=output.  Enabling
 * timing adds event records + a synchronise per call. */
int sdfa_set_timing(sdfa_handle *h, int enable);
int sdfa_last_timing(const sdfa_handle *h, float ms[4]);

/* Host-side plan introspection for tests (no CUDA needed).  `what` selects a buffer; returns its size in
 * bytes, copies min(size, cap) bytes into dst when dst != NULL.  See csrc/plan.hpp for the list. */
long long sdfa_debug_get(const sdfa_handle *h, const char *what, void *dst, long long cap);

#ifdef __cplusplus
}
#endif
#endif /* SDFA_B200_H_ */
