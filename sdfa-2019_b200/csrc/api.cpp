// api.cpp -- the C ABI declared in include/sdfa_b200.h: handle lifetime, uploads, and the host side of
// every entry point.  Mirrors what deformation/cpp/src/pybind.cpp:13-126 does around the reference solver
// (argument checks, singleton state) but reports errors through status codes instead of exit(1).
#include "../../include/sdfa_b200.h"
#include "device_plan.hpp"
#include "plan.hpp"

#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <map>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include <nvtx3/nvToolsExt.h>

using namespace sdfa;

static thread_local std::string g_err;
static int fail(int code, const std::string &msg) { g_err = msg; return code; }
#define CUDA_TRY(expr)                                                                              \
    do {                                                                                            \
        cudaError_t e__ = (expr);                                                                   \
        if (e__ != cudaSuccess)                                                                     \
            return fail(SDFA_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));       \
    } while (0)

// NVTX range per entry point / stage (the reference only has wall-clock prints)
struct NvtxRange {
    explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

struct sdfa_handle {
    HostPlan host;
    DevicePlan dev;
    // One lock per handle: every entry point that takes a handle holds it for the duration of the call (the calls
    // enqueue work and return, so this serialises enqueueing, not the GPU).  Recursive because the legacy entry points
    // are built from the setters.
    std::recursive_mutex mu;
    std::vector<void *> allocs;           // every device allocation, freed in destroy
    std::vector<void *> pca_allocs;       // the PCA buffers of the current sdfa_set_pca (replaced by the next one)
    int pipe_chunk = -1;                  // sdfa_set_option("pipe_chunk"): -1 auto, 0 never chunk, > 0 frames per chunk
    cudaEvent_t ev_async = nullptr;       // recorded after the last stream-ordered call: table updates wait for it
    bool async_pending = false;
    cudaStream_t last_stream = nullptr;
    bool free_only = false;               // output mode of the call in progress (set under the lock by the entry point)
    std::vector<int32_t> eq_src_host;     // current equation -> source triangle map
    int n_src_tris = 0;
    bool has_pca = false, has_full_pca = false;
    // growable device scratch, two sets so that the *_host entry points can overlap the copies of one
    // chunk with the kernels of the next (each set is used on its own stream)
    struct Workspace {
        float *rhs = nullptr; size_t rhs_cap = 0;          // tile-major scratch [tiles][n_free][3][32]
        float *dgrad_c = nullptr; size_t dgrad_c_cap = 0;  // compact decoded dgrad
        float *ximg_s = nullptr; size_t ximg_s_cap = 0;    // split coefficient tile images (tensor-core decode)
        float *ximg_r = nullptr; size_t ximg_r_cap = 0;
        float *io_in = nullptr; size_t io_in_cap = 0;      // staging for *_host entry points
        float *io_in2 = nullptr; size_t io_in2_cap = 0;
        float *io_out = nullptr; size_t io_out_cap = 0;
        cudaStream_t stream = nullptr;
        // chunk pipeline: the output kernel of chunk i runs on `side` under the kernels of chunk i + 1
        float *rhs2 = nullptr; size_t rhs2_cap = 0;
        cudaStream_t side = nullptr;
        cudaEvent_t ev_solved[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
    } ws[2];
    bool timing = false;
    cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    float last_ms[4] = {0, 0, 0, 0};
};

const char *sdfa_last_error(void) { return g_err.c_str(); }

template <typename T>
static int upload(sdfa_handle *h, const std::vector<T> &v, const T **dst) {
    void *p = nullptr;
    size_t bytes = std::max<size_t>(v.size() * sizeof(T), 16);
    CUDA_TRY(cudaMalloc(&p, bytes));
    h->allocs.push_back(p);
    if (!v.empty()) CUDA_TRY(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    *dst = static_cast<const T *>(p);
    return SDFA_OK;
}
template <typename T>
static int upload_mut(sdfa_handle *h, const std::vector<T> &v, T **dst) {
    const T *p = nullptr;
    int rc = upload(h, v, &p);
    *dst = const_cast<T *>(p);
    return rc;
}
static int grow(float **buf, size_t *cap, size_t floats) {
    if (*cap >= floats) return SDFA_OK;
    if (*buf) cudaFree(*buf);
    *buf = nullptr; *cap = 0;
    CUDA_TRY(cudaMalloc((void **)buf, floats * sizeof(float)));
    *cap = floats;
    return SDFA_OK;
}

// Makes the handle's device current for the duration of an entry point and puts the caller's device back afterwards.
struct DeviceGuard {
    int prev = -1;
    bool armed = false;
    int enter(int device) {
        CUDA_TRY(cudaGetDevice(&prev));
        if (prev != device) { CUDA_TRY(cudaSetDevice(device)); armed = true; }
        return SDFA_OK;
    }
    ~DeviceGuard() { if (armed) cudaSetDevice(prev); }
};
static int need_device(sdfa_handle *h, DeviceGuard &g) {
    if (h->dev.device < 0)
        return fail(SDFA_ERR_CUDA, "handle was created without a CUDA device; this library has no CPU path");
    return g.enter(h->dev.device);
}
// Entry-point prologue: NULL check, the handle's lock, its device current (restored on return).
#define ENTER_DEVICE(h)                                                                             \
    if (!(h)) return fail(SDFA_ERR_ARG, "NULL handle");                                             \
    std::lock_guard<std::recursive_mutex> lock__((h)->mu);                                          \
    DeviceGuard guard__;                                                                            \
    if ((rc = need_device((h), guard__))) return rc
// Stream-ordered entry points leave a marker behind; whoever rewrites device tables, reuses the staging buffers from
// another stream or frees buffers waits for it first (handle state was not stream-safe).
static int mark_async(sdfa_handle *h, cudaStream_t s) {
    if (!h->ev_async) CUDA_TRY(cudaEventCreateWithFlags(&h->ev_async, cudaEventDisableTiming));
    CUDA_TRY(cudaEventRecord(h->ev_async, s));
    h->async_pending = true;
    return SDFA_OK;
}
// A stream-ordered call on another stream than the previous one shares its workspaces: order the two on the device.
static int order_after_last(sdfa_handle *h, cudaStream_t s) {
    if (h->async_pending && s != h->last_stream) CUDA_TRY(cudaStreamWaitEvent(s, h->ev_async, 0));
    h->last_stream = s;
    return SDFA_OK;
}
static int wait_async(sdfa_handle *h) {
    if (h->async_pending) {
        CUDA_TRY(cudaEventSynchronize(h->ev_async));
        h->async_pending = false;
    }
    return SDFA_OK;
}

static void default_eq_src(sdfa_handle *h) {
    // no correspondences given at call time: block i reads source triangle i for i < n_tris, the remaining
    // blocks stay zero (deform_triangle_impl.hpp:224, :249-253)
    const HostPlan &p = h->host;
    h->eq_src_host.assign(p.n_eq, -2);
    for (int k = 0; k < std::min(p.n_eq, p.n_tris); ++k) h->eq_src_host[k] = k;
    h->n_src_tris = p.n_tris;
}

static int upload_eq_src(sdfa_handle *h) {
    h->has_pca = false;                       // the compact basis is laid out per equation source: re-pack on change
    if (h->dev.device < 0) return SDFA_OK;
    DeviceGuard guard;
    int rc;
    if ((rc = guard.enter(h->dev.device))) return rc;
    if ((rc = wait_async(h))) return rc;      // kernels of an earlier stream-ordered call may still read these tables
    CUDA_TRY(cudaMemcpy(h->dev.eq_src, h->eq_src_host.data(), h->eq_src_host.size() * 4, cudaMemcpyHostToDevice));
    // the assembly walks carry every equation's current source triangle next to it
    const AssemblyPlan &ap = h->host.asmplan;
    std::vector<int4> walk(ap.warp_sched.size());
    for (size_t b = 0; b < ap.blocks.size(); ++b)
        for (int w = 0; w < ASM_WARPS_PER_BLOCK; ++w)
            for (int i = ap.warp_ptr[b * ASM_WARPS_PER_BLOCK + w]; i < ap.warp_ptr[b * ASM_WARPS_PER_BLOCK + w + 1]; ++i) {
                const int e = ap.warp_sched[i];
                const int g = ap.blocks[b].eq_begin + e;
                walk[i] = e >= 0 ? make_int4(e, h->eq_src_host[ap.eq_id[g]], ap.eq_slot[g], 0) : make_int4(e, -1, 0, 0);
            }
    CUDA_TRY(cudaMemcpy(h->dev.asm_walk, walk.data(), walk.size() * sizeof(int4), cudaMemcpyHostToDevice));
    std::vector<int32_t> local(ap.eq_id.size());
    for (size_t g = 0; g < local.size(); ++g) local[g] = h->eq_src_host[ap.eq_id[g]];
    CUDA_TRY(cudaMemcpy(h->dev.asm_eq_src_local, local.data(), local.size() * 4, cudaMemcpyHostToDevice));
    return SDFA_OK;
}

// compact slot -> (source triangle * 9 + component) of the frame-tiled decode layout, -1 where nothing is decoded
static std::vector<int32_t> compact_map(const sdfa_handle *h) {
    const AssemblyPlan &ap = h->host.asmplan;
    std::vector<int32_t> map((size_t)ap.compact_stride, -1);
    for (size_t g = 0; g < ap.slot_eq.size(); ++g) {
        const int src = h->eq_src_host[ap.slot_eq[g]];
        if (src < 0) continue;
        for (int j = 0; j < 6; ++j) map[g * 6 + j] = src * 9 + j;
        for (int j = 0; j < 3; ++j) map[(size_t)ap.compact_s_rows + g * 3 + j] = src * 9 + 6 + j;
    }
    return map;
}

// The output kernel's per-chunk tables for the vertex list `verts` (one output row per entry): element -> free line
// (or constant), line -> scratch offset + base.  x_base is kept in the Cholesky order (row iperm[f]); lines are indexed
// by scratch row.
static int upload_out_tables(sdfa_handle *h, DevicePlan::OutTables &t, const std::vector<int> &verts) {
    const HostPlan &p = h->host;
    DevicePlan &d = h->dev;
    constexpr int VC = 64, EC = VC * 3;
    const int nv = (int)verts.size(), chunks = (nv + VC - 1) / VC;
    std::vector<int16_t> line_of((size_t)chunks * EC, -1);
    std::vector<float> cval((size_t)chunks * EC, 0.f), hi, lo;
    std::vector<int32_t> ptr(chunks + 1, 0), off;
    int max_lines = 0;
    for (int k = 0; k < chunks; ++k) {
        ptr[k] = (int32_t)off.size();
        for (int e = 0; e < EC; ++e) {
            const int i = k * VC + e / 3, c = e % 3;
            if (i >= nv) break;
            const int v = verts[i];
            const int f = p.vi_to_free[v];
            if (f < 0) { cval[(size_t)k * EC + e] = p.cnst_pos[(size_t)p.vi_to_cnst[v] * 3 + c]; continue; }
            const double x = p.x_base[(size_t)p.iperm[f] * 3 + c];
            line_of[(size_t)k * EC + e] = (int16_t)((int)off.size() - ptr[k]);
            off.push_back(p.scratch_row[f] * d.layout.row_stride + c * d.layout.c_stride);
            hi.push_back((float)x);
            lo.push_back((float)(x - (double)hi.back()));
        }
        max_lines = std::max(max_lines, (int)off.size() - ptr[k]);
    }
    ptr[chunks] = (int32_t)off.size();
    t.n_rows = nv;
    t.max_lines = max_lines;
    CUDA_TRY(cudaMemcpy(t.line_of, line_of.data(), line_of.size() * 2, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(t.cval, cval.data(), cval.size() * 4, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(t.line_ptr, ptr.data(), ptr.size() * 4, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(t.line_off, off.data(), off.size() * 4, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(t.line_hi, hi.data(), hi.size() * 4, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(t.line_lo, lo.data(), lo.size() * 4, cudaMemcpyHostToDevice));
    return SDFA_OK;
}

static int upload_base(sdfa_handle *h) {
    if (h->dev.device < 0) return SDFA_OK;
    const HostPlan &p = h->host;
    DevicePlan &d = h->dev;
    DeviceGuard guard;
    int rc;
    if ((rc = guard.enter(d.device))) return rc;
    if ((rc = wait_async(h))) return rc;      // kernels of an earlier stream-ordered call may still read these tables
    std::vector<int> all(p.n_verts);
    for (int v = 0; v < p.n_verts; ++v) all[v] = v;
    if ((rc = upload_out_tables(h, d.out_full, all))) return rc;
    if ((rc = upload_out_tables(h, d.out_free, p.free_to_vi))) return rc;
    // expansion table: output element -> element of the free rows, or the constrained value
    std::vector<int32_t> src((size_t)p.n_verts * 3, -1);
    std::vector<float> cval((size_t)p.n_verts * 3, 0.f);
    for (int v = 0; v < p.n_verts; ++v)
        for (int c = 0; c < 3; ++c) {
            if (p.vi_to_free[v] >= 0) src[(size_t)v * 3 + c] = p.vi_to_free[v] * 3 + c;
            else cval[(size_t)v * 3 + c] = p.cnst_pos[(size_t)p.vi_to_cnst[v] * 3 + c];
        }
    CUDA_TRY(cudaMemcpy(d.exp_src, src.data(), src.size() * 4, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(d.exp_cval, cval.data(), cval.size() * 4, cudaMemcpyHostToDevice));
    return SDFA_OK;
}

int sdfa_create(sdfa_handle **out, const float *verts, int n_verts, const uint32_t *tris, int n_tris,
                const uint32_t *cnsts, int n_cnsts, const uint32_t *corr_count, double reg, int device) {
    return sdfa_create_with(out, verts, n_verts, tris, n_tris, cnsts, n_cnsts, corr_count, reg, device, nullptr);
}

// "key=value;key=value" -> map; unknown keys are an error so that a typo does not silently select the default
static int parse_options(const char *options, std::map<std::string, std::string> &kv) {
    static const char *known[] = {"solver", "pipe_chunk", "frames_per_tile", "asm_rows", "ts_leaf", "asm_gather", "decode", "output", "output_frames"};
    const std::string text = options ? options : "";
    size_t at = 0;
    while (at < text.size()) {
        size_t end = text.find(';', at);
        if (end == std::string::npos) end = text.size();
        const std::string item = text.substr(at, end - at);
        at = end + 1;
        if (item.empty()) continue;
        const size_t eq = item.find('=');
        if (eq == std::string::npos) return fail(SDFA_ERR_ARG, "sdfa_create_with: option '" + item + "' is not key=value");
        const std::string key = item.substr(0, eq);
        bool ok = false;
        for (const char *k : known) ok |= key == k;
        if (!ok) return fail(SDFA_ERR_ARG, "sdfa_create_with: unknown option '" + key + "'");
        kv[key] = item.substr(eq + 1);
    }
    return SDFA_OK;
}

int sdfa_create_with(sdfa_handle **out, const float *verts, int n_verts, const uint32_t *tris, int n_tris,
                     const uint32_t *cnsts, int n_cnsts, const uint32_t *corr_count, double reg, int device,
                     const char *options) {
    std::map<std::string, std::string> opt;
    {
        int orc;
        if ((orc = parse_options(options, opt))) return orc;
    }
    // an option given in `options` wins over the environment variable of the same purpose (tuning / A-B runs only)
    auto setting = [&](const char *key, const char *env) -> std::string {
        auto it = opt.find(key);
        if (it != opt.end()) return it->second;
        const char *v = std::getenv(env);
        return v ? v : "";
    };
    if (!out || !verts || !tris || n_verts <= 0 || n_tris <= 0 || n_cnsts < 0 || (n_cnsts > 0 && !cnsts))
        return fail(SDFA_ERR_ARG, "sdfa_create: bad arguments");
    *out = nullptr;
    sdfa_handle *h = new sdfa_handle();
    HostPlan &p = h->host;
    p.n_verts = n_verts; p.n_tris = n_tris; p.n_cnsts = n_cnsts; p.reg = reg;
    p.verts.assign(verts, verts + (size_t)n_verts * 3);
    p.tris.assign(tris, tris + (size_t)n_tris * 3);
    if (n_cnsts) p.cnsts.assign(cnsts, cnsts + n_cnsts);
    if (corr_count) p.corr_count.assign(corr_count, corr_count + n_tris);
    std::string err;
    int rc;
    try {
        if ((rc = build_system(p, err)) != 0) { delete h; return fail(SDFA_ERR_ARG, "sdfa_create: " + err); }
        if ((rc = order_and_factor(p, err)) != 0) { delete h; return fail(SDFA_ERR_FACTOR, "sdfa_create: " + err); }
        compute_base_solution(p, nullptr);
        {
            // frames per solve tile: 32 unless the resident rows of a large factor do not fit in the 227 KB of
            // shared memory a CTA can have, then 16 or 8 (config 5: subdivided template)
            auto envi = [](const char *k, int d) { const char *v = std::getenv(k); return v ? std::atoi(v) : d; };
            const std::string fpt = setting("frames_per_tile", "SDFA_FRAMES_PER_TILE");
            const int want = fpt.empty() ? 32 : std::atoi(fpt.c_str());
            for (int f = want; f >= 8; f /= 2) {
                build_solve_program(p, envi("SDFA_PIECE_CAP", 64), envi("SDFA_SUPERNODE_CAP", 32), envi("SDFA_SUBTREE_CAP", 16), f);
                if (solve_smem_bytes(p.prog.n_slots, f) <= 227 * 1024) break;
            }
        }
        {
            // solver: the tensor-core plan when the template fits it, else the SIMT sweeps (SDFA_SOLVER=simt|tensor forces)
            std::string solver = setting("solver", "SDFA_SOLVER");
            if (solver.empty()) solver = "auto";
            if (solver != "auto" && solver != "simt" && solver != "tensor") {
                delete h;
                return fail(SDFA_ERR_ARG, "sdfa_create: solver must be auto, simt or tensor");
            }
            if (solver != "simt") {
                const std::string lm = setting("ts_leaf", "SDFA_TS_LEAF");
                build_tensor_plan(p, lm.empty() ? 64 : std::atoi(lm.c_str()));
            } else p.tplan.why_not = "solver=simt";
            if (solver == "tensor" && !p.tplan.valid) {
                std::string why = p.tplan.why_not;
                delete h;
                return fail(SDFA_ERR_UNSUPPORTED, "sdfa_create: solver=tensor but the template does not fit: " + why);
            }
            p.use_tensor = p.tplan.valid;
            p.scratch_row = p.use_tensor ? p.tplan.row_of_free : p.iperm;
        }
        {
            // row blocks of the assembly kernel: at most ASM_ROWS_MAX vertices each (two CTAs of [rows][3][64 frames]
            // accumulators per SM), split evenly -- measured on FLAME: 11 even blocks of 115 beat 10 of 128 by 5 %
            const std::string e = setting("asm_rows", "SDFA_ASM_ROWS");
            const int rows_max = e.empty() ? ASM_ROWS_MAX : std::atoi(e.c_str()), n_blocks = (p.n_free + rows_max - 1) / rows_max;
            build_assembly_plan(p, /*rows_per_block=*/(p.n_free + n_blocks - 1) / std::max(n_blocks, 1), ASM_MAX_EQ);
        }
    } catch (const std::exception &e) {
        delete h;
        return fail(SDFA_ERR_UNSUPPORTED, std::string("sdfa_create: ") + e.what());
    }
    default_eq_src(h);
    {
        const std::string pc = setting("pipe_chunk", "SDFA_PIPE_CHUNK");
        h->pipe_chunk = pc.empty() ? -1 : std::atoi(pc.c_str());
    }
    {
        const std::string g = setting("asm_gather", "SDFA_ASM_GATHER");
        h->dev.asm_gather_gen = g.empty() ? 0 : std::atoi(g.c_str());   // 0: chosen per call by the row size
        const std::string og = setting("output", "SDFA_OUTPUT");
        h->dev.out_gen = og.empty() ? 2 : std::atoi(og.c_str());
        const std::string of = setting("output_frames", "SDFA_OUTPUT_FRAMES");
        if (!of.empty()) {
            const int fc = std::atoi(of.c_str());
            if (fc != 8 && fc != 16 && fc != 32) { delete h; return fail(SDFA_ERR_ARG, "sdfa_create: output_frames must be 8, 16 or 32"); }
            h->dev.out_fc = fc;
        }
        const std::string dk = setting("decode", "SDFA_DECODE");
        if (!dk.empty() && dk != "f16" && dk != "tf32") { delete h; return fail(SDFA_ERR_ARG, "sdfa_create: decode must be f16 or tf32"); }
        h->dev.decode_fp16 = dk != "tf32";
    }
    h->dev.device = device;
    h->dev.n_verts = n_verts; h->dev.n_tris = n_tris; h->dev.n_cnsts = n_cnsts;
    h->dev.n_free = p.n_free; h->dev.n_eq = p.n_eq;
    if (device >= 0) {
        auto up = [&]() -> int {
            DeviceGuard guard;
            int gr;
            if ((gr = guard.enter(device))) return gr;
            cudaDeviceProp prop;
            CUDA_TRY(cudaGetDeviceProperties(&prop, device));
            if (prop.major < 10)
                return fail(SDFA_ERR_CUDA, "sdfa_create: device is not sm_100-class (kernels are built for sm_100a only)");
            h->dev.sm_count = prop.multiProcessorCount;
            if (!p.use_tensor && solve_smem_bytes(p.prog.n_slots, p.prog.frames_per_tile) > (size_t)prop.sharedMemPerBlockOptin)
                return fail(SDFA_ERR_UNSUPPORTED, "sdfa_create: solve state (" + std::to_string(p.prog.n_slots) +
                                                      " rows) does not fit in shared memory");
            DevicePlan &d = h->dev;
            const AssemblyPlan &ap = p.asmplan;
            std::vector<int4> blocks;
            for (auto &b : ap.blocks) blocks.push_back(make_int4(b.eq_begin, b.eq_end, b.row_begin, b.row_end));
            int r;
            if ((r = upload_mut(h, blocks, &d.asm_blocks))) return r;
            {
                // equation record: U0, U1 and, in the two spare floats, the block rows of the three corners
                std::vector<float4> meta(ap.eq_id.size() * 2);
                for (size_t g = 0; g < ap.eq_id.size(); ++g) {
                    float rec[8];
                    std::memcpy(rec, &ap.eq_u[g * 8], 24);
                    std::memcpy(rec + 6, &ap.eq_rows[g * 4], 8);
                    std::memcpy(&meta[2 * g], rec, 32);
                }
                if ((r = upload(h, meta, &d.asm_eq_meta))) return r;
            }
            if ((r = upload(h, ap.row_perm, &d.asm_row_perm))) return r;
            {
                std::vector<int4> walk(ap.warp_sched.size(), make_int4(0, -1, 0, 0));
                if ((r = upload_mut(h, walk, &d.asm_walk))) return r;          // filled by upload_eq_src
                if ((r = upload(h, ap.warp_ptr, &d.asm_warp_ptr))) return r;
                if ((r = upload(h, ap.row_ptr, &d.asm_row_ptr))) return r;
                if ((r = upload(h, ap.inc, &d.asm_inc))) return r;
                std::vector<int32_t> local(ap.eq_id.size(), -1);
                if ((r = upload_mut(h, local, &d.asm_eq_src_local))) return r;
                d.asm_max_walk = 0;
                for (size_t i = 0; i + 1 < ap.warp_ptr.size(); ++i) d.asm_max_walk = std::max(d.asm_max_walk, ap.warp_ptr[i + 1] - ap.warp_ptr[i]);
            }
            d.n_asm_blocks = (int)ap.blocks.size();
            d.asm_max_eq = ap.max_eq_per_block;
            d.asm_max_rows = ap.max_rows_per_block;
            std::vector<int32_t> tmp(p.n_eq, 0);
            if ((r = upload_mut(h, tmp, &d.eq_src))) return r;
            d.compact_stride = ap.compact_stride;
            d.compact_s_rows = ap.compact_s_rows;
            if ((r = upload(h, p.prog.bytes, &d.prog))) return r;
            if ((r = upload(h, p.prog.stage_off, &d.stage_off))) return r;
            if ((r = upload(h, p.prog.io_desc, &d.io_desc))) return r;
            if ((r = upload(h, p.prog.io_phase, &d.io_phase))) return r;
            d.n_stages = (int)p.prog.stage_off.size() - 1;
            d.n_slots = p.prog.n_slots;
            d.n_phases_fwd = p.prog.n_phases_fwd;
            d.n_phases_bwd = p.prog.n_phases_bwd;
            d.frames_per_tile = p.prog.frames_per_tile;
            d.layout.FL = d.frames_per_tile;
            d.layout.tile_stride = (long long)p.n_free * slot_words(d.frames_per_tile);
            d.layout.row_stride = 3 * d.frames_per_tile;
            d.layout.c_stride = d.frames_per_tile;
            d.use_tensor = p.use_tensor;
            if (p.use_tensor) {
                const TensorPlan &tp = p.tplan;
                if (solve_tc_smem_bytes((int)tp.mma.size(), (int)tp.epi.size()) > (size_t)prop.sharedMemPerBlockOptin)
                    return fail(SDFA_ERR_UNSUPPORTED, "sdfa_create: tensor solve program does not fit in shared memory");
                if ((r = upload(h, tp.mma, &d.ts_mma))) return r;
                if ((r = upload(h, tp.epi, &d.ts_epi))) return r;
                if ((r = upload(h, tp.matrix, &d.ts_matrix))) return r;
                if ((r = upload(h, tp.chunk_off, &d.ts_chunk_off))) return r;
                d.ts_n_mma = (int)tp.mma.size();
                d.ts_n_epi = (int)tp.epi.size();
                d.ts_n_chunks = (int)tp.chunk_off.size() - 1;
                d.ts_n_mma_events = tp.n_mma_events;
                d.ts_n_epi_events = tp.n_epi_events;
                d.ts_n_ring_ops = tp.n_ring_ops;
                d.layout.FL = TS_COLS;
                d.layout.tile_stride = 3LL * p.n_free * TS_COLS;
                d.layout.row_stride = TS_COLS;
                d.layout.c_stride = p.n_free * TS_COLS;
            }
            for (DevicePlan::OutTables *t : {&d.out_full, &d.out_free}) {
                const size_t chunks = ((size_t)(t == &d.out_full ? p.n_verts : p.n_free) + 63) / 64;
                std::vector<int16_t> z16(chunks * 192, -1);
                std::vector<float> zf(chunks * 192, 0.f), zl((size_t)p.n_free * 3, 0.f);
                std::vector<int32_t> zp(chunks + 1, 0), zo((size_t)p.n_free * 3, 0);
                if ((r = upload_mut(h, z16, &t->line_of))) return r;
                if ((r = upload_mut(h, zf, &t->cval))) return r;
                if ((r = upload_mut(h, zp, &t->line_ptr))) return r;
                if ((r = upload_mut(h, zo, &t->line_off))) return r;
                if ((r = upload_mut(h, zl, &t->line_hi))) return r;
                if ((r = upload_mut(h, zl, &t->line_lo))) return r;
            }
            {
                std::vector<int32_t> zs((size_t)p.n_verts * 3, -1);
                std::vector<float> zc((size_t)p.n_verts * 3, 0.f);
                if ((r = upload_mut(h, zs, &d.exp_src))) return r;
                if ((r = upload_mut(h, zc, &d.exp_cval))) return r;
            }
            if ((r = upload_base(h))) return r;
            if ((r = upload_eq_src(h))) return r;
            if (!p.use_tensor) CUDA_TRY(configure_solve(d));
            CUDA_TRY(configure_decode_tc(d));
            CUDA_TRY(configure_decode_tc16(d));
            if (std::getenv("SDFA_SOLVE_PROFILE")) {
                std::vector<long long> zero(std::max((size_t)h->dev.sm_count * 4 * 8, (5 * p.tplan.mma.size() + 6 * p.tplan.epi.size())), 0);
                if ((r = upload_mut(h, zero, &d.solve_prof))) return r;
            }
            return SDFA_OK;
        };
        rc = up();
        if (rc != SDFA_OK) { std::string keep = g_err; sdfa_destroy(h); g_err = keep; return rc; }
    }
    *out = h;
    return SDFA_OK;
}

void sdfa_destroy(sdfa_handle *h) {
    if (!h) return;
    if (h->dev.device >= 0) {
        DeviceGuard guard;
        guard.enter(h->dev.device);
        if (h->async_pending) cudaEventSynchronize(h->ev_async);
        if (h->ev_async) cudaEventDestroy(h->ev_async);
        for (void *p : h->allocs) cudaFree(p);
        for (void *p : h->pca_allocs) cudaFree(p);
        for (auto &w : h->ws) {
            for (float *p : {w.rhs, w.rhs2, w.dgrad_c, w.io_in, w.io_out, w.io_in2, w.ximg_s, w.ximg_r}) if (p) cudaFree(p);
            if (w.stream) cudaStreamDestroy(w.stream);
            if (w.side) cudaStreamDestroy(w.side);
            for (cudaEvent_t e : {w.ev_solved[0], w.ev_solved[1], w.ev_out[0], w.ev_out[1]}) if (e) cudaEventDestroy(e);
        }
        for (auto &e : h->ev) if (e) cudaEventDestroy(e);
    }
    delete h;
}

int sdfa_info(const sdfa_handle *h, int *n_verts, int *n_tris, int *n_cnsts, int *n_free, int *n_eq, int *n_active,
              long long *nnz_l) {
    if (!h) return fail(SDFA_ERR_ARG, "sdfa_info: NULL handle");
    const HostPlan &p = h->host;
    if (n_verts) *n_verts = p.n_verts;
    if (n_tris) *n_tris = p.n_tris;
    if (n_cnsts) *n_cnsts = p.n_cnsts;
    if (n_free) *n_free = p.n_free;
    if (n_eq) *n_eq = p.n_eq;
    if (n_active) *n_active = p.n_active;
    if (nnz_l) *nnz_l = (long long)p.l_rowidx.size();
    return SDFA_OK;
}

int sdfa_set_constraint_positions(sdfa_handle *h, const float *cnst_verts_host) {
    if (!h) return fail(SDFA_ERR_ARG, "NULL handle");
    std::lock_guard<std::recursive_mutex> lock(h->mu);
    HostPlan &p = h->host;
    if (p.n_cnsts == 0) return SDFA_OK;
    // unchanged positions: nothing to redo
    bool same = true;
    for (int i = 0; i < p.n_cnsts * 3 && same; ++i) {
        float want = cnst_verts_host ? cnst_verts_host[i] : p.verts[(size_t)p.cnsts[i / 3] * 3 + i % 3];
        same = (std::memcmp(&want, &p.cnst_pos[i], 4) == 0);
    }
    if (same) return SDFA_OK;
    compute_base_solution(p, cnst_verts_host);
    return upload_base(h);
}

int sdfa_set_correspondences(sdfa_handle *h, const uint32_t *corr_count, const uint32_t *corr_faces, int n_src_tris) {
    if (!h) return fail(SDFA_ERR_ARG, "NULL handle");
    std::lock_guard<std::recursive_mutex> lock(h->mu);
    const HostPlan &p = h->host;
    std::vector<int32_t> prev = h->eq_src_host;
    if (!corr_count) {
        default_eq_src(h);
    } else {
        if (!corr_faces) return fail(SDFA_ERR_ARG, "sdfa_set_correspondences: corr_faces is NULL");
        // deform_triangle_impl.hpp:246-269: the block counter fi advances by max(1,count) per target triangle
        std::vector<int32_t> m;
        for (int i = 0; i < p.n_tris; ++i) {
            if (corr_count[i] > 0) {
                for (uint32_t j = 0; j < corr_count[i]; ++j) {
                    uint32_t s = corr_faces[m.size()];
                    if ((int)s >= n_src_tris) return fail(SDFA_ERR_ARG, "sdfa_set_correspondences: source triangle out of range");
                    m.push_back((int32_t)s);
                }
            } else m.push_back(-1);
        }
        if ((int)m.size() != p.n_eq)
            return fail(SDFA_ERR_ARG, "sdfa_set_correspondences: counts do not match the ones given to sdfa_create");
        h->eq_src_host.swap(m);
        h->n_src_tris = n_src_tris;
    }
    if (prev == h->eq_src_host) return SDFA_OK;
    return upload_eq_src(h);
}

static int time_mark(sdfa_handle *h, int i, cudaStream_t s) {
    if (!h->timing) return SDFA_OK;
    if (!h->ev[i]) CUDA_TRY(cudaEventCreate(&h->ev[i]));
    CUDA_TRY(cudaEventRecord(h->ev[i], s));
    return SDFA_OK;
}
static int time_finish(sdfa_handle *h, cudaStream_t s, bool decoded) {
    if (!h->timing) return SDFA_OK;
    CUDA_TRY(cudaStreamSynchronize(s));
    for (int i = 0; i < 4; ++i) h->last_ms[i] = 0.f;
    if (decoded) CUDA_TRY(cudaEventElapsedTime(&h->last_ms[0], h->ev[0], h->ev[1]));
    CUDA_TRY(cudaEventElapsedTime(&h->last_ms[1], h->ev[1], h->ev[2]));
    CUDA_TRY(cudaEventElapsedTime(&h->last_ms[2], h->ev[2], h->ev[3]));
    CUDA_TRY(cudaEventElapsedTime(&h->last_ms[3], h->ev[3], h->ev[4]));
    return SDFA_OK;
}

static size_t out_row_floats(const sdfa_handle *h) { return (size_t)(h->free_only ? h->dev.n_free : h->dev.n_verts) * 3; }

// One pass over frames [0, n_frames): assembly, solve and -- on stream `so` -- the output kernel.
static int reconstruct_pass(sdfa_handle *h, float *rhs, const float *dgrad_dev, long long stride, bool staged, int mode,
                            int n_frames, float *out_dev, cudaStream_t s, cudaStream_t so, cudaEvent_t solved) {
    int rc;
    if ((rc = time_mark(h, 1, s))) return rc;
    CUDA_TRY(launch_assembly(h->dev, dgrad_dev, stride, staged, n_frames, mode, rhs, s));
    if ((rc = time_mark(h, 2, s))) return rc;
    if (h->dev.use_tensor) CUDA_TRY(launch_solve_tc(h->dev, rhs, n_frames, s));
    else CUDA_TRY(launch_solve(h->dev, rhs, n_frames, s));
    if ((rc = time_mark(h, 3, s))) return rc;
    if (so != s) {
        CUDA_TRY(cudaEventRecord(solved, s));
        CUDA_TRY(cudaStreamWaitEvent(so, solved, 0));
    }
    CUDA_TRY(launch_output(h->dev, h->free_only ? h->dev.out_free : h->dev.out_full, rhs, n_frames, out_dev, so));
    return time_mark(h, 4, s);
}

static int grow_scratch(sdfa_handle *h, float **buf, size_t *cap, int n_frames, cudaStream_t s) {
    const size_t need = scratch_floats(h->dev, n_frames);
    if (*cap >= need) return SDFA_OK;
    int rc;
    if ((rc = grow(buf, cap, need))) return rc;
    // columns of a partially filled tensor tile that no frame owns are solved too: keep them finite
    CUDA_TRY(cudaMemsetAsync(*buf, 0, need * sizeof(float), s));
    return SDFA_OK;
}

// Frames per chunk of the device-resident entry points (0: no chunking): four waves of 128-frame solve tiles.  Larger
// batches run chunk by chunk, which bounds the workspace (compact dgrad 94 KB + scratch 2 x 18 KB per frame of a chunk,
// not of the batch); the output kernel of a chunk runs on a second stream under the next chunk's kernels.  Measured on
// the 75 600-frame bench batch the overlap is worth nothing (one-wave chunks: 7.31 ms against 7.28 ms in one pass --
// the persistent decode and solve kernels leave no room for a second resident CTA), so the chunk is sized for memory.
// Per-kernel timing runs unchunked.
// Frames on grid.y: the gather assembly launches n / 16 rows of CTAs, so one pass takes at most this many frames.
static const int MAX_FRAMES_PER_PASS = 65535 * 16;
static int pipe_chunk(const sdfa_handle *h, int n_frames) {
    const int opt = h->pipe_chunk;                     // sdfa_set_option("pipe_chunk"); default from SDFA_PIPE_CHUNK at create
    if (h->timing || opt == 0) return n_frames > MAX_FRAMES_PER_PASS ? MAX_FRAMES_PER_PASS / 128 * 128 : 0;
    int chunk = opt > 0 ? (opt + 127) / 128 * 128 : h->dev.sm_count * 128 * 4;
    if (opt <= 0) {
        // large templates: keep each of the two scratch buffers under 2 GiB (config 5: 248 KB per frame -> 8448 frames)
        const size_t per_128 = scratch_floats(h->dev, 128) * sizeof(float);
        const size_t fit = ((size_t)2 << 30) / std::max<size_t>(per_128, 1);
        if ((size_t)chunk / 128 > fit) chunk = (int)std::max<size_t>(fit, 1) * 128;
    }
    return n_frames > chunk ? chunk : 0;
}

static int pipe_setup(sdfa_handle::Workspace &w) {
    if (!w.side) CUDA_TRY(cudaStreamCreateWithFlags(&w.side, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
        if (!w.ev_solved[i]) CUDA_TRY(cudaEventCreateWithFlags(&w.ev_solved[i], cudaEventDisableTiming));
        if (!w.ev_out[i]) CUDA_TRY(cudaEventCreateWithFlags(&w.ev_out[i], cudaEventDisableTiming));
    }
    return SDFA_OK;
}

// Runs `decode(f0, nf)` (may be empty) and the reconstruction chunk by chunk; dgrad_of(f0) is the chunk's input.
template <class Decode, class Input>
static int reconstruct_chunks(sdfa_handle *h, sdfa_handle::Workspace &w, int chunk, Decode decode, Input dgrad_of,
                              long long stride, bool staged, int mode, int n_frames, float *out_dev, cudaStream_t s) {
    int rc;
    if ((rc = pipe_setup(w))) return rc;
    if ((rc = grow_scratch(h, &w.rhs, &w.rhs_cap, chunk, s))) return rc;
    if ((rc = grow_scratch(h, &w.rhs2, &w.rhs2_cap, chunk, s))) return rc;
    const size_t row_out = out_row_floats(h);
    int i = 0;
    for (int f0 = 0; f0 < n_frames; f0 += chunk, ++i) {
        const int nf = std::min(chunk, n_frames - f0), b = i & 1;
        if (i >= 2) CUDA_TRY(cudaStreamWaitEvent(s, w.ev_out[b], 0));      // the scratch buffer is free again
        if ((rc = decode(f0, nf))) return rc;
        if ((rc = reconstruct_pass(h, b ? w.rhs2 : w.rhs, dgrad_of(f0), stride, staged, mode, nf,
                                   out_dev + (size_t)f0 * row_out, s, w.side, w.ev_solved[b]))) return rc;
        CUDA_TRY(cudaEventRecord(w.ev_out[b], w.side));
    }
    for (int b = 0; b < std::min(i, 2); ++b) CUDA_TRY(cudaStreamWaitEvent(s, w.ev_out[b], 0));
    return SDFA_OK;
}

static int reconstruct_core(sdfa_handle *h, sdfa_handle::Workspace &w, const float *dgrad_dev, long long stride,
                            bool staged, int mode, int n_frames, float *out_dev, cudaStream_t s, bool decoded) {
    int rc;
    const int chunk = staged ? 0 : pipe_chunk(h, n_frames);
    if (chunk)
        return reconstruct_chunks(h, w, chunk, [](int, int) { return (int)SDFA_OK; },
                                  [&](int f0) { return dgrad_dev + (long long)f0 * stride; }, stride, false, mode,
                                  n_frames, out_dev, s);
    if ((rc = grow_scratch(h, &w.rhs, &w.rhs_cap, n_frames, s))) return rc;
    if ((rc = reconstruct_pass(h, w.rhs, dgrad_dev, stride, staged, mode, n_frames, out_dev, s, s, nullptr))) return rc;
    return time_finish(h, s, decoded);
}

static int sdfa_reconstruct_dev_impl(sdfa_handle *h, const float *dgrad_dev, long long dgrad_stride, int n_frames, float *out_dev,
                         void *stream, bool free_only) {
    int rc;
    ENTER_DEVICE(h);
    h->free_only = free_only;
    if (n_frames < 0 || (n_frames > 0 && (!dgrad_dev || !out_dev))) return fail(SDFA_ERR_ARG, "sdfa_reconstruct_dev: bad arguments");
    if (n_frames == 0) return SDFA_OK;
    long long stride = dgrad_stride ? dgrad_stride : (long long)h->n_src_tris * 9;
    if (stride < (long long)h->n_src_tris * 9) return fail(SDFA_ERR_ARG, "sdfa_reconstruct_dev: dgrad_stride is shorter than a frame");
    NvtxRange nvtx("sdfa_reconstruct_dev");
    cudaStream_t s = (cudaStream_t)stream;
    if ((rc = order_after_last(h, s))) return rc;
    if ((rc = reconstruct_core(h, h->ws[0], dgrad_dev, stride, false, ASM_DGRAD, n_frames, out_dev, s, false))) return rc;
    return mark_async(h, s);
}

int sdfa_reconstruct_dev(sdfa_handle *h, const float *dgrad_dev, long long dgrad_stride, int n_frames, float *out_dev, void *stream) {
    return sdfa_reconstruct_dev_impl(h, dgrad_dev, dgrad_stride, n_frames, out_dev, stream, false);
}
int sdfa_reconstruct_free_dev(sdfa_handle *h, const float *dgrad_dev, long long dgrad_stride, int n_frames, float *out_dev, void *stream) {
    return sdfa_reconstruct_dev_impl(h, dgrad_dev, dgrad_stride, n_frames, out_dev, stream, true);
}

// Frames per chunk of the host-buffer entry points: copies of chunk i overlap the kernels of chunk i+1.
static const int HOST_CHUNK = 4096;

static int ws_stream(sdfa_handle::Workspace &w) {
    if (!w.stream) CUDA_TRY(cudaStreamCreateWithFlags(&w.stream, cudaStreamNonBlocking));
    return SDFA_OK;
}

static int sdfa_reconstruct_host_impl(sdfa_handle *h, const float *dgrad_host, int n_frames, float *out_host, bool free_only) {
    int rc;
    ENTER_DEVICE(h);
    h->free_only = free_only;
    if (n_frames < 0 || (n_frames > 0 && (!dgrad_host || !out_host))) return fail(SDFA_ERR_ARG, "sdfa_reconstruct_host: bad arguments");
    const size_t row_in = (size_t)h->n_src_tris * 9, row_out = out_row_floats(h);
    NvtxRange nvtx("sdfa_reconstruct_host");
    if ((rc = wait_async(h))) return rc;
    for (int f0 = 0, k = 0; f0 < n_frames; f0 += HOST_CHUNK, k ^= 1) {
        const int nf = std::min(HOST_CHUNK, n_frames - f0);
        sdfa_handle::Workspace &w = h->ws[k];
        if ((rc = ws_stream(w))) return rc;
        CUDA_TRY(cudaStreamSynchronize(w.stream));            // this set's previous chunk has left the device
        if ((rc = grow(&w.io_in, &w.io_in_cap, (size_t)HOST_CHUNK * row_in))) return rc;
        if ((rc = grow(&w.io_out, &w.io_out_cap, (size_t)HOST_CHUNK * row_out))) return rc;
        CUDA_TRY(cudaMemcpyAsync(w.io_in, dgrad_host + (size_t)f0 * row_in, (size_t)nf * row_in * 4, cudaMemcpyHostToDevice, w.stream));
        if ((rc = reconstruct_core(h, w, w.io_in, (long long)row_in, false, ASM_DGRAD, nf, w.io_out, w.stream, false))) return rc;
        CUDA_TRY(cudaMemcpyAsync(out_host + (size_t)f0 * row_out, w.io_out, (size_t)nf * row_out * 4, cudaMemcpyDeviceToHost, w.stream));
    }
    for (auto &w : h->ws) if (w.stream) CUDA_TRY(cudaStreamSynchronize(w.stream));
    return SDFA_OK;
}

int sdfa_reconstruct_host(sdfa_handle *h, const float *dgrad_host, int n_frames, float *out_host) {
    return sdfa_reconstruct_host_impl(h, dgrad_host, n_frames, out_host, false);
}
int sdfa_reconstruct_free_host(sdfa_handle *h, const float *dgrad_host, int n_frames, float *out_host) {
    return sdfa_reconstruct_host_impl(h, dgrad_host, n_frames, out_host, true);
}

static int single_frame(sdfa_handle *h, const double *in, long long len, int mode, const float *cnst, float *out_host) {
    int rc;
    const HostPlan &p = h->host;
    if (p.n_cnsts > 0 && !cnst)
        return fail(SDFA_ERR_ARG, "cnst_verts is not given, but " + std::to_string(p.n_cnsts) +
                                      " constraints (reference asserts: deform_triangle_impl.hpp:274)");
    if ((rc = sdfa_set_constraint_positions(h, p.n_cnsts > 0 ? cnst : nullptr))) return rc;
    std::vector<float> f32((size_t)len);
    for (long long i = 0; i < len; ++i) f32[(size_t)i] = (float)in[i];
    sdfa_handle::Workspace &w = h->ws[0];
    h->free_only = false;
    if ((rc = wait_async(h))) return rc;
    if (w.stream) CUDA_TRY(cudaStreamSynchronize(w.stream));
    if ((rc = grow(&w.io_in, &w.io_in_cap, (size_t)len))) return rc;
    if ((rc = grow(&w.io_out, &w.io_out_cap, (size_t)p.n_verts * 3))) return rc;
    CUDA_TRY(cudaMemcpyAsync(w.io_in, f32.data(), (size_t)len * 4, cudaMemcpyHostToDevice, 0));
    if ((rc = reconstruct_core(h, w, w.io_in, len, false, mode, 1, w.io_out, 0, false))) return rc;
    CUDA_TRY(cudaMemcpyAsync(out_host, w.io_out, (size_t)p.n_verts * 12, cudaMemcpyDeviceToHost, 0));
    CUDA_TRY(cudaStreamSynchronize(0));
    return SDFA_OK;
}

int sdfa_get_mesh_f64(sdfa_handle *h, const double *dgrad_host, long long dgrad_len, const float *cnst_verts_host,
                      const uint32_t *corr_count, const uint32_t *corr_faces, long long corr_faces_len, float *out_host) {
    int rc;
    ENTER_DEVICE(h);
    if (!dgrad_host || !out_host || dgrad_len <= 0 || dgrad_len % 9) return fail(SDFA_ERR_ARG, "sdfa_get_mesh_f64: bad arguments");
    const HostPlan &p = h->host;
    const int n_src = (int)(dgrad_len / 9);
    if (corr_count) {
        if (corr_faces_len < p.n_eq) return fail(SDFA_ERR_ARG, "sdfa_get_mesh_f64: corr_faces shorter than the number of equation blocks");
        if ((rc = sdfa_set_correspondences(h, corr_count, corr_faces, n_src))) return rc;
    } else {
        if (n_src < std::min(p.n_tris, p.n_eq)) return fail(SDFA_ERR_ARG, "sdfa_get_mesh_f64: deform_grad shorter than 9*n_tris");
        if ((rc = sdfa_set_correspondences(h, nullptr, nullptr, n_src))) return rc;
        h->n_src_tris = n_src;
    }
    return single_frame(h, dgrad_host, dgrad_len, ASM_DGRAD, cnst_verts_host, out_host);
}

int sdfa_get_mesh_from_dm_f64(sdfa_handle *h, const double *dmat_host, long long dmat_len, const float *cnst_verts_host,
                              float *out_host) {
    int rc;
    ENTER_DEVICE(h);
    const HostPlan &p = h->host;
    if (!dmat_host || !out_host || dmat_len < (long long)p.n_tris * 9) return fail(SDFA_ERR_ARG, "sdfa_get_mesh_from_dm_f64: bad arguments");
    if ((rc = sdfa_set_correspondences(h, nullptr, nullptr, (int)(dmat_len / 9)))) return rc;
    return single_frame(h, dmat_host, dmat_len, ASM_MATRIX, cnst_verts_host, out_host);
}

// ---------------------------------------------------------------------------------------------
int sdfa_set_pca(sdfa_handle *h, const float *compT_scale, const float *means_scale, int k_scale,
                 const float *compT_rotat, const float *means_rotat, int k_rotat) {
    int rc;
    ENTER_DEVICE(h);
    if (!compT_scale || !means_scale || !compT_rotat || !means_rotat || k_scale <= 0 || k_rotat <= 0)
        return fail(SDFA_ERR_ARG, "sdfa_set_pca: bad arguments");
    DevicePlan &d = h->dev;
    const int nt = h->n_src_tris;
    // everything is validated and built on the host before the current basis is touched
    const AssemblyPlan &ap = h->host.asmplan;
    std::vector<int32_t> src_s, src_r;
    for (size_t g = 0; g < ap.slot_eq.size(); ++g) {
        const int src = h->eq_src_host[ap.slot_eq[g]];
        if (src >= nt) return fail(SDFA_ERR_ARG, "sdfa_set_pca: basis has fewer triangles than the correspondences refer to");
        for (int j = 0; j < 6; ++j) src_s.push_back(src < 0 ? -1 : src * 6 + j);
        for (int j = 0; j < 3; ++j) src_r.push_back(src < 0 ? -1 : src * 3 + j);
    }
    // tensor-core path: one GEMM row per slot of the frame-tiled compact layout (scale part, rotation part), split
    // into TF32 hi/lo and stored as tile images; the means ride along as an extra K column
    std::vector<float> img_s, img_r;
    const int mt_s = tc_build_basis(compT_scale, means_scale, k_scale, src_s, img_s);
    const int mt_r = tc_build_basis(compT_rotat, means_rotat, k_rotat, src_r, img_r);
    if (mt_s * tc_rows_per_tile() != d.compact_s_rows || mt_r * tc_rows_per_tile() != d.compact_stride - d.compact_s_rows)
        return fail(SDFA_ERR_STATE, "sdfa_set_pca: compact layout and decode tiles disagree");
    // second-generation decode: scaled FP16 hi/lo images of the same rows (decode_tc16.cu), when the widths fit it
    const bool use16 = d.decode_fp16 && tc16_fits(k_scale, k_rotat);
    std::vector<uint16_t> img16_s, img16_r;
    float inv_sw[2] = {1.f, 1.f};
    if (use16) {
        tc16_build_basis(compT_scale, means_scale, k_scale, src_s, img16_s, &inv_sw[0]);
        tc16_build_basis(compT_rotat, means_rotat, k_rotat, src_r, img16_r, &inv_sw[1]);
    }
    // replace the previous basis (every call used to leak ~60 MB): nothing may still be reading it
    if ((rc = wait_async(h))) return rc;
    for (auto &w : h->ws) if (w.stream) CUDA_TRY(cudaStreamSynchronize(w.stream));
    h->has_pca = h->has_full_pca = false;
    for (void *p : h->pca_allocs) cudaFree(p);
    h->pca_allocs.clear();
    d.wfull_scale = d.mfull_scale = d.wfull_rotat = d.mfull_rotat = d.tc_w_scale = d.tc_w_rotat = nullptr;
    d.tc16_w_scale = d.tc16_w_rotat = nullptr;
    d.tc16_ready = false;
    auto put = [&](const float *src, size_t n, float **dst) -> int {
        void *p = nullptr;
        CUDA_TRY(cudaMalloc(&p, std::max<size_t>(n * sizeof(float), 16)));
        h->pca_allocs.push_back(p);
        CUDA_TRY(cudaMemcpy(p, src, n * sizeof(float), cudaMemcpyHostToDevice));
        *dst = static_cast<float *>(p);
        return SDFA_OK;
    };
    if ((rc = put(compT_scale, (size_t)nt * 6 * k_scale, &d.wfull_scale)) || (rc = put(means_scale, (size_t)nt * 6, &d.mfull_scale)) ||
        (rc = put(compT_rotat, (size_t)nt * 3 * k_rotat, &d.wfull_rotat)) || (rc = put(means_rotat, (size_t)nt * 3, &d.mfull_rotat)) ||
        (rc = put(img_s.data(), img_s.size(), &d.tc_w_scale)) || (rc = put(img_r.data(), img_r.size(), &d.tc_w_rotat)))
        return rc;
    if (use16) {
        float *p16s = nullptr, *p16r = nullptr;       // uploaded as raw bytes (two halves per float)
        if ((rc = put(reinterpret_cast<const float *>(img16_s.data()), img16_s.size() / 2, &p16s)) ||
            (rc = put(reinterpret_cast<const float *>(img16_r.data()), img16_r.size() / 2, &p16r)))
            return rc;
        d.tc16_w_scale = reinterpret_cast<const uint16_t *>(p16s);
        d.tc16_w_rotat = reinterpret_cast<const uint16_t *>(p16r);
        d.tc16_inv_sw[0] = inv_sw[0];
        d.tc16_inv_sw[1] = inv_sw[1];
        d.tc16_ready = true;
    }
    d.k_scale = k_scale; d.k_rotat = k_rotat;
    d.tc_mt_scale = mt_s; d.tc_mt_rotat = mt_r;
    d.n_pca_tris = nt;
    h->has_pca = h->has_full_pca = true;
    return SDFA_OK;
}

static int decode_reconstruct_core(sdfa_handle *h, sdfa_handle::Workspace &w, const float *cs_dev, const float *cr_dev,
                                   int n_frames, float *out_dev, cudaStream_t s) {
    int rc;
    const long long stride = h->dev.compact_stride;
    const int chunk = pipe_chunk(h, n_frames), cap = chunk ? chunk : n_frames;
    if ((rc = grow(&w.dgrad_c, &w.dgrad_c_cap, ((size_t)cap + COMPACT_TILE - 1) / COMPACT_TILE * COMPACT_TILE * stride))) return rc;
    if ((rc = grow(&w.ximg_s, &w.ximg_s_cap, std::max(tc_ximg_floats(cap, h->dev.k_scale), tc16_ximg_floats(cap, h->dev.k_scale))))) return rc;
    if ((rc = grow(&w.ximg_r, &w.ximg_r_cap, std::max(tc_ximg_floats(cap, h->dev.k_rotat), tc16_ximg_floats(cap, h->dev.k_rotat))))) return rc;
    auto decode = [&](int f0, int nf) -> int {
        CUDA_TRY((h->dev.tc16_ready ? launch_decode_tc16 : launch_decode_tc)(
            h->dev, cs_dev + (size_t)f0 * h->dev.k_scale, cr_dev + (size_t)f0 * h->dev.k_rotat, nf, w.ximg_s, w.ximg_r, w.dgrad_c, s));
        return SDFA_OK;
    };
    if (chunk)
        return reconstruct_chunks(h, w, chunk, decode, [&](int) { return (const float *)w.dgrad_c; }, stride, true,
                                  ASM_DGRAD, n_frames, out_dev, s);
    if ((rc = time_mark(h, 0, s))) return rc;
    if ((rc = decode(0, n_frames))) return rc;
    return reconstruct_core(h, w, w.dgrad_c, stride, true, ASM_DGRAD, n_frames, out_dev, s, true);
}

static int sdfa_decode_reconstruct_dev_impl(sdfa_handle *h, const float *coeff_scale_dev, const float *coeff_rotat_dev, int n_frames,
                                float *out_dev, void *stream, bool free_only) {
    int rc;
    ENTER_DEVICE(h);
    h->free_only = free_only;
    if (!h->has_pca) return fail(SDFA_ERR_STATE, "sdfa_decode_reconstruct: call sdfa_set_pca first (and again after changing correspondences)");
    if (n_frames < 0 || (n_frames > 0 && (!coeff_scale_dev || !coeff_rotat_dev || !out_dev))) return fail(SDFA_ERR_ARG, "bad arguments");
    if (n_frames == 0) return SDFA_OK;
    NvtxRange nvtx("sdfa_decode_reconstruct_dev");
    cudaStream_t s = (cudaStream_t)stream;
    if ((rc = order_after_last(h, s))) return rc;
    if ((rc = decode_reconstruct_core(h, h->ws[0], coeff_scale_dev, coeff_rotat_dev, n_frames, out_dev, s))) return rc;
    return mark_async(h, s);
}

int sdfa_decode_reconstruct_dev(sdfa_handle *h, const float *coeff_scale_dev, const float *coeff_rotat_dev, int n_frames, float *out_dev, void *stream) {
    return sdfa_decode_reconstruct_dev_impl(h, coeff_scale_dev, coeff_rotat_dev, n_frames, out_dev, stream, false);
}
int sdfa_decode_reconstruct_free_dev(sdfa_handle *h, const float *coeff_scale_dev, const float *coeff_rotat_dev, int n_frames, float *out_dev, void *stream) {
    return sdfa_decode_reconstruct_dev_impl(h, coeff_scale_dev, coeff_rotat_dev, n_frames, out_dev, stream, true);
}

static int sdfa_decode_reconstruct_host_impl(sdfa_handle *h, const float *coeff_scale_host, const float *coeff_rotat_host,
                                 int n_frames, float *out_host, bool free_only) {
    int rc;
    ENTER_DEVICE(h);
    h->free_only = free_only;
    if (!h->has_pca) return fail(SDFA_ERR_STATE, "sdfa_decode_reconstruct: call sdfa_set_pca first");
    if (n_frames < 0 || (n_frames > 0 && (!coeff_scale_host || !coeff_rotat_host || !out_host))) return fail(SDFA_ERR_ARG, "bad arguments");
    const size_t ks = (size_t)h->dev.k_scale, kr = (size_t)h->dev.k_rotat, row_out = out_row_floats(h);
    NvtxRange nvtx("sdfa_decode_reconstruct_host");
    if ((rc = wait_async(h))) return rc;
    for (int f0 = 0, k = 0; f0 < n_frames; f0 += HOST_CHUNK, k ^= 1) {
        const int nf = std::min(HOST_CHUNK, n_frames - f0);
        sdfa_handle::Workspace &w = h->ws[k];
        if ((rc = ws_stream(w))) return rc;
        CUDA_TRY(cudaStreamSynchronize(w.stream));
        if ((rc = grow(&w.io_in, &w.io_in_cap, (size_t)HOST_CHUNK * ks))) return rc;
        if ((rc = grow(&w.io_in2, &w.io_in2_cap, (size_t)HOST_CHUNK * kr))) return rc;
        if ((rc = grow(&w.io_out, &w.io_out_cap, (size_t)HOST_CHUNK * row_out))) return rc;
        CUDA_TRY(cudaMemcpyAsync(w.io_in, coeff_scale_host + (size_t)f0 * ks, (size_t)nf * ks * 4, cudaMemcpyHostToDevice, w.stream));
        CUDA_TRY(cudaMemcpyAsync(w.io_in2, coeff_rotat_host + (size_t)f0 * kr, (size_t)nf * kr * 4, cudaMemcpyHostToDevice, w.stream));
        if ((rc = decode_reconstruct_core(h, w, w.io_in, w.io_in2, nf, w.io_out, w.stream))) return rc;
        CUDA_TRY(cudaMemcpyAsync(out_host + (size_t)f0 * row_out, w.io_out, (size_t)nf * row_out * 4, cudaMemcpyDeviceToHost, w.stream));
    }
    for (auto &w : h->ws) if (w.stream) CUDA_TRY(cudaStreamSynchronize(w.stream));
    return SDFA_OK;
}

int sdfa_decode_reconstruct_host(sdfa_handle *h, const float *coeff_scale_host, const float *coeff_rotat_host, int n_frames, float *out_host) {
    return sdfa_decode_reconstruct_host_impl(h, coeff_scale_host, coeff_rotat_host, n_frames, out_host, false);
}
int sdfa_decode_reconstruct_free_host(sdfa_handle *h, const float *coeff_scale_host, const float *coeff_rotat_host, int n_frames, float *out_host) {
    return sdfa_decode_reconstruct_host_impl(h, coeff_scale_host, coeff_rotat_host, n_frames, out_host, true);
}

int sdfa_decode_dgrad_dev(sdfa_handle *h, const float *coeff_scale_dev, const float *coeff_rotat_dev, int n_frames,
                          float *dgrad_dev, void *stream) {
    int rc;
    ENTER_DEVICE(h);
    if (!h->has_full_pca) return fail(SDFA_ERR_STATE, "sdfa_decode_dgrad_dev: call sdfa_set_pca first");
    if (n_frames < 0 || (n_frames > 0 && (!coeff_scale_dev || !coeff_rotat_dev || !dgrad_dev))) return fail(SDFA_ERR_ARG, "bad arguments");
    if (h->n_src_tris != h->dev.n_pca_tris)
        return fail(SDFA_ERR_STATE, "sdfa_decode_dgrad_dev: the source triangle count changed since sdfa_set_pca; call it again");
    if (n_frames == 0) return SDFA_OK;
    cudaStream_t s = (cudaStream_t)stream;
    if ((rc = order_after_last(h, s))) return rc;
    CUDA_TRY(launch_decode_full(h->dev, coeff_scale_dev, coeff_rotat_dev, n_frames, dgrad_dev, s));
    return mark_async(h, s);
}

int sdfa_decode_compact_dev(sdfa_handle *h, const float *coeff_scale_dev, const float *coeff_rotat_dev, int n_frames,
                            float *dgrad_compact_dev, void *stream) {
    int rc;
    ENTER_DEVICE(h);
    if (!h->has_pca) return fail(SDFA_ERR_STATE, "sdfa_decode_compact_dev: call sdfa_set_pca first");
    if (n_frames < 0 || (n_frames > 0 && (!coeff_scale_dev || !coeff_rotat_dev || !dgrad_compact_dev))) return fail(SDFA_ERR_ARG, "bad arguments");
    if (n_frames == 0) return SDFA_OK;
    cudaStream_t s = (cudaStream_t)stream;
    if ((rc = order_after_last(h, s))) return rc;
    sdfa_handle::Workspace &w = h->ws[0];
    if ((rc = grow(&w.ximg_s, &w.ximg_s_cap, std::max(tc_ximg_floats(n_frames, h->dev.k_scale), tc16_ximg_floats(n_frames, h->dev.k_scale))))) return rc;
    if ((rc = grow(&w.ximg_r, &w.ximg_r_cap, std::max(tc_ximg_floats(n_frames, h->dev.k_rotat), tc16_ximg_floats(n_frames, h->dev.k_rotat))))) return rc;
    CUDA_TRY((h->dev.tc16_ready ? launch_decode_tc16 : launch_decode_tc)(h->dev, coeff_scale_dev, coeff_rotat_dev, n_frames, w.ximg_s,
                                                                         w.ximg_r, dgrad_compact_dev, s));
    return mark_async(h, s);
}

int sdfa_compact_layout(const sdfa_handle *h, int32_t *map, int cap) {
    if (!h) return -1;
    const std::vector<int32_t> m = compact_map(h);
    if (map) std::memcpy(map, m.data(), sizeof(int32_t) * (size_t)std::min((int)m.size(), cap));
    return (int)m.size();
}

int sdfa_get_deform_grad_host(const float *verts_a, const float *verts_b, int n_verts, const uint32_t *tris, int n_tris,
                              double eps, int as_matrix, int device, double *out_host) {
    if (!verts_a || !verts_b || !tris || !out_host || n_verts <= 0 || n_tris <= 0)
        return fail(SDFA_ERR_ARG, "sdfa_get_deform_grad_host: bad arguments");
    for (int i = 0; i < n_tris * 3; ++i)
        if (tris[i] >= (uint32_t)n_verts) return fail(SDFA_ERR_ARG, "sdfa_get_deform_grad_host: triangle index out of range");
    if (device < 0) return fail(SDFA_ERR_CUDA, "sdfa_get_deform_grad_host: needs a CUDA device; this library has no CPU path");
    DeviceGuard guard;
    {
        int gr;
        if ((gr = guard.enter(device))) return gr;
    }
    float *da = nullptr, *db = nullptr;
    uint32_t *dt = nullptr;
    double *dout = nullptr;
    auto cleanup = [&]() { cudaFree(da); cudaFree(db); cudaFree(dt); cudaFree(dout); };
    cudaError_t e;
    if ((e = cudaMalloc((void **)&da, (size_t)n_verts * 12)) || (e = cudaMalloc((void **)&db, (size_t)n_verts * 12)) ||
        (e = cudaMalloc((void **)&dt, (size_t)n_tris * 12)) || (e = cudaMalloc((void **)&dout, (size_t)n_tris * 72)) ||
        (e = cudaMemcpy(da, verts_a, (size_t)n_verts * 12, cudaMemcpyHostToDevice)) ||
        (e = cudaMemcpy(db, verts_b, (size_t)n_verts * 12, cudaMemcpyHostToDevice)) ||
        (e = cudaMemcpy(dt, tris, (size_t)n_tris * 12, cudaMemcpyHostToDevice)) ||
        (e = launch_deform_grad(da, db, 0, dt, n_tris, 1, eps, as_matrix, dout, true, 0)) ||
        (e = cudaMemcpy(out_host, dout, (size_t)n_tris * 72, cudaMemcpyDeviceToHost))) {
        cleanup();
        return fail(SDFA_ERR_CUDA, std::string("sdfa_get_deform_grad_host: ") + cudaGetErrorString(e));
    }
    cleanup();
    return SDFA_OK;
}

int sdfa_deform_grad_batch_dev(const float *verts_a_dev, const float *verts_b_dev, int n_verts, const uint32_t *tris_dev,
                               int n_tris, int n_frames, double eps, int as_matrix, float *out_dev, void *stream) {
    if (!verts_a_dev || !verts_b_dev || !tris_dev || !out_dev || n_verts <= 0 || n_tris <= 0 || n_frames < 0)
        return fail(SDFA_ERR_ARG, "sdfa_deform_grad_batch_dev: bad arguments");
    CUDA_TRY(launch_deform_grad(verts_a_dev, verts_b_dev, (long long)n_verts * 3, tris_dev, n_tris, n_frames, eps, as_matrix,
                                out_dev, false, (cudaStream_t)stream));
    return SDFA_OK;
}

int sdfa_seek_dev(const float *seq_dev, int n_src, long long width, const double *timestamps_host,
                  const double *query_host, int n_query, float *out_dev, void *stream) {
    if (!seq_dev || !timestamps_host || n_src <= 0 || width <= 0 || n_query < 0 || (n_query > 0 && (!query_host || !out_dev)))
        return fail(SDFA_ERR_ARG, "sdfa_seek_dev: bad arguments");
    if (n_query == 0) return SDFA_OK;
    // saber.stream.seek's search and its two special cases (stream.py:23-46), per query
    std::vector<int2> pairs(n_query);
    std::vector<double> w(n_query, 1.0);
    const double *t = timestamps_host;
    for (int q = 0; q < n_query; ++q) {
        const double ts = query_host[q];
        int left = 0, right = n_src, m = (left + right) / 2;
        while (left < right) {
            m = (left + right) / 2;
            const double tm = t[m], tn = m + 1 < n_src ? t[m + 1] : ts + 1;
            if (tm <= ts && ts < tn) break;
            else if (tm > ts) right = m;
            else left = m + 1;
        }
        if (m >= n_src) m = n_src - 1;
        if (ts < t[m] || ts > t[n_src - 1] || m + 1 >= n_src) pairs[q] = make_int2(m, m);
        else {
            pairs[q] = make_int2(m, m + 1);
            w[q] = (t[m + 1] - ts) / (t[m + 1] - t[m]);
        }
    }
    cudaStream_t s = (cudaStream_t)stream;
    int2 *dp = nullptr;
    double *dw = nullptr;
    CUDA_TRY(cudaMallocAsync((void **)&dp, sizeof(int2) * n_query, s));
    CUDA_TRY(cudaMallocAsync((void **)&dw, sizeof(double) * n_query, s));
    CUDA_TRY(cudaMemcpyAsync(dp, pairs.data(), sizeof(int2) * n_query, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(dw, w.data(), sizeof(double) * n_query, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaStreamSynchronize(s));                   // the pageable host tables above must outlive the copies
    CUDA_TRY(launch_seek(seq_dev, width, dp, dw, n_query, out_dev, s));
    CUDA_TRY(cudaFreeAsync(dp, s));
    CUDA_TRY(cudaFreeAsync(dw, s));
    return SDFA_OK;
}

int sdfa_free_vertices(const sdfa_handle *h, int32_t *ids, int cap) {
    if (!h) return -1;
    const std::vector<int> &f = h->host.free_to_vi;
    if (ids) for (int i = 0; i < std::min((int)f.size(), cap); ++i) ids[i] = f[i];
    return (int)f.size();
}

int sdfa_expand_free_dev(sdfa_handle *h, const float *free_dev, int n_frames, float *out_dev, void *stream) {
    int rc;
    ENTER_DEVICE(h);
    if (n_frames < 0 || (n_frames > 0 && (!free_dev || !out_dev))) return fail(SDFA_ERR_ARG, "sdfa_expand_free_dev: bad arguments");
    if (n_frames == 0) return SDFA_OK;
    cudaStream_t s = (cudaStream_t)stream;
    if ((rc = order_after_last(h, s))) return rc;
    for (int f0 = 0; f0 < n_frames; f0 += MAX_FRAMES_PER_PASS)
        CUDA_TRY(launch_expand(h->dev, free_dev + (size_t)f0 * h->dev.n_free * 3, std::min(MAX_FRAMES_PER_PASS, n_frames - f0),
                               out_dev + (size_t)f0 * h->dev.n_verts * 3, s));
    return mark_async(h, s);
}

int sdfa_set_option(sdfa_handle *h, const char *name, long long value) {
    if (!h || !name) return fail(SDFA_ERR_ARG, "sdfa_set_option: bad arguments");
    std::lock_guard<std::recursive_mutex> lock(h->mu);
    const std::string n(name);
    if (n == "pipe_chunk") {
        if (value < -1 || value > (1 << 30)) return fail(SDFA_ERR_ARG, "sdfa_set_option: pipe_chunk out of range");
        h->pipe_chunk = (int)value;
        return SDFA_OK;
    }
    return fail(SDFA_ERR_ARG, "sdfa_set_option: unknown option '" + n + "'");
}

long long sdfa_launch_count(void) { return launch_counter(); }

int sdfa_set_timing(sdfa_handle *h, int enable) {
    if (!h) return fail(SDFA_ERR_ARG, "NULL handle");
    std::lock_guard<std::recursive_mutex> lock(h->mu);
    h->timing = enable != 0;
    return SDFA_OK;
}
int sdfa_last_timing(const sdfa_handle *h, float ms[4]) {
    if (!h || !ms) return fail(SDFA_ERR_ARG, "bad arguments");
    for (int i = 0; i < 4; ++i) ms[i] = h->last_ms[i];
    return SDFA_OK;
}

// ---------------------------------------------------------------------------------------------
template <typename T>
static long long give(const std::vector<T> &v, void *dst, long long cap) {
    long long bytes = (long long)(v.size() * sizeof(T));
    if (dst && cap > 0) std::memcpy(dst, v.data(), (size_t)std::min(bytes, cap));
    return bytes;
}

long long sdfa_debug_get(const sdfa_handle *h, const char *what, void *dst, long long cap) {
    if (!h || !what) return -1;
    const HostPlan &p = h->host;
    const std::string w(what);
    if (w == "perm") return give(p.perm, dst, cap);
    if (w == "parent") return give(p.parent, dst, cap);
    if (w == "free_to_vi") return give(p.free_to_vi, dst, cap);
    if (w == "l_colptr") return give(p.l_colptr, dst, cap);
    if (w == "l_rowidx") return give(p.l_rowidx, dst, cap);
    if (w == "l_val") return give(p.l_val, dst, cap);
    if (w == "m_colptr") return give(p.m_colptr, dst, cap);
    if (w == "m_rowidx") return give(p.m_rowidx, dst, cap);
    if (w == "m_val") return give(p.m_val, dst, cap);
    if (w == "x_base") return give(p.x_base, dst, cap);
    if (w == "active_eq") return give(p.active_eq, dst, cap);
    if (w == "tri_u") return give(p.tri_u, dst, cap);
    if (w == "prog") return give(p.prog.bytes, dst, cap);
    if (w == "stage_off") return give(p.prog.stage_off, dst, cap);
    if (w == "io_desc") return give(p.prog.io_desc, dst, cap);
    if (w == "io_phase") return give(p.prog.io_phase, dst, cap);
    if (w == "eq_src") return give(h->eq_src_host, dst, cap);
    if (w == "asm_eq_id") return give(p.asmplan.eq_id, dst, cap);
    if (w == "asm_eq_u") return give(p.asmplan.eq_u, dst, cap);
    if (w == "asm_row_perm") return give(p.asmplan.row_perm, dst, cap);
    if (w == "asm_eq_rows") return give(p.asmplan.eq_rows, dst, cap);
    if (w == "asm_colour_ptr") return give(p.asmplan.colour_ptr, dst, cap);
    if (w == "scratch_row") return give(p.scratch_row, dst, cap);
    if (w == "decode_kind") { std::vector<int32_t> v = {h->dev.tc16_ready ? 16 : 32}; return give(v, dst, cap); }
    if (w == "compact_tile") { std::vector<int32_t> v = {COMPACT_TILE}; return give(v, dst, cap); }
    if (w == "ts_mma") return give(p.tplan.mma, dst, cap);
    if (w == "ts_epi") return give(p.tplan.epi, dst, cap);
    if (w == "ts_matrix") return give(p.tplan.matrix, dst, cap);
    if (w == "ts_chunk_off") return give(p.tplan.chunk_off, dst, cap);
    if (w == "ts_why_not") { std::vector<char> v(p.tplan.why_not.begin(), p.tplan.why_not.end()); return give(v, dst, cap); }
    if (w == "ts_stats") {
        std::vector<long long> v = {p.use_tensor ? 1 : 0, p.tplan.valid ? 1 : 0, (long long)p.tplan.mma.size(), (long long)p.tplan.epi.size(),
                                    (long long)p.tplan.chunk_off.size() - 1, (long long)p.tplan.matrix.size(), p.tplan.n_mma_events,
                                    p.tplan.n_epi_events, p.tplan.n_nodes, p.tplan.n_leaves, p.tplan.nk_products, p.tplan.tmem_fwd,
                                    p.tplan.tmem_bwd, (long long)solve_tc_smem_bytes((int)p.tplan.mma.size(), (int)p.tplan.epi.size()),
                                    p.tplan.n_streams, p.tplan.n_ring_ops};
        return give(v, dst, cap);
    }
    if (w == "asm_blocks") {
        std::vector<int> b;
        for (auto &x : p.asmplan.blocks) { b.push_back(x.eq_begin); b.push_back(x.eq_end); b.push_back(x.row_begin); b.push_back(x.row_end); b.push_back(x.n_colours); }
        return give(b, dst, cap);
    }
    if (w == "solve_prof") {
        std::vector<long long> v(std::max((size_t)h->dev.sm_count * 4 * 8, (5 * p.tplan.mma.size() + 6 * p.tplan.epi.size())), 0);
        if (h->dev.solve_prof) cudaMemcpy(v.data(), h->dev.solve_prof, v.size() * 8, cudaMemcpyDeviceToHost);
        return give(v, dst, cap);
    }
    if (w == "stats") {
        std::vector<long long> s = {p.prog.n_slots, p.prog.n_phases_fwd, p.prog.n_steps_fwd, p.prog.n_steps_bwd,
                                    p.prog.n_entries, (long long)p.prog.stage_off.size() - 1,
                                    (long long)p.prog.bytes.size(), p.asmplan.max_eq_per_block,
                                    (long long)p.asmplan.eq_id.size(), (long long)p.asmplan.blocks.size(),
                                    (long long)solve_smem_bytes(p.prog.n_slots, p.prog.frames_per_tile), p.prog.n_supernodes,
                                    (long long)p.prog.io_desc.size(), p.prog.n_entries_padded, p.prog.frames_per_tile};
        return give(s, dst, cap);
    }
    return -1;
}
