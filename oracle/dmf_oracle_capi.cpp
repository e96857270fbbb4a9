// dmf_oracle_capi.cpp -- C entry points over dmf_oracle.hpp so that tests/, smoke() and bench.py's
// cpu_baseline / --impl reference legs can drive the CPU oracle through ctypes.
// TEST INFRASTRUCTURE, NOT PRODUCT CODE (see dmf_oracle.hpp header: pinned against
// oracle/_ref, the reference's own headers compiled here; PARITY UNPINNED only for Eigen's internal op order).
#include "dmf_oracle.hpp"
#include <chrono>
#include <cstdio>
#ifdef _OPENMP
#include <omp.h>
#endif

using namespace dmf_oracle;

namespace {
Affine pose_from12(const float* p) {   // row-major 3x4
    Affine T;
    for (int r = 0; r < 3; r++) for (int c = 0; c < 4; c++) T.m[r][c] = p[4*r+c];
    return T;
}
Camera cam_from(const float* K, int H, int W) { return Camera(std::vector<float>(K, K+9), H, W); }
void put_counters(const Counters& c, long long* out) {
    if (!out) return;
    out[0]=c.samples; out[1]=c.inbounds; out[2]=c.hits; out[3]=c.oob; out[4]=c.runaway;
}
}

extern "C" {

void orc_set_eigen_order(int o) { set_eigen_order(o); }

// setDimensions + setVolumeSize + constructVolume.  flat: 0 = reference pointer grid, 1 = flat index grid.
void* orc_volume_new(const double* bounds, const int* dims, int flat) {
    auto* v = new VoxelVolume(flat != 0);
    v->setDimensions(bounds[0],bounds[1],bounds[2],bounds[3],bounds[4],bounds[5]);
    v->setVolumeSize(dims[0],dims[1],dims[2]);
    v->constructVolume();
    return v;
}
void orc_volume_free(void* h) { delete (VoxelVolume*)h; }
void orc_volume_info(void* h, int* dims, double* deltas, double* voxel_size) {
    auto* v = (VoxelVolume*)h;
    dims[0]=v->xdim_; dims[1]=v->ydim_; dims[2]=v->zdim_;
    deltas[0]=v->xdelta_; deltas[1]=v->ydelta_; deltas[2]=v->zdelta_;
    *voxel_size = v->voxel_size_;
}
long orc_volume_integrate(void* h, const float* xyz, const float* nrm, long n) { return ((VoxelVolume*)h)->integrate(xyz, nrm, n); }
long orc_volume_num_occupied(void* h) { return (long)((VoxelVolume*)h)->occupied_cells_.size(); }
void orc_volume_get_occupied(void* h, u64* ids) { auto* v=(VoxelVolume*)h; std::copy(v->occupied_cells_.begin(), v->occupied_cells_.end(), ids); }
long orc_volume_num_normals(void* h) { long n=0; for (auto* p : ((VoxelVolume*)h)->pool_) n += (long)p->normals.size(); return n; }
// CSR of per-voxel normals in occupied_cells_ order: offsets[n_occ+1], normals[3*total]
void orc_volume_get_normals(void* h, unsigned* offsets, float* normals) {
    auto* v=(VoxelVolume*)h; unsigned o=0; size_t i=0;
    for (auto* p : v->pool_) { offsets[i++]=o; for (auto& n : p->normals) { normals[3*o]=n.v[0]; normals[3*o+1]=n.v[1]; normals[3*o+2]=n.v[2]; o++; } }
    offsets[i]=o;
}
void orc_volume_get_marks(void* h, int* view, unsigned char* good) {
    auto* v=(VoxelVolume*)h; size_t i=0;
    for (auto* p : v->pool_) { view[i]=p->view; good[i]=p->good?1:0; i++; }
}
void orc_volume_clear_marks(void* h) { for (auto* p : ((VoxelVolume*)h)->pool_) { p->view=0; p->good=false; } }

// One forward routine.  mode: 0 rayTraceAndGetPoints, 1 rayTraceAndGetGoodPoints, 2 rayTraceAndClassify,
// 3 rayTrace, 4 rayTraceAndGetMinimum.  Optional outputs may be NULL.  ids_out capacity = ids_cap.
// Returns number of ids the reference would return (may exceed ids_cap; only ids_cap are written).
long orc_forward(void* h, const float* K, int H, int W, const float* pose12, int mode, int zdelta, int sparse, int view,
                 int* depth_img, float* points, u64* hit_voxel, u64* ids_out, long ids_cap,
                 int* found_any, int* min_depth, long long* counters) {
    auto* vol = (VoxelVolume*)h;
    RayTracingEngine eng(cam_from(K,H,W));
    Counters c; PixelOut po; po.depth=depth_img; po.point=points; po.voxel=hit_voxel;
    bool want_po = depth_img || points || hit_voxel;
    int md = -1;
    auto res = eng.forward((RayTracingEngine::Mode)mode, *vol, pose_from12(pose12), zdelta, sparse != 0, view,
                           &md, counters ? &c : nullptr, want_po ? &po : nullptr);
    if (found_any) *found_any = res.first ? 1 : 0;
    if (min_depth) *min_depth = md;
    put_counters(c, counters);
    long n = (long)res.second.size();
    if (ids_out) std::copy(res.second.begin(), res.second.begin() + std::min(n, ids_cap), ids_out);
    return n;
}

// Carve-mode extension (PixelOut::observed; not in the reference): runs one forward routine and ORs the voxels of all visited
// in-bounds samples into `observed` (uint32 words over the padded index space, caller-zeroed or accumulated across calls).
// Returns the number of words the grid needs; with observed == NULL nothing is run.
long orc_forward_observed(void* h, const float* K, int H, int W, const float* pose12, int mode, int zdelta, int sparse, int view,
                          unsigned* observed, long long* counters) {
    auto* vol = (VoxelVolume*)h;
    size_t nbits = (size_t)(vol->xdim_+1)*(size_t)(vol->ydim_+1)*(size_t)(vol->zdim_+1);
    long nwords = (long)(((nbits + 31) / 32 + 7) / 8 * 8);
    if (!observed) return nwords;
    RayTracingEngine eng(cam_from(K,H,W));
    Counters c; PixelOut po; po.observed = observed;
    int md = -1;
    eng.forward((RayTracingEngine::Mode)mode, *vol, pose_from12(pose12), zdelta, sparse != 0, view, &md, counters ? &c : nullptr, &po);
    put_counters(c, counters);
    return nwords;
}

// reverseRayTraceFast (fast=1) / reverseRayTrace (fast=0).  visible_flags (n_occ bytes, fast only):
// bit0 = not occluded, bit1 = emitted as good.
long orc_reverse(void* h, const float* K, int H, int W, const float* pose12, int fast, int viz, int dead_work,
                 u64* ids_out, long ids_cap, int* found_any, unsigned char* visible_flags, long long* counters) {
    auto* vol = (VoxelVolume*)h;
    RayTracingEngine eng(cam_from(K,H,W));
    eng.dead_neighbor_work_ = dead_work != 0;
    Counters c; std::vector<char> flags;
    IdList res = fast ? eng.reverseRayTraceFast(*vol, pose_from12(pose12), viz != 0, 1, counters ? &c : nullptr, visible_flags ? &flags : nullptr)
                      : eng.reverseRayTrace(*vol, pose_from12(pose12), viz != 0, 1, counters ? &c : nullptr);
    if (found_any) *found_any = res.first ? 1 : 0;
    if (visible_flags && fast) std::copy(flags.begin(), flags.end(), visible_flags);
    put_counters(c, counters);
    long n = (long)res.second.size();
    if (ids_out) std::copy(res.second.begin(), res.second.begin() + std::min(n, ids_cap), ids_out);
    return n;
}

void orc_zbuffer(void* h, const float* K, int H, int W, const float* pose12, int* depth_img, long long* counter) {
    RayTracingEngine eng(cam_from(K,H,W));
    eng.rayTraceVolume(*(VoxelVolume*)h, pose_from12(pose12), depth_img, counter);
}

// Affine3f::inverse() restated (E5): out12 row-major 3x4.
void orc_affine_inverse(const float* pose12, float* out12) {
    Affine r = pose_from12(pose12).inverse();
    for (int i = 0; i < 3; i++) for (int j = 0; j < 4; j++) out12[4*i+j] = r.m[i][j];
}

// The host libm's float acos composed with degree() -- exposes the exact predicate value the oracle uses.
int orc_degree_acosf(float d) { return degree(std::acos(d)); }

// greedySetCover over CSR sets (sorted ids).  Returns number selected; selected_out capacity n_sets.
long orc_greedy_set_cover(const u64* ids, const long long* offsets, long n_sets, u64* selected_out) {
    std::vector<std::vector<u64>> sets(n_sets);
    for (long i = 0; i < n_sets; i++) sets[i].assign(ids + offsets[i], ids + offsets[i+1]);
    auto sel = greedySetCover(sets);
    std::copy(sel.begin(), sel.end(), selected_out);
    return (long)sel.size();
}

int orc_will_collide(void* h, const float* a, const float* b, int guard_coords, long long* steps) {
    Vec3 va = {{a[0],a[1],a[2]}}, vb = {{b[0],b[1],b[2]}};
    return willCollide(*(VoxelVolume*)h, va, vb, guard_coords != 0, steps) ? 1 : 0;
}

// optimizeCameraPosition(volume, engine, res, camera): returns mid; out12 = repositioned camera
unsigned orc_optimize_standoff(void* h, const float* K, int H, int W, const float* pose12, float* out12) {
    RayTracingEngine eng(cam_from(K,H,W));
    unsigned mid = 0;
    Affine r = optimizeCameraPosition(*(VoxelVolume*)h, eng, pose_from12(pose12), &mid);
    for (int i = 0; i < 3; i++) for (int j = 0; j < 4; j++) out12[4*i+j] = r.m[i][j];
    return mid;
}

// Timed CPU baseline: n_views poses (12 floats each) through one forward/reverse routine, no instrumentation,
// no per-pixel outputs -- i.e. exactly the reference's work.  threads = 1 is the reference's own (serial) path;
// threads > 1 distributes views over OpenMP threads (read-only modes only: 0, 1, 4 and reverse with viz = 0).
// kind: 0..4 forward mode, 10 = reverseRayTraceFast, 11 = reverseRayTraceFast incl. dead neighbour work,
// 12 = reverseRayTrace.  Returns seconds; total ids returned -> *n_ids_total (keeps the work observable).
double orc_time_views(void* h, const float* K, int H, int W, const float* poses, long n_views, int kind,
                      int zdelta, int sparse, int threads, long long* n_ids_total) {
    auto* vol = (VoxelVolume*)h;
    RayTracingEngine eng(cam_from(K,H,W));
    eng.dead_neighbor_work_ = (kind == 11);
    long long total = 0;
    auto t0 = std::chrono::steady_clock::now();
#ifdef _OPENMP
    #pragma omp parallel for schedule(dynamic,1) num_threads(threads > 0 ? threads : 1) reduction(+:total)
#endif
    for (long i = 0; i < n_views; i++) {
        Affine T = pose_from12(poses + 12*i);
        if (kind <= 4) {
            int md = -1;
            auto r = eng.forward((RayTracingEngine::Mode)kind, *vol, T, zdelta, sparse != 0, 1, &md, nullptr, nullptr);
            total += (long long)r.second.size() + (md >= 0 ? md : 0);
        } else if (kind == 12) {
            total += (long long)eng.reverseRayTrace(*vol, T, false).second.size();
        } else {
            total += (long long)eng.reverseRayTraceFast(*vol, T, false).second.size();
        }
    }
    auto t1 = std::chrono::steady_clock::now();
    if (n_ids_total) *n_ids_total = total;
    return std::chrono::duration<double>(t1 - t0).count();
}

int orc_max_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

}  // extern "C"
