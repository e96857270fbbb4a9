// ref_capi.cpp -- TEST INFRASTRUCTURE.  C entry points over the reference's OWN hot-path headers, compiled from where they
// lie (/root/reference/include, read-only, never copied) against oracle/ref_shim (a minimal Eigen/PCL look-alike), into
// oracle/_ref/libref_dmf.so.  Used to (1) validate the restatement in dmf_oracle.hpp against the real reference code and
// (2) time the reference's CPU path (bench.py cpu_baseline kind "reference").  The reference source cannot travel to the
// GPU box; the built library does.
//
// Same C API shape as dmf_oracle_capi.cpp (subset), prefix ref_.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <chrono>
#include <memory>
#include <vector>
#include <tuple>
#include <algorithm>
#include <Eigen/Dense>
using namespace Eigen;   // the reference's CommonUtilities.hpp names Vector3f unqualified; its drivers get this from other headers

// The reference's integratePointCloud overloads fall off the end of a bool function (Volume.hpp:197,228): undefined
// behaviour that g++ -O3 turns into an unreachable.  Compile Volume.hpp without optimisation so the call returns.
#pragma GCC push_options
#pragma GCC optimize("O0", "no-unreachable-traps")
#include <Volume.hpp>
#pragma GCC pop_options
#include <RayTracingEngine.hpp>
// The consumers next to the hot path: greedySetCover, positionCamera(s), generateSphere, repositionCamera,
// optimizeCameraPosition -- the reference's Algorithms.hpp, unmodified.  Its PCL I/O / viewer includes resolve to empty
// stand-ins and its TransformationUtilities.hpp to a one-function stand-in (oracle/ref_shim); it relies on its drivers
// having included <numeric> / <iterator> / <iostream> and the engine headers before it.
#include <numeric>
#include <iterator>
#include <iostream>
#include <Algorithms.hpp>
// (FileRoutines.hpp -- read/writeCameraLocations, the pose-file wire format of SURVEY 8f-4 -- arrives through the driver below)
// The driver tests/CameraPathGen.cpp itself, for the three functions it defines next to the path: willCollide (:128-156),
// repositionCamerasSampled (:94-126) and setCover (:158-181).  Its viewer / TSP / PCL-filter headers resolve to stand-ins in
// oracle/ref_shim, and its main() is renamed away.
#include <climits>
#include <mutex>
#define main ref_unused_camerapathgen_main
#include <CameraPathGen.cpp>
#undef main
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {
Eigen::Affine3f pose_from12(const float* p) {
    Eigen::Affine3f T = Eigen::Affine3f::Identity();
    for (int r = 0; r < 3; r++) for (int c = 0; c < 4; c++) T(r, c) = p[4 * r + c];
    return T;
}
// FINDING: the forward routines declare `bool found[H][W]={false};` on a variable-length array (RayTracingEngine.hpp:272,
// 315,381,451).  g++ zero-initialises only the first element of a VLA; the rest is whatever the stack held (checked with
// g++ 13.3: 54 315 of 307 200 entries non-zero after a call that dirtied the stack).  The reference's results therefore
// depend on stack garbage (UB).  The intended semantics -- no pixel found yet -- is what the oracle and the CUDA path
// implement; to compare against the real source we hand it a clean stack: zero a region below the current frame right
// before every call, so the VLA the callee carves out of it reads as all-false.
__attribute__((noinline)) void scrub_stack() {
    volatile char pad[3 << 20];
    for (size_t i = 0; i < sizeof pad; i += 64) pad[i] = 0;          // touch
    std::memset((void*)pad, 0, sizeof pad);
    asm volatile("" ::: "memory");
}
struct Quiet {   // the reference prints inside the hot path (RayTracingEngine.hpp:224,537-538)
    std::streambuf* old;
    Quiet() : old(std::cout.rdbuf(nullptr)) {}
    ~Quiet() { std::cout.rdbuf(old); }
};
}

extern "C" {

void* ref_volume_new(const double* bounds, const int* dims) {
    auto* v = new VoxelVolume();
    v->setDimensions(bounds[0], bounds[1], bounds[2], bounds[3], bounds[4], bounds[5]);
    v->setVolumeSize(dims[0], dims[1], dims[2]);
    v->constructVolume();
    return v;
}
void ref_volume_free(void* h) { delete (VoxelVolume*)h; }
void ref_volume_info(void* h, int* dims, double* deltas, double* voxel_size) {
    auto* v = (VoxelVolume*)h;
    dims[0] = v->xdim_; dims[1] = v->ydim_; dims[2] = v->zdim_;
    deltas[0] = v->xdelta_; deltas[1] = v->ydelta_; deltas[2] = v->zdelta_;
    *voxel_size = v->voxel_size_;
}
void ref_volume_integrate(void* h, const float* xyz, const float* nrm, long n) {
    auto* v = (VoxelVolume*)h;
    pcl::PointCloud<pcl::PointXYZRGB>::Ptr cloud(new pcl::PointCloud<pcl::PointXYZRGB>);
    pcl::PointCloud<pcl::Normal>::Ptr normals(new pcl::PointCloud<pcl::Normal>);
    for (long i = 0; i < n; i++) {
        pcl::PointXYZRGB p; p.x = xyz[3 * i]; p.y = xyz[3 * i + 1]; p.z = xyz[3 * i + 2];
        cloud->points.push_back(p);
        pcl::Normal q; q.normal[0] = nrm[3 * i]; q.normal[1] = nrm[3 * i + 1]; q.normal[2] = nrm[3 * i + 2];
        normals->points.push_back(q);
    }
    v->integratePointCloud(cloud, normals);
}
long ref_volume_num_occupied(void* h) { return (long)((VoxelVolume*)h)->occupied_cells_.size(); }
void ref_volume_get_occupied(void* h, unsigned long long* ids) { auto* v = (VoxelVolume*)h; std::copy(v->occupied_cells_.begin(), v->occupied_cells_.end(), ids); }
void ref_volume_get_marks(void* h, int* view, unsigned char* good) {
    auto* v = (VoxelVolume*)h; size_t i = 0;
    for (auto id : v->occupied_cells_) { int x, y, z; std::tie(x, y, z) = v->getVoxelCoords(id); Voxel* q = v->voxels_[x][y][z]; view[i] = q->view; good[i] = q->good ? 1 : 0; i++; }
}
void ref_volume_clear_marks(void* h) {
    auto* v = (VoxelVolume*)h;
    for (auto id : v->occupied_cells_) { int x, y, z; std::tie(x, y, z) = v->getVoxelCoords(id); v->voxels_[x][y][z]->view = 0; v->voxels_[x][y][z]->good = false; }
}

// mode: 0 rayTraceAndGetPoints, 1 rayTraceAndGetGoodPoints, 2 rayTraceAndClassify, 3 rayTrace, 4 rayTraceAndGetMinimum
long ref_forward(void* h, const float* K, int H, int W, const float* pose12, int mode, int zdelta, int sparse, int view,
                 unsigned long long* ids_out, long ids_cap, int* found_any, int* min_depth) {
    Quiet q;
    std::vector<float> Kv(K, K + 9);
    Camera cam(Kv, H, W);
    RayTracingEngine eng(cam);
    Eigen::Affine3f T = pose_from12(pose12);
    VoxelVolume& vol = *(VoxelVolume*)h;
    std::pair<bool, std::vector<unsigned long long int>> res(false, {});
    int md = -1;
    scrub_stack();
    switch (mode) {
        case 0: res = eng.rayTraceAndGetPoints(vol, T, zdelta, sparse != 0); break;
        case 1: res = eng.rayTraceAndGetGoodPoints(vol, T, zdelta, sparse != 0); break;
        case 2: eng.rayTraceAndClassify(vol, T, zdelta, view, sparse != 0); break;
        case 3: eng.rayTrace(vol, T, zdelta, sparse != 0); break;
        default: md = eng.rayTraceAndGetMinimum(vol, T, zdelta, sparse != 0); res.first = md >= 0; break;
    }
    if (found_any) *found_any = res.first ? 1 : 0;
    if (min_depth) *min_depth = md;
    long n = (long)res.second.size();
    if (ids_out) std::copy(res.second.begin(), res.second.begin() + std::min(n, ids_cap), ids_out);
    return n;
}

long ref_reverse(void* h, const float* K, int H, int W, const float* pose12, int fast, int viz,
                 unsigned long long* ids_out, long ids_cap, int* found_any) {
    Quiet q;
    std::vector<float> Kv(K, K + 9);
    Camera cam(Kv, H, W);
    RayTracingEngine eng(cam);
    VoxelVolume& vol = *(VoxelVolume*)h;
    auto res = fast ? eng.reverseRayTraceFast(vol, pose_from12(pose12), viz != 0) : eng.reverseRayTrace(vol, pose_from12(pose12), viz != 0);
    if (found_any) *found_any = res.first ? 1 : 0;
    long n = (long)res.second.size();
    if (ids_out) std::copy(res.second.begin(), res.second.begin() + std::min(n, ids_cap), ids_out);
    return n;
}

void ref_zbuffer(void* h, const float* K, int H, int W, const float* pose12) {
    Quiet q;
    std::vector<float> Kv(K, K + 9);
    Camera cam(Kv, H, W);
    RayTracingEngine eng(cam);
    Eigen::Affine3f T = pose_from12(pose12);
    eng.rayTraceVolume(*(VoxelVolume*)h, T);
}

// seconds for n_views calls of one routine; kind 0..4 forward mode, 10 reverseRayTraceFast, 12 reverseRayTrace.
// threads > 1 spreads views over OpenMP threads (read-only kinds 0, 1, 4, 10, 12 with viz = false).
double ref_time_views(void* h, const float* K, int H, int W, const float* poses, long n_views, int kind, int zdelta, int sparse, int threads, long long* total_out) {
    Quiet q;
    std::vector<float> Kv(K, K + 9);
    Camera cam(Kv, H, W);
    RayTracingEngine eng(cam);
    VoxelVolume& vol = *(VoxelVolume*)h;
    long long total = 0;
    auto t0 = std::chrono::steady_clock::now();
#ifdef _OPENMP
    #pragma omp parallel for schedule(dynamic,1) num_threads(threads > 0 ? threads : 1) reduction(+:total)
#endif
    for (long i = 0; i < n_views; i++) {
        Eigen::Affine3f T = pose_from12(poses + 12 * i);
        scrub_stack();
        if (kind == 0) total += (long long)eng.rayTraceAndGetPoints(vol, T, zdelta, sparse != 0).second.size();
        else if (kind == 1) total += (long long)eng.rayTraceAndGetGoodPoints(vol, T, zdelta, sparse != 0).second.size();
        else if (kind == 4) total += eng.rayTraceAndGetMinimum(vol, T, zdelta, sparse != 0);
        else if (kind == 12) total += (long long)eng.reverseRayTrace(vol, T, false).second.size();
        else total += (long long)eng.reverseRayTraceFast(vol, T, false).second.size();
    }
    auto t1 = std::chrono::steady_clock::now();
    if (total_out) *total_out = total;
    return std::chrono::duration<double>(t1 - t0).count();
}

// ---- Algorithms.hpp -------------------------------------------------------------------------------------------------
// greedySetCover(candidate_sets) over CSR sets (sorted ids, Algorithms.hpp:38-86).  Returns the number selected.
long ref_greedy_set_cover(const unsigned long long* ids, const long long* offsets, long n_sets, unsigned long long* selected_out) {
    Quiet q;
    std::vector<std::vector<unsigned long long int>> sets(n_sets);
    for (long i = 0; i < n_sets; i++) sets[i].assign(ids + offsets[i], ids + offsets[i + 1]);
    auto sel = Algorithms::greedySetCover(sets);
    std::copy(sel.begin(), sel.end(), selected_out);
    return (long)sel.size();
}

// positionCameras(locations, distance) (Algorithms.hpp:282-298 -> positionCamera :190-236): n surface points + normals
// -> n poses (12 floats each, row-major 3x4).  The normal flip of :286-292 is applied to a private copy.
void ref_position_cameras(const float* xyz, const float* normals, long n, unsigned distance, float* poses12) {
    Quiet q;
    pcl::PointCloud<pcl::PointXYZRGBNormal>::Ptr loc(new pcl::PointCloud<pcl::PointXYZRGBNormal>);
    for (long i = 0; i < n; i++) {
        pcl::PointXYZRGBNormal p;
        p.x = xyz[3 * i]; p.y = xyz[3 * i + 1]; p.z = xyz[3 * i + 2];
        p.normal[0] = normals[3 * i]; p.normal[1] = normals[3 * i + 1]; p.normal[2] = normals[3 * i + 2];
        loc->points.push_back(p);
    }
    auto cams = Algorithms::positionCameras(loc, distance);
    for (long i = 0; i < n; i++) for (int r = 0; r < 3; r++) for (int c = 0; c < 4; c++) poses12[12 * i + 4 * r + c] = cams[i](r, c);
}

// generateSphere(radius, sphere, transformation, z_threshold, factor) (Algorithms.hpp:88-112).  Returns the number of
// points generated; writes at most cap of them.
long ref_generate_sphere(double radius, const float* T12, double z_threshold, double factor, float* xyz_out, long cap) {
    Quiet q;
    pcl::PointCloud<pcl::PointXYZRGB>::Ptr sphere(new pcl::PointCloud<pcl::PointXYZRGB>);
    Algorithms::generateSphere(radius, sphere, pose_from12(T12), z_threshold, factor);
    const long n = (long)sphere->points.size();
    for (long i = 0; i < n && i < cap; i++) { xyz_out[3 * i] = sphere->points[i].x; xyz_out[3 * i + 1] = sphere->points[i].y; xyz_out[3 * i + 2] = sphere->points[i].z; }
    return n;
}

// repositionCamera(camera, distance) (Algorithms.hpp:180-188 -> moveCamera :170-178)
void ref_reposition_camera(const float* pose12, unsigned distance, float* out12) {
    Eigen::Affine3f r = Algorithms::repositionCamera(pose_from12(pose12), distance);
    for (int i = 0; i < 3; i++) for (int j = 0; j < 4; j++) out12[4 * i + j] = r(i, j);
}

// optimizeCameraPosition(volume, engine, resolution, camera) (Algorithms.hpp:394-421): the binary search over the stand-off
// with two reverseRayTrace casts per step.  out12 = the repositioned camera it returns.
void ref_optimize_camera_position(void* h, const float* K, int H, int W, const float* pose12, float* out12) {
    Quiet q;
    std::vector<float> Kv(K, K + 9);
    Camera cam(Kv, H, W);
    RayTracingEngine engine(cam);
    Eigen::Affine3f r = Algorithms::optimizeCameraPosition(*(VoxelVolume*)h, engine, 1, pose_from12(pose12));
    for (int i = 0; i < 3; i++) for (int j = 0; j < 4; j++) out12[4 * i + j] = r(i, j);
}

// ---- FileRoutines.hpp: the camera-pose text files the drivers exchange (:69-112) ------------------------------------------
void ref_write_camera_locations(const char* filename, const float* poses12, long n) {
    std::vector<Eigen::Affine3f> T;
    for (long i = 0; i < n; i++) T.push_back(pose_from12(poses12 + 12 * i));
    writeCameraLocations(filename, T);
}
// returns the number of poses in the file; writes at most cap of them
long ref_read_camera_locations(const char* filename, float* poses12, long cap) {
    Quiet q;
    auto T = readCameraLocations(filename);
    for (long i = 0; i < (long)T.size() && i < cap; i++) for (int r = 0; r < 3; r++) for (int c = 0; c < 4; c++) poses12[12 * i + 4 * r + c] = T[i](r, c);
    return (long)T.size();
}

// ---- tests/CameraPathGen.cpp: the driver's own helpers next to the path -------------------------------------------------------
// willCollide(volume, a, b) (:128-156): 1 mm float march from a towards b, true at the first occupied voxel
int ref_will_collide(void* h, const float* a, const float* b) {
    Quiet q;
    return willCollide(*(VoxelVolume*)h, Eigen::Vector3f(a[0], a[1], a[2]), Eigen::Vector3f(b[0], b[1], b[2])) ? 1 : 0;
}
// repositionCamerasSampled(cameras, volume, cam) (:94-126): rayTraceAndGetMinimum per camera, then 0.3 m back from the hit
void ref_reposition_cameras_sampled(void* h, const float* K, int H, int W, const float* poses12, long n, float* out12) {
    Quiet q;
    std::vector<float> Kv(K, K + 9);
    Camera cam(Kv, H, W);
    std::vector<Eigen::Affine3f> cams;
    for (long i = 0; i < n; i++) cams.push_back(pose_from12(poses12 + 12 * i));
    scrub_stack();
    auto out = repositionCamerasSampled(cams, *(VoxelVolume*)h, cam);
    for (long i = 0; i < n; i++) for (int r = 0; r < 3; r++) for (int c = 0; c < 4; c++) out12[12 * i + 4 * r + c] = out[i](r, c);
}
// setCover(engine, volume, camera_locations, resolution, sparse) (:158-181): reverseRayTraceFast per camera, sort, greedySetCover
long ref_set_cover(void* h, const float* K, int H, int W, const float* poses12, long n, unsigned long long* selected_out) {
    Quiet q;
    std::vector<float> Kv(K, K + 9);
    Camera cam(Kv, H, W);
    RayTracingEngine engine(cam);
    std::vector<Eigen::Affine3f> cams;
    for (long i = 0; i < n; i++) cams.push_back(pose_from12(poses12 + 12 * i));
    auto sel = setCover(engine, *(VoxelVolume*)h, cams, 1, false);
    std::copy(sel.begin(), sel.end(), selected_out);
    return (long)sel.size();
}

int ref_max_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

}  // extern "C"
