// DmfAlgorithms.hpp -- batched B200 replacements for the ray-marching helpers the reference keeps OUTSIDE
// RayTracingEngine: the driver-local willCollide() (tests/CameraPathGen.cpp:128-156, CameraMotionTSP.cpp:236-261,
// CameraMotionPlanner.cpp:246-271), the setCover() drivers (tests/SetCover.cpp:214-244, CameraPathGen.cpp:158-186) and
// Algorithms::optimizeCameraPosition(volume, engine, res, Affine3f) (include/Algorithms.hpp:394-421) and the driver-local
// repositionCamerasSampled() (tests/CameraPathGen.cpp:94-126).
//
// Same argument lists and results as the functions they replace (minus the stdout chatter), so a driver swaps
//     if (willCollide(volume, a, b) == true)          ->   if (dmf_dropin::willCollide(volume, a, b) == true)
// and the O(n^2) edge loops / per-camera loops can move to the batched forms below (one kernel launch for all
// segments / all cameras).  Include after <Volume.hpp> and <RayTracingEngine.hpp>.  No host fallback.
#pragma once
#include <cstdint>
#include <vector>
#include <RayTracingEngine.hpp>

namespace dmf_dropin {

// n segments a[i] -> b[i] in one launch.  guard_coords: the validCoords guard of CameraPathGen.cpp:147 (true) or
// the unguarded copies in CameraMotionTSP.cpp / CameraMotionPlanner.cpp (false).
template <class Volume>
inline std::vector<uint8_t> segmentsCollide(Volume& volume, const std::vector<Eigen::Vector3f>& a, const std::vector<Eigen::Vector3f>& b, bool guard_coords = true)
{
    dmf_ctx* ctx = sync_volume(volume);
    const size_t n = a.size() < b.size() ? a.size() : b.size();
    std::vector<float> fa(3 * n), fb(3 * n);
    for (size_t i = 0; i < n; i++) for (int k = 0; k < 3; k++) { fa[3 * i + k] = a[i](k); fb[3 * i + k] = b[i](k); }
    std::vector<uint8_t> out(n);
    if (n) must(dmf_segments_collide(ctx, fa.data(), fb.data(), (int)n, guard_coords ? 1 : 0, out.data()), "dmf_segments_collide");
    return out;
}

// bool willCollide(VoxelVolume& volume, Vector3f a, Vector3f b)   (tests/CameraPathGen.cpp:128)
template <class Volume>
inline bool willCollide(Volume& volume, Eigen::Vector3f a, Eigen::Vector3f b, bool guard_coords = true)
{
    return segmentsCollide(volume, std::vector<Eigen::Vector3f>{a}, std::vector<Eigen::Vector3f>{b}, guard_coords)[0] != 0;
}

// The drivers' edge-validity double loop (CameraPathGen.cpp:318-330: every ordered pair x != y of camera centres) as
// ONE call.  Row-major n x n, [x][y] = willCollide(volume, centre_x, centre_y); the diagonal is 0.
template <class Volume>
inline std::vector<uint8_t> collisionMatrix(Volume& volume, const std::vector<Eigen::Affine3f>& cameras, bool guard_coords = true)
{
    const size_t n = cameras.size();
    std::vector<Eigen::Vector3f> a, b;
    a.reserve(n * n); b.reserve(n * n);
    for (size_t x = 0; x < n; x++) for (size_t y = 0; y < n; y++) {
        if (x == y) continue;
        Eigen::Vector3f pa, pb;
        for (int k = 0; k < 3; k++) { pa(k) = cameras[x](k, 3); pb(k) = cameras[y](k, 3); }
        a.push_back(pa); b.push_back(pb);
    }
    std::vector<uint8_t> hit = segmentsCollide(volume, a, b, guard_coords), m(n * n, 0);
    size_t i = 0;
    for (size_t x = 0; x < n; x++) for (size_t y = 0; y < n; y++) if (x != y) m[x * n + y] = hit[i++];
    return m;
}

// Algorithms::optimizeCameraPosition for many cameras: every bisection step marches all still-active cameras at once.
template <class Volume>
inline std::vector<Eigen::Affine3f> optimizeCameraPositions(Volume& volume, RayTracingEngine engine, const std::vector<Eigen::Affine3f>& cameras,
                                                            unsigned low = 300, unsigned high = 600, std::vector<uint32_t>* mid_out = nullptr)
{
    dmf_ctx* ctx = sync(engine.cam_, volume);
    const size_t n = cameras.size();
    std::vector<float> in(12 * n), out(12 * n);
    std::vector<uint32_t> mid(n);
    for (size_t i = 0; i < n; i++) pose12(cameras[i], &in[12 * i]);
    if (n) must(dmf_optimize_standoff(ctx, in.data(), (int)n, low, high, mid.data(), out.data()), "dmf_optimize_standoff");
    std::vector<Eigen::Affine3f> res(n, Eigen::Affine3f::Identity());
    for (size_t i = 0; i < n; i++) for (int r = 0; r < 3; r++) for (int c = 0; c < 4; c++) res[i](r, c) = out[12 * i + 4 * r + c];
    if (mid_out) *mid_out = mid;
    return res;
}

// Affine3f optimizeCameraPosition(VoxelVolume&, RayTracingEngine, int resolution_single_dimension, Affine3f camera)
// (include/Algorithms.hpp:394; the resolution argument is unused there as well)
template <class Volume>
inline Eigen::Affine3f optimizeCameraPosition(Volume& volume, RayTracingEngine engine, int /*resolution_single_dimension*/, Eigen::Affine3f camera)
{
    return optimizeCameraPositions(volume, engine, std::vector<Eigen::Affine3f>{camera})[0];
}

// vector<Affine3f> repositionCamerasSampled(vector<Affine3f> cameras, VoxelVolume& volume, Camera cam)
// (tests/CameraPathGen.cpp:94-126): rayTraceAndGetMinimum for every camera as ONE batched cast (zdelta = 1, sparse: the
// defaults the driver uses), then the driver's own arithmetic per camera -- 0.3 m back from the nearest hit along the optical
// axis; a camera that hits nothing keeps its place.
template <class Volume>
inline std::vector<Eigen::Affine3f> repositionCamerasSampled(const std::vector<Eigen::Affine3f>& cameras, Volume& volume, Camera cam)
{
    const size_t n = cameras.size();
    std::vector<Eigen::Affine3f> new_locations(cameras);
    if (!n) return new_locations;
    dmf_ctx* ctx = sync(cam, volume);
    std::vector<float> poses(12 * n);
    for (size_t i = 0; i < n; i++) pose12(cameras[i], &poses[12 * i]);
    std::vector<int32_t> nearest(n, -1);
    dmf_forward_params p = {DMF_MODE_MINIMUM, 1, 1, 1, DMF_GRID_AUTO, 0};
    dmf_forward_out out = {};
    out.min_depth = nearest.data();
    must(dmf_forward(ctx, &p, poses.data(), (int)n, &out), "dmf_forward");
    for (size_t i = 0; i < n; i++) {
        if (nearest[i] == -1) continue;                                       // CameraPathGen.cpp:108-113
        const Eigen::Affine3f& camera = cameras[i];
        Eigen::Vector3f bz(camera(0,2), camera(1,2), camera(2,2));
        Eigen::Vector3f current_pt(camera(0,3), camera(1,3), camera(2,3));
        Eigen::Vector3f intersection_pt = current_pt + bz*(float(nearest[i])/1000.0);      // :116, scalar narrowed to float by Eigen
        Eigen::Vector3f new_point = intersection_pt - bz*(0.3);                             // :117
        for (int k = 0; k < 3; k++) new_locations[i](k,3) = new_point(k);
    }
    return new_locations;
}

// vector<unsigned long long int> setCover(RayTracingEngine engine, VoxelVolume& volume, vector<Affine3f> camera_locations,
//                                         int resolution_single_dimension, bool sparse = true)
// (tests/SetCover.cpp:214-244): reverseRayTraceFast for every location + Algorithms::greedySetCover.  The per-view loop
// (:218-240) is dealt over every GPU of the process's group (dmf_sweep_reverse: views r, r+N, ... on GPU r, the volume
// replicated GPU to GPU, finished visibility rows pushed to the peers by the march kernels), then the greedy loop runs on
// the device over the gathered rows.  Returns the selected location indices in selection order, like the reference.
template <class Volume>
inline std::vector<unsigned long long int> setCover(RayTracingEngine engine, Volume& volume, const std::vector<Eigen::Affine3f>& camera_locations,
                                                    int /*resolution_single_dimension*/ = 1, bool /*sparse*/ = true)
{
    dmf_ctx* ctx = sync(engine.cam_, volume);
    const size_t n = camera_locations.size();
    if (!n) return {};
    const size_t words = (size_t)dmf_visibility_words(ctx);
    std::vector<float> poses(12 * n);
    for (size_t i = 0; i < n; i++) pose12(camera_locations[i], &poses[12 * i]);
    if (dmf_comm* comm = sync_group(engine.cam_, volume)) {
        must(dmf_sweep_reverse(comm, 1, poses.data(), (int)n, nullptr), "dmf_sweep_reverse");
        std::vector<int32_t> selected(n); int n_selected = 0;
        must(dmf_sweep_set_cover(comm, selected.data(), &n_selected), "dmf_sweep_set_cover");
        return std::vector<unsigned long long int>(selected.begin(), selected.begin() + n_selected);
    }
    std::vector<uint64_t> vis(n * (words ? words : 1));
    dmf_reverse_out out = {};
    out.visibility = vis.data();
    must(dmf_reverse(ctx, 1, 0, poses.data(), (int)n, &out), "dmf_reverse");
    std::vector<int32_t> selected(n); int n_selected = 0;
    must(dmf_greedy_set_cover(ctx, vis.data(), (int)n, words, selected.data(), &n_selected), "dmf_greedy_set_cover");
    return std::vector<unsigned long long int>(selected.begin(), selected.begin() + n_selected);
}

}  // namespace dmf_dropin
