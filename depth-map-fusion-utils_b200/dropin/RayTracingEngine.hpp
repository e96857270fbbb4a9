// RayTracingEngine.hpp -- drop-in for the reference's include/RayTracingEngine.hpp (class RayTracingEngine :27-40).
//
// Same class, same public member `cam_`, same eight method signatures, same default arguments on the out-of-class
// definitions (:45,136,229,268,311,377,447,498), so tests/*.cpp and include/Algorithms.hpp compile and relink unchanged.
// Every method body forwards to libdmf_b200.so (include/dmf_b200.h): the per-pixel / per-voxel marches run as sm_100a
// CUDA kernels.  There is no host fallback: if the library cannot reach a B200 the process aborts with the library's
// error message (the reference has no error convention a caller could handle).
//
// The engine is passed BY VALUE all over the reference (tests/SetCover.cpp:218, include/Algorithms.hpp:364), so it owns
// no device state; that lives in one process-global context, and the device mirror of a VoxelVolume is refreshed
// whenever (&volume, volume.revision()) changes.  Like the reference, this header does not include Volume.hpp itself:
// include <Volume.hpp> first.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <memory>
#include <utility>
#include <vector>
#include <Eigen/Dense>
#include <Eigen/Core>

#include <Camera.hpp>
#include <dmf_b200.h>

using namespace Eigen;
using namespace std;

constexpr double k_AngleMin = 0;
constexpr double k_AngleMax = 90;
constexpr double k_ZMin = 0.20;
constexpr double k_ZMax = 1.0;

namespace dmf_dropin {

inline void must(int rc, const char* what)
{
    if (rc != 0) { std::fprintf(stderr, "libdmf_b200: %s failed: %s\n", what, dmf_last_error()); std::abort(); }
}

// One per process.  By default it forms a group over every visible GPU (dmf_comm_init_all: this one host thread drives them
// all): single-view calls run on GPU 0, the per-view loops of the drivers (setCover) are dealt over all of them.
// DMF_GPUS=n limits the group to the first n devices; DMF_DEVICE=d pins everything to one device (no group).
struct Global {
    dmf_ctx* ctx = nullptr;
    dmf_comm* comm = nullptr;
    const void* volume = nullptr;
    unsigned long long revision = ~0ull;
    unsigned long long replicated_revision = ~0ull;     // revision the other GPUs of the group hold
    const void* replicated_volume = nullptr;
    Global()
    {
        const char* dev = std::getenv("DMF_DEVICE");
        const char* gpus = std::getenv("DMF_GPUS");
        if (dev) { must(dmf_create(&ctx, std::atoi(dev)), "dmf_create"); return; }
        must(dmf_comm_init_all(&comm, gpus ? std::atoi(gpus) : 0), "dmf_comm_init_all");
        ctx = dmf_comm_ctx(comm, 0);
    }
    ~Global() { if (comm) dmf_comm_destroy(comm); else dmf_destroy(ctx); }
    int gpus() const
    {
        int world = 1;
        if (comm) dmf_comm_info(comm, &world, nullptr, nullptr, nullptr);
        return world;
    }
};
inline Global& global() { static Global g; return g; }

inline void pose12(const Eigen::Affine3f& T, float out[12])
{
    for (int r = 0; r < 3; r++) for (int c = 0; c < 4; c++) out[4 * r + c] = T(r, c);
}

// make the device mirror of the volume current (re-upload only when (&volume, revision) changed)
template <class Volume>
inline dmf_ctx* sync_volume(Volume& volume)
{
    Global& g = global();
    if (g.volume != (const void*)&volume || g.revision != volume.revision()) {
        const double bounds[6] = {volume.xmin_, volume.xmax_, volume.ymin_, volume.ymax_, volume.zmin_, volume.zmax_};
        const double delta[3] = {volume.xdelta_, volume.ydelta_, volume.zdelta_};
        const int dim[3] = {volume.xdim_, volume.ydim_, volume.zdim_};
        std::vector<uint32_t> offsets; std::vector<float> normals;
        volume.exportNormals(offsets, normals);
        std::vector<uint64_t> ids(volume.occupied_cells_.begin(), volume.occupied_cells_.end());
        must(dmf_upload_volume(g.ctx, bounds, delta, dim, ids.data(), ids.size(), offsets.data(), normals.data()), "dmf_upload_volume");
        g.volume = &volume; g.revision = volume.revision();
    }
    return g.ctx;
}

// the group's other GPUs hold the current volume and camera (GPU 0 -> peers, GPU to GPU); no-op for a single device
template <class Volume>
inline dmf_comm* sync_group(Camera& cam, Volume& volume)
{
    Global& g = global();
    sync_volume(volume);
    if (!g.comm) return nullptr;
    must(dmf_comm_set_camera(g.comm, cam.intrinsics().data(), cam.getHeight(), cam.getWidth()), "dmf_comm_set_camera");
    if (g.replicated_volume != (const void*)&volume || g.replicated_revision != volume.revision()) {
        must(dmf_comm_replicate_volume(g.comm, 0), "dmf_comm_replicate_volume");
        g.replicated_volume = &volume; g.replicated_revision = volume.revision();
    }
    return g.comm;
}

// volume as above + camera intrinsics every call (cheap)
template <class Volume>
inline dmf_ctx* sync(Camera& cam, Volume& volume)
{
    must(dmf_set_camera(global().ctx, cam.intrinsics().data(), cam.getHeight(), cam.getWidth()), "dmf_set_camera");
    return sync_volume(volume);
}

// Voxel::view / Voxel::good live on the host objects in the reference; push them before and pull them after a call
// that may change them, so host readers (VisualizationUtilities.hpp:398-405) see exactly what the reference would leave.
template <class Volume>
inline void push_marks(dmf_ctx* ctx, Volume& volume)
{
    const size_t n = volume.occupied_cells_.size();
    std::vector<int32_t> view(n); std::vector<uint8_t> good(n);
    for (size_t i = 0; i < n; i++) { auto* v = volume.voxelOf(i); view[i] = v->view; good[i] = v->good ? 1 : 0; }
    must(dmf_upload_marks(ctx, view.data(), good.data(), n), "dmf_upload_marks");
}
template <class Volume>
inline void pull_marks(dmf_ctx* ctx, Volume& volume)
{
    const size_t n = volume.occupied_cells_.size();
    std::vector<int32_t> view(n); std::vector<uint8_t> good(n);
    must(dmf_download_marks(ctx, view.data(), good.data(), n), "dmf_download_marks");
    for (size_t i = 0; i < n; i++) { auto* v = volume.voxelOf(i); v->view = view[i]; v->good = good[i] != 0; }
}

template <class Volume>
inline std::pair<bool, std::vector<unsigned long long int>> forward_ids(Camera& cam, Volume& volume, Eigen::Affine3f& T, int mode, int zdelta, bool sparse)
{
    dmf_ctx* ctx = sync(cam, volume);
    float pose[12]; pose12(T, pose);
    dmf_forward_params p = {mode, zdelta, sparse ? 1 : 0, 1, DMF_GRID_AUTO, 0};
    // a view returns each voxel at most once and at most one id per pixel; the buffer is left uninitialised (zero-filling
    // 2.4 MB per call cost more than the march of one VGA view)
    const size_t cap = std::min((size_t)cam.getHeight() * cam.getWidth(), volume.occupied_cells_.size()) + 1;
    std::unique_ptr<uint64_t[]> ids(new uint64_t[cap]);
    int64_t offsets[2] = {0, 0};
    int32_t found = 0;
    dmf_forward_out out = {};
    out.found_any = &found; out.ids = ids.get(); out.ids_offsets = offsets; out.ids_capacity = cap;
    must(dmf_forward(ctx, &p, pose, 1, &out), "dmf_forward");
    return std::make_pair(found != 0, std::vector<unsigned long long int>(ids.get(), ids.get() + offsets[1]));
}

template <class Volume>
inline void forward_marks(Camera& cam, Volume& volume, Eigen::Affine3f& T, int mode, int zdelta, int view, bool sparse)
{
    dmf_ctx* ctx = sync(cam, volume);
    push_marks(ctx, volume);
    float pose[12]; pose12(T, pose);
    dmf_forward_params p = {mode, zdelta, sparse ? 1 : 0, view, DMF_GRID_AUTO, 0};
    dmf_forward_out out = {};
    must(dmf_forward(ctx, &p, pose, 1, &out), "dmf_forward");
    pull_marks(ctx, volume);
}

template <class Volume>
inline std::pair<bool, std::vector<unsigned long long int>> reverse_ids(Camera& cam, Volume& volume, const Eigen::Affine3f& T, bool fast, bool viz)
{
    dmf_ctx* ctx = sync(cam, volume);
    if (viz) push_marks(ctx, volume);
    float pose[12]; pose12(T, pose);
    const size_t cap = 2 * volume.occupied_cells_.size() + 65;      // (the drifting whole-grid scan may visit a voxel twice)
    std::unique_ptr<uint64_t[]> ids(new uint64_t[cap]);
    int64_t offsets[2] = {0, 0};
    int32_t found = 0;
    dmf_reverse_out out = {};
    out.found_any = &found; out.ids = ids.get(); out.ids_offsets = offsets; out.ids_capacity = cap;
    must(dmf_reverse(ctx, fast ? 1 : 0, viz ? 1 : 0, pose, 1, &out), "dmf_reverse");
    if (viz) pull_marks(ctx, volume);
    return std::make_pair(found != 0, std::vector<unsigned long long int>(ids.get(), ids.get() + offsets[1]));
}

}  // namespace dmf_dropin

class RayTracingEngine
{
    public:
        Camera cam_;
        RayTracingEngine(Camera &cam);
        void rayTrace(VoxelVolume& volume,Eigen::Affine3f& transformation,int zdelta,bool sparse);
        void rayTraceAndClassify(VoxelVolume& volume,Eigen::Affine3f& transformation,int zdelta,int view,bool sparse);
        void rayTraceVolume(VoxelVolume& volume,Eigen::Affine3f& transformation);
        int rayTraceAndGetMinimum(VoxelVolume& volume,Eigen::Affine3f& transformation,int zdelta,bool sparse);
        std::pair<bool,std::vector<unsigned long long int>> rayTraceAndGetGoodPoints(VoxelVolume& volume,Eigen::Affine3f& transformation,int zdelta,bool sparse);
        std::pair<bool,std::vector<unsigned long long int>> rayTraceAndGetPoints(VoxelVolume& volume,Eigen::Affine3f& transformation,int zdelta,bool sparse);
        std::pair<bool,std::vector<unsigned long long int>> reverseRayTrace(VoxelVolume& volume, Eigen::Affine3f transformation,bool viz,int zdelta);
        std::pair<bool,std::vector<unsigned long long int>> reverseRayTraceFast(VoxelVolume& volume, Eigen::Affine3f transformation,bool viz,int zdelta);
};

inline RayTracingEngine::RayTracingEngine(Camera &cam):cam_(cam){}

// dmf_reverse(fast = 0): whole-grid scan, march from 1 mm, "good" = depth window only
inline std::pair<bool,std::vector<unsigned long long int>> RayTracingEngine::reverseRayTrace(VoxelVolume& volume, Eigen::Affine3f transformation,bool viz, int zdelta = 1)
{
    (void)zdelta;   // unused by the reference as well
    return dmf_dropin::reverse_ids(cam_, volume, transformation, false, viz);
}

// dmf_reverse(fast = 1): one march per occupied voxel, from 50 mm, normal test inside the 0.2..1.0 m window
inline std::pair<bool,std::vector<unsigned long long int>> RayTracingEngine::reverseRayTraceFast(VoxelVolume& volume, Eigen::Affine3f transformation,bool viz, int zdelta = 1)
{
    (void)zdelta;
    return dmf_dropin::reverse_ids(cam_, volume, transformation, true, viz);
}

inline int RayTracingEngine::rayTraceAndGetMinimum(VoxelVolume& volume,Eigen::Affine3f& transformation,int zdelta = 1,bool sparse=true)
{
    dmf_ctx* ctx = dmf_dropin::sync(cam_, volume);
    float pose[12]; dmf_dropin::pose12(transformation, pose);
    dmf_forward_params p = {DMF_MODE_MINIMUM, zdelta, sparse ? 1 : 0, 1, DMF_GRID_AUTO, 0};
    int32_t min_depth = -1;
    dmf_forward_out out = {};
    out.min_depth = &min_depth;
    dmf_dropin::must(dmf_forward(ctx, &p, pose, 1, &out), "dmf_forward");
    return min_depth;
}

inline void RayTracingEngine::rayTrace(VoxelVolume& volume,Eigen::Affine3f& transformation,int zdelta = 10,bool sparse=true)
{
    dmf_dropin::forward_marks(cam_, volume, transformation, DMF_MODE_MARK, zdelta, 1, sparse);
}

inline void RayTracingEngine::rayTraceAndClassify(VoxelVolume& volume,Eigen::Affine3f& transformation,int zdelta=10,int view = 1,bool sparse=true)
{
    dmf_dropin::forward_marks(cam_, volume, transformation, DMF_MODE_CLASSIFY, zdelta, view, sparse);
}

inline std::pair<bool,std::vector<unsigned long long int>> RayTracingEngine::rayTraceAndGetGoodPoints(VoxelVolume& volume,Eigen::Affine3f& transformation,int zdelta=10,bool sparse=true)
{
    return dmf_dropin::forward_ids(cam_, volume, transformation, DMF_MODE_GOOD_POINTS, zdelta, sparse);
}

inline std::pair<bool,std::vector<unsigned long long int>> RayTracingEngine::rayTraceAndGetPoints(VoxelVolume& volume,Eigen::Affine3f& transformation,int zdelta=10,bool sparse=true)
{
    return dmf_dropin::forward_ids(cam_, volume, transformation, DMF_MODE_POINTS, zdelta, sparse);
}

inline void RayTracingEngine::rayTraceVolume(VoxelVolume& volume,Eigen::Affine3f& transformation)
{
    dmf_ctx* ctx = dmf_dropin::sync(cam_, volume);
    dmf_dropin::push_marks(ctx, volume);
    float pose[12]; dmf_dropin::pose12(transformation, pose);
    dmf_dropin::must(dmf_zbuffer(ctx, pose, nullptr, nullptr), "dmf_zbuffer");
    dmf_dropin::pull_marks(ctx, volume);
}
