// dmf_host.cuh -- host-side state behind the C ABI: the context, grow-only device buffers, the host mirror of
// VoxelVolume used by dmf_volume_from_points, and the libm acosf bisection.
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>
#include "dmf_device.cuh"
#include "../../include/dmf_b200.h"

namespace dmf {

inline std::string& last_error() { static thread_local std::string e; return e; }
inline int fail(const char* fmt, ...) {
    char buf[1024];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    last_error() = buf;
    return 1;
}
#define DMF_CUDA(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return dmf::fail("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)
#define DMF_TRY(call) do { int r_ = (call); if (r_) return r_; } while (0)

// Bumped whenever a device or pinned buffer of the library moves: a captured CUDA graph holds raw pointers, so it is only
// replayed while this has not changed since its capture.
inline std::atomic<unsigned long long>& alloc_generation() { static std::atomic<unsigned long long> g{1}; return g; }   // (contexts may live on different host threads)

// grow-only device allocation
struct DevBuf {
    void* p = nullptr; size_t cap = 0;
    int reserve(size_t bytes) {
        if (bytes <= cap) return 0;
        alloc_generation()++;
        if (p) { cudaFree(p); p = nullptr; cap = 0; }
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { p = nullptr; return fail("cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e)); }
        cap = want;
        return 0;
    }
    void release() { if (p) { cudaFree(p); alloc_generation()++; } p = nullptr; cap = 0; }
    template <class T> T* as() const { return (T*)p; }
};

// grow-only pinned host staging (small results leave the device through it: one asynchronous copy, one synchronisation)
struct HostStage {
    void* p = nullptr; size_t cap = 0;
    int reserve(size_t bytes) {
        if (bytes <= cap) return 0;
        alloc_generation()++;
        if (p) { cudaFreeHost(p); p = nullptr; cap = 0; }
        size_t want = bytes + bytes / 4 + 4096;
        cudaError_t e = cudaMallocHost(&p, want);
        if (e != cudaSuccess) { p = nullptr; return fail("cudaMallocHost(%zu) failed: %s", want, cudaGetErrorString(e)); }
        cap = want;
        return 0;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

// Host restatement of the data-model half of VoxelVolume (reference include/Volume.hpp:89-128,135-170,199-233);
// used only to turn a point cloud into (occupied ids, normals CSR) -- the march itself never runs on the host.
struct HostVolume {
    double vmin[3], vmax[3], delta[3];
    int dim[3];
    std::vector<uint64_t> occupied;                     // occupied_cells_
    std::vector<std::vector<float>> normals;            // per occupied voxel: flat xyz list
    std::unordered_map<uint64_t, uint32_t> index;       // id -> ordinal

    void construct(const double b[6], const int dims[3]) {
        for (int a = 0; a < 3; a++) {
            vmin[a] = b[2 * a]; vmax[a] = b[2 * a + 1];
            delta[a] = (vmax[a] - vmin[a]) / dims[a];             // setVolumeSize  :114-116
            dim[a] = (int)((vmax[a] - vmin[a]) / delta[a]);       // constructVolume :121-123 (truncation)
        }
    }
    bool valid_point(float x, float y, float z) const {           // validPoints :230-233
        return !(x >= vmax[0] || y >= vmax[1] || z >= vmax[2] || x <= vmin[0] || y <= vmin[1] || z <= vmin[2]);
    }
    static uint64_t hash_id(int x, int y, int z) {                // getHashId :143-148
        uint64_t h = (uint64_t)(long long)x;
        return (h << 40) ^ (uint64_t)(long long)(y << 20) ^ (uint64_t)(long long)z;
    }
    size_t integrate(const float* xyz, const float* nrm, size_t n) {   // integratePointCloud :199-228 (:172-197 if nrm == null)
        size_t used = 0;
        for (size_t i = 0; i < n; i++) {
            float px = xyz[3 * i], py = xyz[3 * i + 1], pz = xyz[3 * i + 2];
            if (!valid_point(px, py, pz)) continue;
            int c[3] = {(int)std::floor((px - vmin[0]) / delta[0]), (int)std::floor((py - vmin[1]) / delta[1]), (int)std::floor((pz - vmin[2]) / delta[2])};
            if (c[0] >= dim[0] || c[1] >= dim[1] || c[2] >= dim[2] || c[0] < 0 || c[1] < 0 || c[2] < 0) continue;   // validCoords :211
            uint64_t h = hash_id(c[0], c[1], c[2]);
            auto it = index.find(h);
            uint32_t o;
            if (it == index.end()) { o = (uint32_t)occupied.size(); index.emplace(h, o); occupied.push_back(h); normals.emplace_back(); }
            else o = it->second;
            if (nrm) { normals[o].push_back(nrm[3 * i]); normals[o].push_back(nrm[3 * i + 1]); normals[o].push_back(nrm[3 * i + 2]); }
            used++;
        }
        return used;
    }
};

// degree(acos(d)) in [0,90] with the HOST libm's float acos, exactly as the reference evaluates it
// (CommonUtilities.hpp:17; RayTracingEngine.hpp:211-212): this is a property of the machine the reference would run on.
inline bool host_angle_ok(float d) {
    float a = std::acos(d);
    double deg = ((double)a * 180) / 3.14159;
    if (!(deg > -2147483649.0 && deg < 2147483648.0)) return false;   // int(NaN) = INT_MIN on x86
    int angle = (int)deg;
    return angle >= 0 && angle <= 90;
}
inline float next_up(float f) { return std::nextafter(f, INFINITY); }
inline float next_down(float f) { return std::nextafter(f, -INFINITY); }

// Smallest d with host_angle_ok(d) (bisection over the float order), plus the band where monotonicity fails.
inline AngleTest bisect_angle_test() {
    auto ord = [](float f) { int32_t i; std::memcpy(&i, &f, 4); return i < 0 ? (int32_t)0x80000000 - i : i; };
    auto from = [](int32_t o) { int32_t i = o < 0 ? (int32_t)0x80000000 - o : o; float f; std::memcpy(&f, &i, 4); return f; };
    int32_t lo = ord(-1.0f), hi = ord(1.0f);      // ok(lo) false, ok(hi) true
    while (hi - lo > 1) { int32_t mid = lo + (hi - lo) / 2; if (host_angle_ok(from(mid))) hi = mid; else lo = mid; }
    AngleTest t; t.dot_min = from(hi); t.band_lo = 0.f; t.band_hi = 0.f;
    // scan a neighbourhood for non-monotonic flips
    int32_t first_bad = 0, last_bad = 0; bool any = false;
    for (int32_t o = hi - 4096; o <= hi + 4096; o++) {
        bool expect = o >= hi, got = host_angle_ok(from(o));
        if (expect != got) { if (!any) first_bad = o; last_bad = o; any = true; }
    }
    if (any) { t.band_lo = from(first_bad); t.band_hi = from(last_bad + 1); }
    return t;
}

// A single-view call is a dozen small launches: replaying them as one captured CUDA graph removes the gaps between them.
// A call shape (key) is run directly the first time (it sizes every buffer), captured the second time, replayed afterwards --
// as long as no buffer of the library has moved (alloc_generation) and the key (mode, camera, volume epoch, ...) is the same.
struct CallGraph {
    cudaGraphExec_t exec = nullptr;
    unsigned long long key = 0, gen = 0, cand_key = 0, cand_gen = 0;
    unsigned n_kernels = 0;
    bool disabled = false;
    void drop() { if (exec) cudaGraphExecDestroy(exec); exec = nullptr; key = cand_key = 0; }
};

struct TableKey {
    float K[9]; int H, W, z0, zdelta, cstride, rstride;
    bool operator==(const TableKey& o) const { return std::memcmp(this, &o, sizeof *this) == 0; }
};

}  // namespace dmf

struct dmf_ctx {
    int device = 0;
    cudaStream_t stream = nullptr, copy_stream = nullptr, aux_stream = nullptr;   // aux: every other sub-launch of a split march
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    cudaEvent_t ev_k0 = nullptr, ev_k1 = nullptr;     // bracket the kernels of the last call
    cudaEvent_t ev_compute[2] = {nullptr, nullptr}, ev_copied[2] = {nullptr, nullptr};
    cudaEvent_t ev_h0 = nullptr, ev_h1 = nullptr;     // bracket the dominant march kernel of the last call
    cudaEvent_t ev_p0 = nullptr, ev_p1 = nullptr, ev_b0 = nullptr, ev_b1 = nullptr;   // volume preparation: structures / distance bytes
    bool timed = false, hot_timed = false, prepare_timed = false, bytes_timed = false;
    cudaEvent_t ev_last = nullptr; cudaStream_t last_stream = nullptr; bool last_stream_valid = false;   // the previous *_dev call (stream ordering of shared scratch)
    // camera
    bool cam_set = false; float K[9]; int H = 0, W = 0;
    // volume
    bool vol_set = false;
    VolDev vol{};
    double bounds[6]; double voxel_size = 0; size_t n_occ = 0, n_normals = 0;
    std::vector<uint64_t> h_occ; std::vector<uint32_t> h_noff; std::vector<float> h_normals;   // host mirrors (dmf_volume_get_*)
    bool defer_first_view = false;                    // sharded CLASSIFY: leave first_view for dmf_comm_fuse_marks instead of applying it locally
    bool mirror_valid = false;                        // false: the volume was built / received on the device, mirrors fetched on demand
    dmf::DevBuf d_scan, d_dt_tmp, d_macro_dist[2], d_err;   // scan scratch, distance-transform ping-pong, macro-cell distances, error words
    dmf::DevBuf d_bricks /* bit grid words */, d_macro, d_prefix, d_rank2occ, d_bytes, d_noff, d_normals, d_occ_ids, d_centroid_hash;
    bool bytes_built = false;
    dmf::DevBuf d_rev_perm;                     // occupied ordinals in Morton order of their voxels (work order of k_reverse<FAST>)
    dmf::DevBuf d_tile_rec;                     // k_tile_start: [n_views][tiles] 64-bit records (sample intervals + cone pre-march) read by k_forward_line
    int auto_uses = 0;                    // forward calls with DMF_GRID_AUTO since the volume was uploaded
    int reverse_format = DMF_GRID_BYTE;   // grid the reverse march probes (dmf_set_reverse_format)
    dmf::DevBuf d_view_mark, d_good_bits, d_first_view;
    // carve mode (DMF_FWD_CARVE): observed-voxel bit grid, same layout and word count as the occupancy bit grid; zeroed on first use
    dmf::DevBuf d_observed; size_t n_grid_words = 0; bool observed_ready = false;
    // reverseRayTrace / rayTraceVolume float-accumulated axes
    dmf::DevBuf d_axis[3]; int n_axis[3] = {0, 0, 0};
    // projectPoint tables
    bool tables_valid = false; dmf::TableKey tkey{};
    dmf::DevBuf d_xtab, d_ytab, d_ztab, d_kstart, d_dcx, d_dcy, d_clearance; int S = 0, Wc = 0, Hc = 0; float dcx_max = 0, dcy_max = 0;
    // scratch
    dmf::DevBuf d_poses[2], d_inv_poses, d_out[2][8], d_first_key, d_ray_key, d_ray_occ, d_tmp_a, d_tmp_b, d_out_occ, d_n_ids, d_offsets, d_ids;
    dmf::DevBuf d_misc[4];
    dmf::DevBuf d_counters;
    dmf::HostStage stage;                             // pinned staging of the id-list calls
    dmf::CallGraph graph_fwd_ids, graph_rev_ids;      // captured single-view id-list calls (dmf_forward / dmf_reverse with one pose)
    bool capturing = false;                           // a stream capture is in progress: no timing-event records inside
    unsigned long long volume_epoch = 0;              // bumped by every volume (re)build: part of the graph keys
    AngleTest angle{};
    uint64_t launches = 0;
};
