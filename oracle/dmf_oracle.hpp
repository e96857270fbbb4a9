// dmf_oracle.hpp -- CPU ORACLE for the RayTracingEngine hot path.
//
// TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may build, link, import
// or execute anything in oracle/.  The product (depth-map-fusion-utils_b200/)
// never includes this file and has no CPU fallback.
//
// PINNING STATUS.  The reference (REXJJ/depth-map-fusion-utils) ships no golden
// vectors and no assertions on ray-tracing output, and as a whole cannot be built here
// (Eigen >= 3.3, PCL >= 1.7 ... are absent; CMakeLists.txt:3-29).  This file is a
// from-scratch restatement of the reference's arithmetic, step for step.  Its LOGIC is
// pinned: tests/test_reference_build_cpu.py compares it with the reference's own
// Camera/Volume/RayTracingEngine headers compiled unmodified against a minimal
// Eigen/PCL shim (oracle/ref_capi.cpp -> oracle/_ref/) -- identical on every routine.
// PARITY UNPINNED for one remainder: the float op order INSIDE the un-vendored Eigen
// 3.3, which both this file and the shim write out as rules E1..E5 below.  The
// alternative order is selectable at run time (set_eigen_order) so tests can COUNT
// how many samples change ("tie cases").
//
// What is restated (reference file:line):
//   Camera::projectPoint / deProjectPoint / transformPoints / validPixel
//                                         include/Camera.hpp:24-45,65-68
//   VoxelVolume::setDimensions / setVolumeSize / constructVolume
//                                         include/Volume.hpp:89-128
//   getHash / getHashId / getVoxel / getVoxelCoords / validCoords
//                                         include/Volume.hpp:135-170
//   integratePointCloud(cloud,normals)    include/Volume.hpp:199-228
//   validPoints / getNeighborHashes       include/Volume.hpp:230-255
//   degree()                              include/CommonUtilities.hpp:17
//   RayTracingEngine (8 methods)          include/RayTracingEngine.hpp:45-564
//   Algorithms::greedySetCover            include/Algorithms.hpp:38-86
//
// Eigen 3.3 arithmetic made explicit (SSE2 build, no FMA: CMakeLists.txt:5):
//   E1  Affine3f * Vector3f  ==  Matrix4f * (x,y,z,1), column-accumulated:
//         q_i = ((m_i0*x + m_i1*y) + m_i2*z) + m_i3        (separate mul / add)
//       alternative (order 1):  q_i = m_i3 + (m_i0*x + (m_i1*y + m_i2*z))
//   E2  3-element reductions (dot, squaredNorm, cofactor sums): a0 + (a1 + a2)
//       alternative (order 1):  (a0 + a1) + a2
//   E3  normalized(): n / sqrtf(squaredNorm) component-wise if squaredNorm > 0
//   E4  Vector3f * double  converts the scalar to float; "/ 1000.0" is a true
//       float division:  pt_i = c_i + ((v_i * float(depth)) / 1000.0f)
//   E5  Affine3f::inverse(): 3x3 cofactor inverse times 1/det, translation
//       = -(Linv * t)
//
// Behaviours reproduced on purpose: samples outside the AABB are skipped, not
// terminating (forward); the reverse march does not stop at the camera; the
// forward "centroid" is sample point + delta/2; float-accumulating full-grid
// loops (reverseRayTrace, rayTraceVolume); int(NaN) == INT_MIN (x86 cvttsd2si).
// Reads that the reference would perform out of bounds (UB) are treated as
// "empty" and counted in Counters::oob.
#pragma once
#include <cmath>
#include <climits>
#include <cstdint>
#include <cstring>
#include <tuple>
#include <unordered_set>
#include <utility>
#include <vector>
#include <algorithm>
#include <iterator>

namespace dmf_oracle {

using u64 = unsigned long long;

// ---- run-time switch for the unpinned Eigen op order (see header) -----------
inline int& eigen_order() { static int o = 0; return o; }
inline void set_eigen_order(int o) { eigen_order() = o; }

inline float sum3(float a0, float a1, float a2) {                      // E2
    return eigen_order() == 0 ? a0 + (a1 + a2) : (a0 + a1) + a2;
}

struct Vec3 {
    float v[3];
    float& operator()(int i) { return v[i]; }
    float operator()(int i) const { return v[i]; }
};
inline Vec3 operator-(const Vec3& a, const Vec3& b) { return {{a.v[0]-b.v[0], a.v[1]-b.v[1], a.v[2]-b.v[2]}}; }
inline float dot(const Vec3& a, const Vec3& b) { return sum3(a.v[0]*b.v[0], a.v[1]*b.v[1], a.v[2]*b.v[2]); }
inline Vec3 normalized(const Vec3& n) {                                // E3
    float z = sum3(n.v[0]*n.v[0], n.v[1]*n.v[1], n.v[2]*n.v[2]);
    if (z > 0.0f) { float s = std::sqrt(z); return {{n.v[0]/s, n.v[1]/s, n.v[2]/s}}; }
    return n;
}

// 3x4 affine (last row 0 0 0 1 implied), element (r,c) like Eigen::Affine3f.
struct Affine {
    float m[3][4];
    float& operator()(int r, int c) { return m[r][c]; }
    float operator()(int r, int c) const { return m[r][c]; }
    static Affine Identity() { Affine a; std::memset(a.m, 0, sizeof a.m); a.m[0][0]=a.m[1][1]=a.m[2][2]=1.0f; return a; }
    Vec3 apply(float x, float y, float z) const {                      // E1
        Vec3 q;
        for (int i = 0; i < 3; i++) {
            if (eigen_order() == 0) q.v[i] = ((m[i][0]*x + m[i][1]*y) + m[i][2]*z) + m[i][3];
            else                    q.v[i] = m[i][3] + (m[i][0]*x + (m[i][1]*y + m[i][2]*z));
        }
        return q;
    }
    Affine inverse() const {                                           // E5
        auto cof = [&](int i, int j) {
            int i1=(i+1)%3, i2=(i+2)%3, j1=(j+1)%3, j2=(j+2)%3;
            return m[i1][j1]*m[i2][j2] - m[i1][j2]*m[i2][j1];
        };
        float c0 = cof(0,0), c1 = cof(1,0), c2 = cof(2,0);
        float det = sum3(c0*m[0][0], c1*m[1][0], c2*m[2][0]);
        float invdet = 1.0f / det;
        Affine r;
        r.m[0][0] = c0*invdet; r.m[0][1] = c1*invdet; r.m[0][2] = c2*invdet;
        r.m[1][0] = cof(0,1)*invdet; r.m[1][1] = cof(1,1)*invdet; r.m[1][2] = cof(2,1)*invdet;
        r.m[2][0] = cof(0,2)*invdet; r.m[2][1] = cof(1,2)*invdet; r.m[2][2] = cof(2,2)*invdet;
        for (int i = 0; i < 3; i++)
            r.m[i][3] = sum3((-r.m[i][0])*m[0][3], (-r.m[i][1])*m[1][3], (-r.m[i][2])*m[2][3]);
        return r;
    }
};

// x86 cvttsd2si semantics for the int(double) casts the reference performs on
// possibly NaN / out-of-range values (UB in ISO C++, INT_MIN on the author's CPU).
inline int to_int_x86(double v) {
    if (!(v > -2147483649.0 && v < 2147483648.0)) return INT_MIN;
    return (int)v;
}

// CommonUtilities.hpp:17
inline int degree(double radian) { return to_int_x86((radian * 180) / 3.14159); }

constexpr double k_AngleMin = 0, k_AngleMax = 90, k_ZMin = 0.20, k_ZMax = 1.0;   // RayTracingEngine.hpp:22-25

// ---- Camera (Camera.hpp:17-86) -----------------------------------------------
class Camera {
    std::vector<float> K_;
    int height_ = 480, width_ = 640;
public:
    Camera() {}
    Camera(const std::vector<float>& K, int height = 480, int width = 640) : K_(K), height_(height), width_(width) {}
    std::tuple<float,float,float> projectPoint(int r, int c, int depth_mm) const {      // :24-31
        double fx = K_[0], cx = K_[2], fy = K_[4], cy = K_[5];
        double z = depth_mm * 0.001;
        double x = z * ((double)c - cx) / fx;
        double y = z * ((double)r - cy) / fy;
        return std::make_tuple((float)x, (float)y, (float)z);
    }
    std::tuple<int,int> deProjectPoint(double x, double y, double z) const {            // :32-38
        double fx = K_[0], cx = K_[2], fy = K_[4], cy = K_[5];
        int c = to_int_x86(std::round((x * fx) / z + cx));
        int r = to_int_x86(std::round((y * fy) / z + cy));
        return std::make_tuple(r, c);
    }
    std::tuple<float,float,float> transformPoints(double x, double y, double z, const Affine& T) const {   // :39-45
        Vec3 p = T.apply((float)x, (float)y, (float)z);
        return std::make_tuple(p.v[0], p.v[1], p.v[2]);
    }
    int getHeight() const { return height_; }
    int getWidth() const { return width_; }
    bool validPixel(int r, int c) const { return r >= 0 && r < height_ && c >= 0 && c < width_; }
};

// ---- Voxel / VoxelVolume (Volume.hpp:29-255) ---------------------------------
struct Voxel {
    int n_pts = 0;
    std::vector<Vec3> normals;
    int view = 0;
    bool good = false;
};

// Two storages behind one accessor: the reference's vector<vector<vector<Voxel*>>>
// (used for the timed CPU baseline) and a flat int32 index (fast golden generation).
// Tests assert both produce identical results.
class VoxelVolume {
public:
    std::vector<u64> occupied_cells_;
    double xmin_=0,xmax_=0,ymin_=0,ymax_=0,zmin_=0,zmax_=0;
    double xcenter_=0,ycenter_=0,zcenter_=0;
    double xdelta_=0,ydelta_=0,zdelta_=0;
    double voxel_size_=0;
    int xdim_=0,ydim_=0,zdim_=0;
    bool flat_ = false;
    std::vector<std::vector<std::vector<Voxel*>>> voxels_;   // pointer grid (faithful)
    std::vector<int32_t> cell_;                               // flat: index into pool_, -1 = empty
    std::vector<Voxel*> pool_;                                // owned voxels in first-insertion order
    long long oob_ = 0;

    explicit VoxelVolume(bool flat = false) : flat_(flat) {}
    ~VoxelVolume() { for (auto* p : pool_) delete p; }
    VoxelVolume(const VoxelVolume&) = delete;
    VoxelVolume& operator=(const VoxelVolume&) = delete;

    void setDimensions(double xmin,double xmax,double ymin,double ymax,double zmin,double zmax) {   // :89-100
        xmin_=xmin; xmax_=xmax; ymin_=ymin; ymax_=ymax; zmin_=zmin; zmax_=zmax;
        xcenter_ = xmin_+(xmax_-xmin_)/2.0; ycenter_ = ymin_+(ymax_-ymin_)/2.0; zcenter_ = zmin_+(zmax_-zmin_)/2.0;
    }
    void setVolumeSize(int xdim,int ydim,int zdim) {                                                 // :109-117
        xdim_=xdim; ydim_=ydim; zdim_=zdim;
        xdelta_=(xmax_-xmin_)/xdim; ydelta_=(ymax_-ymin_)/ydim; zdelta_=(zmax_-zmin_)/zdim;
    }
    bool constructVolume() {                                                                         // :119-128
        xdim_ = (int)((xmax_-xmin_)/xdelta_);       // truncation: may give dim-1 for non-dyadic sizes
        ydim_ = (int)((ymax_-ymin_)/ydelta_);
        zdim_ = (int)((zmax_-zmin_)/zdelta_);
        voxel_size_ = xdelta_*ydelta_*zdelta_;
        if (flat_) cell_.assign((size_t)xdim_*ydim_*zdim_, -1);
        else voxels_ = std::vector<std::vector<std::vector<Voxel*>>>(xdim_, std::vector<std::vector<Voxel*>>(ydim_, std::vector<Voxel*>(zdim_, nullptr)));
        return true;
    }
    // voxels_[x][y][z] of the reference.  Out-of-range reads are UB there; here: empty + counted.
    inline Voxel* at(int x, int y, int z) {
        if ((unsigned)x >= (unsigned)xdim_ || (unsigned)y >= (unsigned)ydim_ || (unsigned)z >= (unsigned)zdim_) { oob_++; return nullptr; }
        if (flat_) { int32_t i = cell_[((size_t)x*ydim_ + y)*zdim_ + z]; return i < 0 ? nullptr : pool_[i]; }
        return voxels_[x][y][z];
    }
    inline void put(int x, int y, int z, Voxel* v) {
        if (flat_) cell_[((size_t)x*ydim_ + y)*zdim_ + z] = (int32_t)(pool_.size() - 1);
        else voxels_[x][y][z] = v;
    }
    inline std::tuple<int,int,int> getVoxel(float x, float y, float z) const {                       // :150-156
        int xv = (int)std::floor((x-xmin_)/xdelta_);
        int yv = (int)std::floor((y-ymin_)/ydelta_);
        int zv = (int)std::floor((z-zmin_)/zdelta_);
        return std::make_tuple(xv,yv,zv);
    }
    inline u64 getHashId(int x, int y, int z) const {                                                // :143-148
        u64 hash = x;
        hash = (hash<<40)^(u64)(long long)(y<<20)^(u64)(long long)z;   // y<<20 and z are ints, sign-extended by the xor
        return hash;
    }
    inline u64 getHash(float x, float y, float z) const {                                            // :135-141
        auto c = getVoxel(x,y,z);
        u64 hash = std::get<0>(c);
        hash = hash<<40^(u64)(long long)(std::get<1>(c)<<20)^(u64)(long long)(std::get<2>(c));
        return hash;
    }
    inline std::tuple<int,int,int> getVoxelCoords(u64 id) const {                                    // :158-165
        constexpr u64 mask = (1<<20)-1;
        return std::make_tuple((int)(id>>40), (int)(id>>20&mask), (int)(id&mask));
    }
    inline bool validCoords(int xid,int yid,int zid) const {                                         // :167-170
        return xid<xdim_&&yid<ydim_&&zid<zdim_&&xid>=0&&yid>=0&&zid>=0;
    }
    inline bool validPoints(float x,float y,float z) const {                                         // :230-233
        return !(x>=xmax_||y>=ymax_||z>=zmax_||x<=xmin_||y<=ymin_||z<=zmin_);
    }
    // integratePointCloud(cloud, normals) (:199-228).  normals may be null => the
    // xyz-only overload (:172-197) which has no validCoords guard.
    long integrate(const float* xyz, const float* nrm, long n) {
        long used = 0;
        for (long i = 0; i < n; i++) {
            float px = xyz[3*i], py = xyz[3*i+1], pz = xyz[3*i+2];
            if (!validPoints(px,py,pz)) continue;
            int x,y,z; std::tie(x,y,z) = getVoxel(px,py,pz);
            if (!validCoords(x,y,z)) { if (!nrm) oob_++; continue; }
            Voxel* v = at(x,y,z);
            if (v == nullptr) {
                occupied_cells_.push_back(getHashId(x,y,z));
                v = new Voxel();
                pool_.push_back(v);
                put(x,y,z,v);
                v->n_pts = 1;
                if (nrm) v->normals.push_back({{nrm[3*i],nrm[3*i+1],nrm[3*i+2]}});
            } else {
                v->n_pts++;
                if (nrm) v->normals.push_back({{nrm[3*i],nrm[3*i+1],nrm[3*i+2]}});
            }
            used++;
        }
        return used;
    }
    std::vector<u64> getNeighborHashes(u64 hash, int K = 1) {                                        // :235-255
        double x,y,z; { int a,b,c; std::tie(a,b,c) = getVoxelCoords(hash); x=a; y=b; z=c; }
        std::vector<u64> neighbors;
        for (int i=-K;i<=K;i++) for (int j=-K;j<=K;j++) for (int k=-K;k<=K;k++) {
            if (((i==j)==k)==0) continue;                // the reference's `i==j==k==0`
            if (validCoords((int)(x+i),(int)(y+j),(int)(z+k)))
                if (at((int)(x+i),(int)(y+j),(int)(z+k)) != nullptr)
                    neighbors.push_back(getHashId((int)(x+i),(int)(y+j),(int)(z+k)));
        }
        return neighbors;
    }
};

// ---- instrumentation (not in the reference; null pointers disable it) ---------
struct Counters {
    long long samples = 0;        // (pixel, z_depth) or (voxel, step) probes evaluated
    long long inbounds = 0;       // probes that passed validPoints (forward: up to and incl. the first hit)
    long long hits = 0;           // forward: rays with a hit;  reverse: visible voxels
    long long oob = 0;            // reads the reference would do out of bounds
    long long runaway = 0;        // reverse marches stopped by the safety cap (infinite loop in the reference)
};
struct PixelOut {                  // optional per-pixel outputs (size H*W each), all may be null
    int32_t* depth = nullptr;      // first-hit z_depth in mm, -1 = no hit / not cast
    float* point = nullptr;        // world-space sample point of the first hit (x,y,z), untouched if none
    u64* voxel = nullptr;          // voxel id of the first hit
    // Carve mode -- an EXTENSION, not in the reference (which never records free space); this restatement is its only
    // definition: every sample that passes validPoints and is visited by its ray up to and including the first hit (the
    // samples `Counters::inbounds` counts) sets the bit of the voxel getVoxel (Volume.hpp:150-156) puts it in.  uint32 words
    // over the padded index space [0,xdim] x [0,ydim] x [0,zdim], bit index = (x*(ydim+1) + y)*(zdim+1) + z; OR-accumulated.
    uint32_t* observed = nullptr;
};

using IdList = std::pair<bool, std::vector<u64>>;

class RayTracingEngine {
public:
    Camera cam_;
    bool dead_neighbor_work_ = false;   // execute the dead getNeighborHashes(...,5) of reverseRayTraceFast (:170-171)
    int reverse_step_cap_ = 1000000;
    explicit RayTracingEngine(const Camera& cam) : cam_(cam) {}

private:
    // One forward probe: pixel (r,c) at z_depth -> world point (floats) -> voxel.  Returns false if outside the AABB.
    inline bool probe(VoxelVolume& vol, const Affine& T, int r, int c, int z_depth,
                      double& x, double& y, double& z, int& xid, int& yid, int& zid) const {
        std::tie(x,y,z) = cam_.projectPoint(r,c,z_depth);
        std::tie(x,y,z) = cam_.transformPoints(x,y,z,T);
        if (vol.validPoints(x,y,z) == false) return false;
        std::tie(xid,yid,zid) = vol.getVoxel(x,y,z);
        return true;
    }
    // "good" predicate of the forward routines (:358-369, :424-438): some stored normal within [0,90] deg of v,
    // inside the 250..600 mm window.
    inline bool forward_good(const Voxel* voxel, const Vec3& v, int z_depth) const {
        for (const auto& n : voxel->normals) {
            int angle_z = degree(std::acos(dot(n, v)));      // float acos overload
            if (z_depth >= 250 && z_depth <= 600)
                if (angle_z >= k_AngleMin && angle_z <= k_AngleMax) return true;
        }
        return false;
    }
    inline Vec3 view_dir(const VoxelVolume& vol, const Affine& T, double x, double y, double z) const {   // :345-349
        Vec3 centroid = {{(float)(x+vol.xdelta_/2.0), (float)(y+vol.ydelta_/2.0), (float)(z+vol.zdelta_/2.0)}};
        Vec3 camera_center = {{T(0,3), T(1,3), T(2,3)}};
        return normalized(camera_center - centroid);
    }

public:
    enum Mode { POINTS = 0, GOOD_POINTS = 1, CLASSIFY = 2, MARK = 3, MINIMUM = 4 };

    // Shared body of the five forward routines.  Loop nest, skip rules and emission rules follow
    //   rayTraceAndGetPoints :447-494, rayTraceAndGetGoodPoints :377-445, rayTraceAndClassify :311-375,
    //   rayTrace :268-309, rayTraceAndGetMinimum :229-264.
    // z_depth outermost, pixels inner; `found` per pixel; ids in discovery order, first occurrence only.
    IdList forward(Mode mode, VoxelVolume& vol, const Affine& T, int zdelta, bool sparse, int view,
                   int* min_depth, Counters* cnt, const PixelOut* po) const {
        const int width = cam_.getWidth(), height = cam_.getHeight();
        std::vector<char> found((size_t)height*width, 0);
        int rdelta = 1, cdelta = 1;
        if (sparse) { rdelta = cdelta = (mode == MINIMUM ? 10 : 5); }          // :236-237 vs :277-278
        const int z0 = (mode == MINIMUM ? 5 : 10);                               // :239 vs :280
        bool point_found = false;
        std::vector<u64> out;
        std::unordered_set<u64> checked;
        if (min_depth) *min_depth = -1;
        if (po && po->depth) std::fill(po->depth, po->depth + (size_t)height*width, -1);
        for (int z_depth = z0; z_depth < k_ZMax*1000; z_depth += zdelta) {
            for (int r = 0; r < height; r += rdelta) {
                for (int c = 0; c < width; c += cdelta) {
                    if (mode != MINIMUM && found[(size_t)r*width+c]) continue;
                    if (cnt) cnt->samples++;
                    double x,y,z; int xid,yid,zid;
                    if (!probe(vol,T,r,c,z_depth,x,y,z,xid,yid,zid)) continue;
                    if (cnt) cnt->inbounds++;
                    if (po && po->observed && mode != MINIMUM) {
                        size_t idx = ((size_t)xid*(size_t)(vol.ydim_+1) + (size_t)yid)*(size_t)(vol.zdim_+1) + (size_t)zid;
                        po->observed[idx >> 5] |= 1u << (idx & 31);
                    }
                    Voxel* voxel = vol.at(xid,yid,zid);
                    if (voxel == nullptr) continue;
                    if (mode == MINIMUM) {                                   // :256-259
                        if (min_depth) *min_depth = z_depth;
                        if (cnt) cnt->hits++;
                        return std::make_pair(true, out);
                    }
                    found[(size_t)r*width+c] = 1;
                    point_found = true;
                    if (cnt) cnt->hits++;
                    u64 id = vol.getHashId(xid,yid,zid);
                    if (po) {
                        size_t p = (size_t)r*width+c;
                        if (po->depth) po->depth[p] = z_depth;
                        if (po->point) { po->point[3*p]=(float)x; po->point[3*p+1]=(float)y; po->point[3*p+2]=(float)z; }
                        if (po->voxel) po->voxel[p] = id;
                    }
                    switch (mode) {
                    case POINTS:                                             // :484-488
                        if (checked.find(id) == checked.end()) { checked.insert(id); out.push_back(id); }
                        break;
                    case MARK:                                               // :302
                        voxel->view = 1;
                        break;
                    case GOOD_POINTS:                                        // :424-439
                        if (forward_good(voxel, view_dir(vol,T,x,y,z), z_depth))
                            if (checked.find(id) == checked.end()) { checked.insert(id); out.push_back(id); }
                        break;
                    case CLASSIFY:                                           // :354-370
                        if (voxel->view == 0) voxel->view = view;
                        if (voxel->good == false)
                            if (forward_good(voxel, view_dir(vol,T,x,y,z), z_depth)) voxel->good = true;
                        break;
                    default: break;
                    }
                }
            }
        }
        if (cnt) cnt->oob = vol.oob_;
        return std::make_pair(point_found, out);
    }

    IdList rayTraceAndGetPoints(VoxelVolume& v, const Affine& T, int zdelta = 10, bool sparse = true) const { return forward(POINTS, v, T, zdelta, sparse, 1, nullptr, nullptr, nullptr); }
    IdList rayTraceAndGetGoodPoints(VoxelVolume& v, const Affine& T, int zdelta = 10, bool sparse = true) const { return forward(GOOD_POINTS, v, T, zdelta, sparse, 1, nullptr, nullptr, nullptr); }
    void rayTraceAndClassify(VoxelVolume& v, const Affine& T, int zdelta = 10, int view = 1, bool sparse = true) const { forward(CLASSIFY, v, T, zdelta, sparse, view, nullptr, nullptr, nullptr); }
    void rayTrace(VoxelVolume& v, const Affine& T, int zdelta = 10, bool sparse = true) const { forward(MARK, v, T, zdelta, sparse, 1, nullptr, nullptr, nullptr); }
    int rayTraceAndGetMinimum(VoxelVolume& v, const Affine& T, int zdelta = 1, bool sparse = true) const { int d; forward(MINIMUM, v, T, zdelta, sparse, 1, &d, nullptr, nullptr); return d; }

private:
    // The 1 mm march of the reverse routines (:81-103, :172-200): from `centroid` along v until the AABB is left.
    inline bool reverse_collides(VoxelVolume& vol, const Vec3& centroid, const Vec3& v, u64 centroid_hash,
                                 int depth0, Counters* cnt) const {
        for (int depth = depth0; ; depth++) {
            if (depth - depth0 > reverse_step_cap_) { if (cnt) cnt->runaway++; return false; }
            float s = (float)(double)depth;                                 // E4
            float px = centroid.v[0] + (v.v[0]*s)/1000.0f;
            float py = centroid.v[1] + (v.v[1]*s)/1000.0f;
            float pz = centroid.v[2] + (v.v[2]*s)/1000.0f;
            double xx = px, yy = py, zz = pz;
            if (cnt) cnt->samples++;
            if (vol.validPoints(xx,yy,zz) == false) break;
            if (cnt) cnt->inbounds++;
            u64 hash = vol.getHash(xx,yy,zz);
            if (hash == centroid_hash) continue;
            int xidn,yidn,zidn; std::tie(xidn,yidn,zidn) = vol.getVoxel(xx,yy,zz);
            if (vol.validCoords(xidn,yidn,zidn) == false) break;
            if (vol.at(xidn,yidn,zidn) != nullptr) return true;
        }
        return false;
    }

public:
    // reverseRayTraceFast :136-226 -- what every shipped driver calls.
    IdList reverseRayTraceFast(VoxelVolume& vol, const Affine& T, bool viz, int /*zdelta*/ = 1, Counters* cnt = nullptr,
                               std::vector<char>* visible_flags = nullptr) const {
        Affine inverseT = T.inverse();
        bool found = false;
        std::vector<u64> good_points;
        if (visible_flags) visible_flags->assign(vol.occupied_cells_.size(), 0);
        size_t ordinal = 0;
        for (auto hashes : vol.occupied_cells_) {
            size_t me = ordinal++;
            int xid,yid,zid; std::tie(xid,yid,zid) = vol.getVoxelCoords(hashes);
            float x = (float)(xid*vol.xdelta_ + vol.xmin_);
            float y = (float)(yid*vol.ydelta_ + vol.ymin_);
            float z = (float)(zid*vol.zdelta_ + vol.zmin_);
            Voxel* voxel = vol.at(xid,yid,zid);
            Vec3 centroid = {{(float)(x+vol.xdelta_/2.0), (float)(y+vol.ydelta_/2.0), (float)(z+vol.zdelta_/2.0)}};
            float xxx,yyy,zzz;
            std::tie(xxx,yyy,zzz) = cam_.transformPoints(x+vol.xdelta_/2.0, y+vol.ydelta_/2.0, z+vol.zdelta_/2.0, inverseT);
            int r,c; std::tie(r,c) = cam_.deProjectPoint(xxx,yyy,zzz);
            u64 centroid_hash = vol.getHash(centroid(0),centroid(1),centroid(2));
            if (cam_.validPixel(r,c) == false) continue;
            Vec3 camera_center = {{T(0,3),T(1,3),T(2,3)}};
            Vec3 v = normalized(camera_center - centroid);
            if (dead_neighbor_work_) {
                auto neighbors = vol.getNeighborHashes(vol.getHash(x,y,z),5);
                std::unordered_set<u64> n_set(neighbors.begin(),neighbors.end());
                asm volatile("" :: "r"(n_set.size()) : "memory");
            }
            bool collided = reverse_collides(vol, centroid, v, centroid_hash, 50, cnt);
            if (collided == false) {
                found = true;
                if (cnt) cnt->hits++;
                if (visible_flags) (*visible_flags)[me] |= 1;
                if (viz) voxel->view = 1;
                if (zzz >= k_ZMin && zzz <= k_ZMax) {
                    for (const auto& n : voxel->normals) {
                        int angle_z = degree(std::acos(dot(n, v)));
                        if (angle_z >= k_AngleMin && angle_z <= k_AngleMax) {
                            if (viz) voxel->good = true;
                            good_points.push_back(centroid_hash);
                            if (visible_flags) (*visible_flags)[me] |= 2;
                            break;
                        }
                    }
                }
            }
        }
        if (cnt) cnt->oob = vol.oob_;
        return std::make_pair(found, good_points);
    }

    // Positions visited by `for(float x=min; x<max; x+=delta)` with a double delta (:54-56, :509-511).
    static std::vector<float> float_axis(double lo, double hi, double delta) {
        std::vector<float> a;
        for (float x = (float)lo; x < hi; x = (float)(x + delta)) { a.push_back(x); if (a.size() > (1u<<22)) break; }
        return a;
    }

    // reverseRayTrace :45-134 -- whole-grid scan, march from depth=1, "good" = depth window only.
    IdList reverseRayTrace(VoxelVolume& vol, const Affine& T, bool viz, int /*zdelta*/ = 1, Counters* cnt = nullptr) const {
        Affine inverseT = T.inverse();
        bool found = false;
        std::vector<u64> good_points;
        auto xs = float_axis(vol.xmin_, vol.xmax_, vol.xdelta_);
        auto ys = float_axis(vol.ymin_, vol.ymax_, vol.ydelta_);
        auto zs = float_axis(vol.zmin_, vol.zmax_, vol.zdelta_);
        for (float x : xs) for (float y : ys) for (float z : zs) {
            int xid,yid,zid; std::tie(xid,yid,zid) = vol.getVoxel(x,y,z);
            Voxel* voxel = vol.at(xid,yid,zid);
            if (voxel == nullptr) continue;
            Vec3 centroid = {{(float)(x+vol.xdelta_/2.0), (float)(y+vol.ydelta_/2.0), (float)(z+vol.zdelta_/2.0)}};
            float xx,yy,zz;
            std::tie(xx,yy,zz) = cam_.transformPoints(x+vol.xdelta_/2.0, y+vol.ydelta_/2.0, z+vol.zdelta_/2.0, inverseT);
            int r,c; std::tie(r,c) = cam_.deProjectPoint(xx,yy,zz);
            u64 centroid_hash = vol.getHash(centroid(0),centroid(1),centroid(2));
            if (cam_.validPixel(r,c) == false) continue;
            Vec3 camera_center = {{T(0,3),T(1,3),T(2,3)}};
            Vec3 v = normalized(camera_center - centroid);
            bool collided = reverse_collides(vol, centroid, v, centroid_hash, 1, cnt);
            if (collided == false) {
                found = true;
                if (cnt) cnt->hits++;
                if (viz) voxel->view = 1;
                if (zz >= k_ZMin && zz <= k_ZMax) {
                    if (viz) voxel->good = true;
                    good_points.push_back(centroid_hash);
                }
            }
        }
        if (cnt) cnt->oob = vol.oob_;
        return std::make_pair(found, good_points);
    }

    // rayTraceVolume :498-564 -- z-buffer splat, then mark voxels whose depth equals the buffer.
    // depth_out (H*W) receives the buffer (not returned by the reference; exposed for parity).
    void rayTraceVolume(VoxelVolume& vol, const Affine& T, int32_t* depth_out = nullptr, long long* counter_out = nullptr) const {
        const int width = cam_.getWidth(), height = cam_.getHeight();
        std::vector<int> depth((size_t)height*width, -1);
        Affine inv = T.inverse();
        long long counter = 0;
        auto xs = float_axis(vol.xmin_, vol.xmax_, vol.xdelta_);
        auto ys = float_axis(vol.ymin_, vol.ymax_, vol.ydelta_);
        auto zs = float_axis(vol.zmin_, vol.zmax_, vol.zdelta_);
        for (int pass = 0; pass < 2; pass++)
            for (float x : xs) for (float y : ys) for (float z : zs) {
                int xid,yid,zid; std::tie(xid,yid,zid) = vol.getVoxel(x,y,z);
                Voxel* voxel = vol.at(xid,yid,zid);
                if (voxel == nullptr) continue;
                float xx,yy,zz;
                std::tie(xx,yy,zz) = cam_.transformPoints(x+vol.xdelta_/2.0, y+vol.ydelta_/2.0, z+vol.zdelta_/2.0, inv);
                int r,c; std::tie(r,c) = cam_.deProjectPoint(xx,yy,zz);
                if (cam_.validPixel(r,c) == false) continue;
                int d = to_int_x86(std::round(zz*1000));          // float multiply, float round (:528)
                int& slot = depth[(size_t)r*width+c];
                if (pass == 0) { slot = (slot == -1) ? d : std::min(slot, d); counter++; }
                else if (slot == d) voxel->view = 1;
            }
        if (depth_out) std::copy(depth.begin(), depth.end(), depth_out);
        if (counter_out) *counter_out = counter;
    }
};

// willCollide (tests/CameraPathGen.cpp:128-156; variants without the validCoords guard at
// tests/CameraMotionTSP.cpp:236-261 and tests/CameraMotionPlanner.cpp:246-271): 1 mm march from a towards b.
inline bool willCollide(VoxelVolume& volume, const Vec3& a, const Vec3& b, bool guard_coords, long long* steps = nullptr) {
    Vec3 ab = a - b;
    double distance = std::sqrt(sum3(ab.v[0]*ab.v[0], ab.v[1]*ab.v[1], ab.v[2]*ab.v[2]));   // Vector3f::norm(): float sqrt, widened
    Vec3 v = normalized(b - a);
    bool collided = false;
    for (int depth = 1; collided == false; depth++) {
        float s = (float)(double)depth;                                            // rule E4
        float px = a.v[0] + (v.v[0]*s)/1000.0f, py = a.v[1] + (v.v[1]*s)/1000.0f, pz = a.v[2] + (v.v[2]*s)/1000.0f;
        double xx = px, yy = py, zz = pz;
        if (depth > distance*1000) break;
        if (steps) (*steps)++;
        if (volume.validPoints(xx,yy,zz) == false) continue;
        int xidn,yidn,zidn; std::tie(xidn,yidn,zidn) = volume.getVoxel(xx,yy,zz);
        if (guard_coords && volume.validCoords(xidn,yidn,zidn) == false) continue;
        if (volume.at(xidn,yidn,zidn) != nullptr) collided = true;
    }
    return collided;
}

// Algorithms::moveCamera / repositionCamera (Algorithms.hpp:170-188): translation -= row2(linear) * distance / 1000, in float
inline Affine repositionCamera(Affine camera, unsigned distance) {
    float d = (float)(double)distance;
    for (int i = 0; i < 3; i++) camera.m[i][3] = camera.m[i][3] - (camera.m[2][i]*d)/1000.0f;
    return camera;
}

// Algorithms::optimizeCameraPosition(volume, engine, res, Affine3f camera) (Algorithms.hpp:394-421): binary search of the
// stand-off in [300,600] mm on the number of ids reverseRayTrace returns.
inline Affine optimizeCameraPosition(VoxelVolume& volume, const RayTracingEngine& engine, const Affine& camera, unsigned* mid_out = nullptr) {
    unsigned low = 300, high = 600, mid = low;
    while (low < high) {
        auto lo_ids = engine.reverseRayTrace(volume, repositionCamera(camera, low), false).second;
        auto hi_ids = engine.reverseRayTrace(volume, repositionCamera(camera, high), false).second;
        mid = (low + high) / 2;
        if (hi_ids.size() > lo_ids.size()) low = mid + 1; else high = mid;
    }
    if (mid_out) *mid_out = mid;
    return repositionCamera(camera, mid);
}

// Algorithms::greedySetCover (Algorithms.hpp:38-86).  candidate_sets must be sorted.
inline std::vector<u64> greedySetCover(const std::vector<std::vector<u64>>& candidate_sets) {
    std::vector<u64> covered, selected_sets, set_ids(candidate_sets.size());
    for (size_t i = 0; i < set_ids.size(); i++) set_ids[i] = i;
    while (true) {
        u64 selected = (u64)-1;
        u64 max_points = 0;
        for (auto x : set_ids) {
            std::vector<u64> difference;
            std::set_difference(candidate_sets[x].begin(),candidate_sets[x].end(),covered.begin(),covered.end(),std::inserter(difference,difference.begin()));
            if (difference.size() > max_points) { max_points = difference.size(); selected = x; }
        }
        if (selected == (u64)-1) break;
        if (max_points < 5) break;
        std::vector<u64> difference;
        std::set_difference(candidate_sets[selected].begin(),candidate_sets[selected].end(),covered.begin(),covered.end(),std::inserter(difference,difference.begin()));
        for (auto x : difference) covered.push_back(x);
        std::sort(covered.begin(),covered.end());
        selected_sets.push_back(selected);
        set_ids.erase(std::remove(set_ids.begin(),set_ids.end(),selected),set_ids.end());
    }
    return selected_sets;
}

}  // namespace dmf_oracle
