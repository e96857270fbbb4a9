// Volume.hpp -- drop-in for the reference's include/Volume.hpp (struct Voxel :29-48, class VoxelVolume :50-255).
//
// The HOST data model is kept exactly as drivers expect it -- voxels_[x][y][z] raw pointers (nullptr = empty),
// occupied_cells_ in first-insertion order, Voxel::view / Voxel::good, all public bounds/delta members -- because
// drivers and other headers read them directly (tests/CameraPathGen.cpp:150, include/Algorithms.hpp:323,
// include/VisualizationUtilities.hpp:197-405).  What changes: RayTracingEngine no longer walks this structure; it
// uploads a compact device mirror (occupied ids + per-voxel normal lists) through dmf_upload_volume whenever
// revision() changed, and writes view/good back after calls that mutate them.
#pragma once
#include <atomic>
#include <cmath>
#include <cstdint>
#include <tuple>
#include <vector>
#include <Eigen/Dense>
#include <Eigen/Core>
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#include <pcl/impl/point_types.hpp>

using namespace std;
using namespace pcl;

struct Voxel
{
    vector<pcl::PointXYZRGB> pts;
    vector<pcl::Normal> normals;
    int view;
    bool good;
    Voxel(pcl::PointXYZRGB pt) : view(0), good(false) { pts.push_back(pt); }
    Voxel(pcl::PointXYZRGB pt, pcl::Normal normal) : view(0), good(false) { pts.push_back(pt); normals.push_back(normal); }
};

// Revisions come from one process-wide counter, so no two states of any two volumes ever share one: a new VoxelVolume that
// happens to live at the address of a destroyed one (stack local in a loop, heap reuse) can never be mistaken for it by the
// engine's (&volume, revision) cache.
inline unsigned long long dmf_next_volume_revision()
{
    static std::atomic<unsigned long long> counter{0};
    return ++counter;
}

class VoxelVolume
{
    unsigned long long revision_ = dmf_next_volume_revision();   // renewed on every change of occupancy / normals / geometry
    public:
    vector<unsigned long long int> occupied_cells_;
    double xmin_, xmax_, ymin_, ymax_, zmin_, zmax_;
    double xcenter_, ycenter_, zcenter_;
    double xdelta_, ydelta_, zdelta_;
    double voxel_size_;
    int xdim_ = 0, ydim_ = 0, zdim_ = 0;
    unsigned long long int hsize_;
    vector<vector<vector<Voxel*>>> voxels_;

    VoxelVolume() {}
    ~VoxelVolume()
    {
        for (auto& plane : voxels_) for (auto& row : plane) for (Voxel* v : row) delete v;
    }
    VoxelVolume(const VoxelVolume&) = delete;             // the reference would double-free on copy
    VoxelVolume& operator=(const VoxelVolume&) = delete;

    void setDimensions(double xmin, double xmax, double ymin, double ymax, double zmin, double zmax)
    {
        xmin_ = xmin; xmax_ = xmax; ymin_ = ymin; ymax_ = ymax; zmin_ = zmin; zmax_ = zmax;
        xcenter_ = xmin_ + (xmax_ - xmin_) / 2.0;
        ycenter_ = ymin_ + (ymax_ - ymin_) / 2.0;
        zcenter_ = zmin_ + (zmax_ - zmin_) / 2.0;
        revision_ = dmf_next_volume_revision();
    }
    void setResolution(double xdelta, double ydelta, double zdelta) { xdelta_ = xdelta; ydelta_ = ydelta; zdelta_ = zdelta; revision_ = dmf_next_volume_revision(); }
    void setVolumeSize(int xdim, int ydim, int zdim)
    {
        xdim_ = xdim; ydim_ = ydim; zdim_ = zdim;
        xdelta_ = (xmax_ - xmin_) / xdim; ydelta_ = (ymax_ - ymin_) / ydim; zdelta_ = (zmax_ - zmin_) / zdim;
        revision_ = dmf_next_volume_revision();
    }
    bool constructVolume()
    {
        // the reference re-derives the dims from the deltas by truncation (:121-123); keep that
        xdim_ = (xmax_ - xmin_) / xdelta_;
        ydim_ = (ymax_ - ymin_) / ydelta_;
        zdim_ = (zmax_ - zmin_) / zdelta_;
        hsize_ = xdim_ * ydim_ * zdim_;
        voxel_size_ = xdelta_ * ydelta_ * zdelta_;
        voxels_.assign(xdim_, vector<vector<Voxel*>>(ydim_, vector<Voxel*>(zdim_, nullptr)));
        occupied_cells_.clear();
        revision_ = dmf_next_volume_revision();
        return true;
    }
    template <typename PointT> bool addPointCloud(typename pcl::PointCloud<PointT>::Ptr) { return true; }

    tuple<int, int, int> getVoxel(float x, float y, float z)
    {
        return make_tuple(int(floor((x - xmin_) / xdelta_)), int(floor((y - ymin_) / ydelta_)), int(floor((z - zmin_) / zdelta_)));
    }
    unsigned long long int getHashId(int x, int y, int z)
    {
        unsigned long long int h = x;
        return (h << 40) ^ (y << 20) ^ z;
    }
    unsigned long long int getHash(float x, float y, float z)
    {
        int a, b, c;
        tie(a, b, c) = getVoxel(x, y, z);
        unsigned long long int h = a;
        return h << 40 ^ (b << 20) ^ c;
    }
    tuple<int, int, int> getVoxelCoords(unsigned long long int id)
    {
        constexpr unsigned long long int low20 = (1 << 20) - 1;
        return make_tuple(int(id >> 40), int(id >> 20 & low20), int(id & low20));
    }
    bool validCoords(int xid, int yid, int zid) { return xid >= 0 && yid >= 0 && zid >= 0 && xid < xdim_ && yid < ydim_ && zid < zdim_; }
    bool validPoints(float x, float y, float z) { return !(x >= xmax_ || y >= ymax_ || z >= zmax_ || x <= xmin_ || y <= ymin_ || z <= zmin_); }

    // xyz-only overload (:172-197): no validCoords guard in the reference; one is applied here instead of writing out of range
    bool integratePointCloud(pcl::PointCloud<pcl::PointXYZRGB>::Ptr cloud) { return integrate(cloud, nullptr); }
    bool integratePointCloud(pcl::PointCloud<pcl::PointXYZRGB>::Ptr cloud, pcl::PointCloud<pcl::Normal>::Ptr normals) { return integrate(cloud, normals.get()); }

    vector<unsigned long long int> getNeighborHashes(unsigned long long int hash, int K = 1)
    {
        int cx, cy, cz;
        tie(cx, cy, cz) = getVoxelCoords(hash);
        vector<unsigned long long int> out;
        for (int i = -K; i <= K; i++) for (int j = -K; j <= K; j++) for (int k = -K; k <= K; k++) {
            if (((i == j) == k) == 0) continue;                       // the reference's `i==j==k==0`
            if (validCoords(cx + i, cy + j, cz + k) && voxels_[cx + i][cy + j][cz + k] != nullptr) out.push_back(getHashId(cx + i, cy + j, cz + k));
        }
        return out;
    }

    // ---- additions (not in the reference): what the GPU engine needs to mirror this volume ----
    unsigned long long revision() const { return revision_; }
    void touch() { revision_ = dmf_next_volume_revision(); }     // call after editing voxels_ / normals by hand
    // CSR of the per-voxel normal lists in occupied_cells_ order
    void exportNormals(vector<uint32_t>& offsets, vector<float>& xyz)
    {
        offsets.assign(occupied_cells_.size() + 1, 0);
        xyz.clear();
        for (size_t i = 0; i < occupied_cells_.size(); i++) {
            int x, y, z;
            tie(x, y, z) = getVoxelCoords(occupied_cells_[i]);
            offsets[i] = (uint32_t)(xyz.size() / 3);
            for (const auto& n : voxels_[x][y][z]->normals) { xyz.push_back(n.normal[0]); xyz.push_back(n.normal[1]); xyz.push_back(n.normal[2]); }
        }
        offsets[occupied_cells_.size()] = (uint32_t)(xyz.size() / 3);
    }
    Voxel* voxelOf(size_t occupied_index)
    {
        int x, y, z;
        tie(x, y, z) = getVoxelCoords(occupied_cells_[occupied_index]);
        return voxels_[x][y][z];
    }

    private:
    bool integrate(pcl::PointCloud<pcl::PointXYZRGB>::Ptr cloud, pcl::PointCloud<pcl::Normal>* normals)
    {
        for (size_t i = 0; i < cloud->points.size(); i++) {
            const pcl::PointXYZRGB pt = cloud->points[i];
            if (!validPoints(pt.x, pt.y, pt.z)) continue;
            int x, y, z;
            tie(x, y, z) = getVoxel(pt.x, pt.y, pt.z);
            if (!validCoords(x, y, z)) continue;
            Voxel*& slot = voxels_[x][y][z];
            if (slot == nullptr) {
                occupied_cells_.push_back(getHashId(x, y, z));
                slot = normals ? new Voxel(pt, normals->points[i]) : new Voxel(pt);
            } else {
                slot->pts.push_back(pt);
                if (normals) slot->normals.push_back(normals->points[i]);
            }
        }
        revision_ = dmf_next_volume_revision();
        return true;    // the reference falls off the end of a bool function here (UB)
    }
};
