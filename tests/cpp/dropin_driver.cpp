// dropin_driver.cpp -- headless driver written the way the reference's tests/*.cpp drivers are (tests/Raytracing.cpp:55-96,
// tests/SetCover.cpp:218-240), but against the DROP-IN headers: it includes <Volume.hpp> and <RayTracingEngine.hpp>,
// builds a VoxelVolume with setDimensions/setVolumeSize/constructVolume/integratePointCloud, constructs
// RayTracingEngine engine(cam) and calls the eight methods with the reference's signatures.  The scene comes from a
// binary file written by tests/test_dropin_gpu.py; the results go to another binary file that the test compares with
// the CPU oracle.  No viewer, no PCD I/O.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <iostream>
#include <vector>

#include <Volume.hpp>
#include <RayTracingEngine.hpp>
#include <DmfAlgorithms.hpp>

template <class T> static void rd(FILE* f, T* p, size_t n) { if (fread(p, sizeof(T), n, f) != n) { fprintf(stderr, "short read\n"); exit(2); } }
template <class T> static void wr(FILE* f, const T* p, size_t n) { if (fwrite(p, sizeof(T), n, f) != n) { fprintf(stderr, "short write\n"); exit(2); } }
static void wr_ids(FILE* f, bool found, const std::vector<unsigned long long int>& ids)
{
    int32_t fl = found ? 1 : 0; int64_t n = (int64_t)ids.size();
    wr(f, &fl, 1); wr(f, &n, 1);
    if (n) wr(f, ids.data(), ids.size());
}
static void wr_marks(FILE* f, VoxelVolume& volume)
{
    int64_t n = (int64_t)volume.occupied_cells_.size();
    wr(f, &n, 1);
    for (int64_t i = 0; i < n; i++) { Voxel* v = volume.voxelOf(i); int32_t view = v->view; uint8_t good = v->good; wr(f, &view, 1); wr(f, &good, 1); }
}
static void clear_marks(VoxelVolume& volume)
{
    for (size_t i = 0; i < volume.occupied_cells_.size(); i++) { Voxel* v = volume.voxelOf(i); v->view = 0; v->good = false; }
}

// the helper every reference driver carries (tests/SetCover.cpp:218-240), minus the greedy step
static vector<vector<unsigned long long int>> regionsCovered(RayTracingEngine engine, VoxelVolume& volume, vector<Affine3f> camera_locations)
{
    vector<vector<unsigned long long int>> regions_covered;
    for (size_t i = 0; i < camera_locations.size(); i++) {
        vector<unsigned long long int> good_points;
        bool found;
        tie(found, good_points) = engine.reverseRayTraceFast(volume, camera_locations[i], false);
        sort(good_points.begin(), good_points.end());
        regions_covered.push_back(good_points);
    }
    return regions_covered;
}

int main(int argc, char** argv)
{
    if (argc < 3) { fprintf(stderr, "usage: %s scene.bin out.bin\n", argv[0]); return 2; }
    FILE* in = fopen(argv[1], "rb");
    if (!in) { perror("scene"); return 2; }
    double bounds[6]; int32_t dims[3]; float Kf[9]; int32_t H, W; int64_t n_pts; int32_t n_poses, zdelta;
    rd(in, bounds, 6); rd(in, dims, 3); rd(in, Kf, 9); rd(in, &H, 1); rd(in, &W, 1); rd(in, &n_pts, 1);
    std::vector<float> pts(3 * n_pts), nrm(3 * n_pts);
    rd(in, pts.data(), pts.size()); rd(in, nrm.data(), nrm.size());
    rd(in, &n_poses, 1);
    std::vector<float> poses(12 * (size_t)n_poses);
    rd(in, poses.data(), poses.size()); rd(in, &zdelta, 1);
    fclose(in);

    pcl::PointCloud<pcl::PointXYZRGB>::Ptr cloud(new pcl::PointCloud<pcl::PointXYZRGB>);
    pcl::PointCloud<pcl::Normal>::Ptr normals(new pcl::PointCloud<pcl::Normal>);
    for (int64_t i = 0; i < n_pts; i++) {
        pcl::PointXYZRGB p; p.x = pts[3 * i]; p.y = pts[3 * i + 1]; p.z = pts[3 * i + 2];
        cloud->points.push_back(p);
        normals->points.push_back(pcl::Normal(nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2]));
    }
    // Setting up the volume (tests/Raytracing.cpp:66-75)
    VoxelVolume volume;
    volume.setDimensions(bounds[0], bounds[1], bounds[2], bounds[3], bounds[4], bounds[5]);
    volume.setVolumeSize(dims[0], dims[1], dims[2]);
    volume.constructVolume();
    volume.integratePointCloud(cloud, normals);
    vector<float> K(Kf, Kf + 9);
    Camera cam(K, H, W);
    RayTracingEngine engine(cam);
    vector<Affine3f> camera_locations;
    for (int i = 0; i < n_poses; i++) {
        Affine3f Q = Eigen::Affine3f::Identity();
        for (int r = 0; r < 3; r++) for (int c = 0; c < 4; c++) Q(r, c) = poses[12 * i + 4 * r + c];
        camera_locations.push_back(Q);
    }

    FILE* out = fopen(argv[2], "wb");
    if (!out) { perror("out"); return 2; }
    int64_t n_occ = (int64_t)volume.occupied_cells_.size();
    wr(out, &n_occ, 1);
    wr(out, volume.occupied_cells_.data(), volume.occupied_cells_.size());
    for (auto& T : camera_locations) {
        bool found; vector<unsigned long long int> ids;
        tie(found, ids) = engine.rayTraceAndGetPoints(volume, T, zdelta, false);      wr_ids(out, found, ids);
        tie(found, ids) = engine.rayTraceAndGetGoodPoints(volume, T);                  wr_ids(out, found, ids);   // defaults: zdelta=10, sparse=true
        tie(found, ids) = engine.reverseRayTraceFast(volume, T, false);                wr_ids(out, found, ids);
        tie(found, ids) = engine.reverseRayTrace(volume, T, false);                    wr_ids(out, found, ids);
        int32_t m = engine.rayTraceAndGetMinimum(volume, T);                           wr(out, &m, 1);            // defaults: zdelta=1, sparse=true
    }
    // mutating routines: marks must end up on the host Voxel objects
    clear_marks(volume);
    for (size_t i = 0; i < camera_locations.size(); i++) engine.rayTraceAndClassify(volume, camera_locations[i], zdelta, int(i) + 1, false);
    wr_marks(out, volume);
    clear_marks(volume);
    engine.rayTrace(volume, camera_locations[0], zdelta, true);
    wr_marks(out, volume);
    clear_marks(volume);
    engine.reverseRayTraceFast(volume, camera_locations[1 % n_poses], true);
    wr_marks(out, volume);
    clear_marks(volume);
    engine.rayTraceVolume(volume, camera_locations[0]);
    wr_marks(out, volume);
    // engine passed by value, as the reference's setCover() does
    auto regions = regionsCovered(engine, volume, camera_locations);
    for (auto& r : regions) wr_ids(out, !r.empty(), r);
    // helpers outside RayTracingEngine (DmfAlgorithms.hpp): willCollide for every ordered pair of camera centres the way
    // CameraPathGen.cpp:318-330 loops, singly and as one batched matrix; optimizeCameraPosition; setCover
    {
        auto matrix = dmf_dropin::collisionMatrix(volume, camera_locations);
        for (size_t x = 0; x < camera_locations.size(); x++) for (size_t y = 0; y < camera_locations.size(); y++) {
            if (x == y) continue;
            Vector3f a, b;
            for (int k = 0; k < 3; k++) { a(k) = camera_locations[x](k, 3); b(k) = camera_locations[y](k, 3); }
            uint8_t single = dmf_dropin::willCollide(volume, a, b) == true ? 1 : 0;
            if (single != matrix[x * camera_locations.size() + y]) { fprintf(stderr, "collisionMatrix != willCollide at %zu,%zu\n", x, y); return 3; }
        }
        wr(out, matrix.data(), matrix.size());
        std::vector<uint32_t> mids;
        auto moved = dmf_dropin::optimizeCameraPositions(volume, engine, camera_locations, 300, 600, &mids);
        Affine3f one = dmf_dropin::optimizeCameraPosition(volume, engine, 1, camera_locations[0]);
        for (int r = 0; r < 3; r++) for (int c = 0; c < 4; c++) if (one(r, c) != moved[0](r, c)) { fprintf(stderr, "optimizeCameraPosition: single != batched\n"); return 3; }
        wr(out, mids.data(), mids.size());
        for (auto& T : moved) { float p[12]; dmf_dropin::pose12(T, p); wr(out, p, 12); }
        auto cover = dmf_dropin::setCover(engine, volume, camera_locations, 1, false);
        wr_ids(out, !cover.empty(), cover);
    }
    // Volumes that come and go at ONE address (a stack local in a loop): the engine's device mirror must follow the volume,
    // never the address.  Round 0 / 2 hold the scene, round 1 is EMPTY (every routine must then do nothing, like the
    // reference -- including the ones that push/pull Voxel::view and Voxel::good).
    for (int round = 0; round < 3; round++) {
        VoxelVolume local;
        local.setDimensions(bounds[0], bounds[1], bounds[2], bounds[3], bounds[4], bounds[5]);
        local.setVolumeSize(dims[0], dims[1], dims[2]);
        local.constructVolume();
        if (round != 1) local.integratePointCloud(cloud, normals);
        int64_t addr = (int64_t)(intptr_t)&local;
        wr(out, &addr, 1);
        bool found; vector<unsigned long long int> ids;
        tie(found, ids) = engine.rayTraceAndGetPoints(local, camera_locations[0], zdelta, false);  wr_ids(out, found, ids);
        tie(found, ids) = engine.reverseRayTraceFast(local, camera_locations[0], true);           wr_ids(out, found, ids);
        engine.rayTrace(local, camera_locations[0], zdelta, true);
        engine.rayTraceAndClassify(local, camera_locations[0], zdelta, 3, true);
        engine.rayTraceVolume(local, camera_locations[0]);
        wr_marks(out, local);
        auto cover = dmf_dropin::setCover(engine, local, camera_locations, 1, false);
        wr_ids(out, !cover.empty(), cover);
    }
    {
        int32_t gpus = dmf_dropin::global().gpus();
        wr(out, &gpus, 1);
    }
    fclose(out);
    std::cout << "dropin_driver: " << n_occ << " occupied voxels, " << n_poses << " poses" << std::endl;
    return 0;
}
