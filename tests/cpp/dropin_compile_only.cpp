// Compile-only check (tests/test_dropin.py, CPU tier): every template of DmfAlgorithms.hpp instantiates against the drop-in
// headers with the reference's argument types.  Never linked or run.
#include <vector>
#include <Eigen/Dense>
#include <Camera.hpp>
#include <Volume.hpp>
#include <RayTracingEngine.hpp>
#include <DmfAlgorithms.hpp>

template std::vector<uint8_t> dmf_dropin::segmentsCollide<VoxelVolume>(VoxelVolume&, const std::vector<Eigen::Vector3f>&, const std::vector<Eigen::Vector3f>&, bool);
template bool dmf_dropin::willCollide<VoxelVolume>(VoxelVolume&, Eigen::Vector3f, Eigen::Vector3f, bool);
template std::vector<uint8_t> dmf_dropin::collisionMatrix<VoxelVolume>(VoxelVolume&, const std::vector<Eigen::Affine3f>&, bool);
template std::vector<Eigen::Affine3f> dmf_dropin::optimizeCameraPositions<VoxelVolume>(VoxelVolume&, RayTracingEngine, const std::vector<Eigen::Affine3f>&, unsigned, unsigned, std::vector<uint32_t>*);
template Eigen::Affine3f dmf_dropin::optimizeCameraPosition<VoxelVolume>(VoxelVolume&, RayTracingEngine, int, Eigen::Affine3f);
template std::vector<Eigen::Affine3f> dmf_dropin::repositionCamerasSampled<VoxelVolume>(const std::vector<Eigen::Affine3f>&, VoxelVolume&, Camera);
template std::vector<unsigned long long int> dmf_dropin::setCover<VoxelVolume>(RayTracingEngine, VoxelVolume&, const std::vector<Eigen::Affine3f>&, int, bool);
