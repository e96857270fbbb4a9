// dmf_volume.cuh -- the march structures of a VoxelVolume built ON THE DEVICE from its occupied id list
// (occupied_cells_, reference include/Volume.hpp:57,216): bit grid, macro-cell bits, rank directory, rank -> ordinal table.
// Round 1 built these in host loops and uploaded ~50 MB of staging vectors per volume; now the host sends the ids (8 B per
// occupied voxel) -- or nothing at all when integratePointCloud itself ran on the GPU (dmf_integrate.cuh) -- and a multi-GPU
// group replicates a volume by broadcasting the ids GPU to GPU and rebuilding locally (dmf_comm.cuh).
#pragma once
#include "dmf_device.cuh"

// err[0] = smallest index of an id outside the grid, err[1] = smallest index whose voxel was already set (duplicate);
// both 0xFFFFFFFF when the list is clean
__global__ void k_vol_mark(const u64* __restrict__ ids, unsigned n_occ, const VolDev v, unsigned* __restrict__ bits, unsigned* __restrict__ macro, unsigned* __restrict__ err) {
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n_occ; i += gridDim.x * blockDim.x) {
        const u64 id = ids[i];
        const u64 x = id >> 40, y = (id >> 20) & 0xFFFFFu, z = id & 0xFFFFFu;                 // getVoxelCoords, Volume.hpp:158-165
        if (x >= (u64)v.dim[0] || y >= (u64)v.dim[1] || z >= (u64)v.dim[2]) { atomicMin(err, i); continue; }
        const unsigned idx = ((unsigned)x * (unsigned)v.pdim[1] + (unsigned)y) * (unsigned)v.pdim[2] + (unsigned)z;
        const unsigned bit = 1u << (idx & 31);
        if (atomicOr(bits + (idx >> 5), bit) & bit) atomicMin(err + 1, i);
        const unsigned m = (((unsigned)x >> 3) * (unsigned)v.mdim[1] + ((unsigned)y >> 3)) * (unsigned)v.mdim[2] + ((unsigned)z >> 3);
        const unsigned mb = 1u << (m & 31);
        if (!(macro[m >> 5] & mb)) atomicOr(macro + (m >> 5), mb);
    }
}

__global__ void k_vol_popc(const unsigned* __restrict__ bits, unsigned* __restrict__ cnt, size_t nwords) {
    for (size_t w = blockIdx.x * (size_t)blockDim.x + threadIdx.x; w < nwords; w += (size_t)gridDim.x * blockDim.x) cnt[w] = (unsigned)__popc(bits[w]);
}

// rank2occ[rank of voxel i in linear order] = i  (prefix = exclusive scan of the word popcounts)
__global__ void k_vol_rank2occ(const u64* __restrict__ ids, unsigned n_occ, const VolDev v, const unsigned* __restrict__ bits, const unsigned* __restrict__ prefix,
                               unsigned* __restrict__ rank2occ) {
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n_occ; i += gridDim.x * blockDim.x) {
        const u64 id = ids[i];
        const unsigned x = (unsigned)(id >> 40), y = (unsigned)((id >> 20) & 0xFFFFFu), z = (unsigned)(id & 0xFFFFFu);
        const unsigned idx = (x * (unsigned)v.pdim[1] + y) * (unsigned)v.pdim[2] + z;
        rank2occ[prefix[idx >> 5] + (unsigned)__popc(bits[idx >> 5] & ((1u << (idx & 31)) - 1u))] = i;
    }
}

// CSR sanity on the device (volumes that never visit the host, e.g. one received from a peer): offsets monotone
__global__ void k_vol_check_csr(const unsigned* __restrict__ noff, unsigned n_occ, unsigned* __restrict__ err) {
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n_occ; i += gridDim.x * blockDim.x)
        if (noff[i + 1] < noff[i]) atomicMin(err + 2, i);
    if (blockIdx.x == 0 && threadIdx.x == 0 && noff[0] != 0u) atomicMin(err + 2, 0u);
}

// ---- work order of the fast reverse march ---------------------------------------------------------------------------------
// k_reverse<FAST> gives every occupied voxel a thread, and a warp waits for its longest march.  occupied_cells_ order is the order
// points were integrated in -- for a scanned surface typically x-major, so that consecutive ordinals alternate between the near and
// the far wall of whatever the object is (two voxels per (x, y) column of a shell): half the lanes of a warp are occluded at once,
// the other half march to the end of the volume.  Marching the ordinals in MORTON order of their voxels puts 32 neighbours of one
// surface patch into a warp: their rays are near-parallel and end alike.  Results are bit sets over ordinals -- order-free.
__device__ __forceinline__ unsigned long long spread3_21(unsigned long long x) {      // 21 bits -> every third bit of 63
    x &= 0x1FFFFFull;
    x = (x | (x << 32)) & 0x1F00000000FFFFull;
    x = (x | (x << 16)) & 0x1F0000FF0000FFull;
    x = (x | (x << 8)) & 0x100F00F00F00F00Full;
    x = (x | (x << 4)) & 0x10C30C30C30C30C3ull;
    x = (x | (x << 2)) & 0x1249249249249249ull;
    return x;
}
__global__ void k_morton_keys(const unsigned long long* __restrict__ occ_ids, unsigned n, unsigned long long* __restrict__ keys, unsigned* __restrict__ vals) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long id = occ_ids[i];
    keys[i] = (spread3_21(id >> 40) << 2) | (spread3_21((id >> 20) & 0xFFFFFull) << 1) | spread3_21(id & 0xFFFFFull);
    vals[i] = i;
}
