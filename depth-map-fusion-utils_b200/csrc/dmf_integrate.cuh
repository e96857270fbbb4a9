// dmf_integrate.cuh -- K0: VoxelVolume::integratePointCloud(cloud, normals) (reference include/Volume.hpp:199-228) on the
// device: a point buffer -> occupied_cells_ in FIRST-INSERTION order + per-voxel normal lists in point order (CSR).
//
//   k_pt_key      per point: validPoints, getVoxel (exact: float filter + double fallback), validCoords -> padded linear index
//   k_pt_first    first[voxel] = min point index that landed in it               (dense uint32 over the padded index space)
//   k_pt_flag     flag[i] = 1 iff point i is the first of its voxel;  exclusive scan(flag) = the voxel's ordinal
//   k_pt_assign   occ_ids[ordinal] = getHashId(voxel);  first[voxel] := ordinal
//   k_pt_count / scan / k_pt_scatter / k_seg_sort / k_gather_normals   CSR of the normals, each list sorted by point index
#pragma once
#include "dmf_device.cuh"

__global__ void k_pt_key(const VolDev v, const float* __restrict__ xyz, size_t n, unsigned* __restrict__ key) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
        unsigned k = 0xFFFFFFFFu;
        if (in_bounds(v, x, y, z)) {                                                     // validPoints :207
            unsigned ne = 0;
            bool unsafe = false;
            int ix = voxel_index_f32(x, v.inv32[0], v.c32[0], v.err32[0], unsafe);
            int iy = voxel_index_f32(y, v.inv32[1], v.c32[1], v.err32[1], unsafe);
            int iz = voxel_index_f32(z, v.inv32[2], v.c32[2], v.err32[2], unsafe);
            if (unsafe) {
                ix = voxel_index(x, v.vmin[0], v.delta[0], v.inv[0], v.c0[0], v.eps[0], ne);
                iy = voxel_index(y, v.vmin[1], v.delta[1], v.inv[1], v.c0[1], v.eps[1], ne);
                iz = voxel_index(z, v.vmin[2], v.delta[2], v.inv[2], v.c0[2], v.eps[2], ne);
            }
            if (coords_valid(v, ix, iy, iz)) k = linear_index(v, ix, iy, iz);            // validCoords :211
        }
        key[i] = k;
    }
}

__global__ void k_pt_first(const unsigned* __restrict__ key, size_t n, unsigned* __restrict__ first) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        if (key[i] != 0xFFFFFFFFu) atomicMin(first + key[i], (unsigned)i);
}

__global__ void k_pt_flag(const unsigned* __restrict__ key, const unsigned* __restrict__ first, size_t n, unsigned* __restrict__ flag) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        flag[i] = (key[i] != 0xFFFFFFFFu && first[key[i]] == (unsigned)i) ? 1u : 0u;
}

__global__ void k_pt_assign(const VolDev v, const unsigned* __restrict__ key, const unsigned* __restrict__ flag, const unsigned* __restrict__ ord,
                            size_t n, unsigned* __restrict__ first, u64* __restrict__ occ_ids) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        if (!flag[i]) continue;
        const unsigned k = key[i], o = ord[i];
        const unsigned pz = (unsigned)v.pdim[2], py = (unsigned)v.pdim[1];
        const int z = (int)(k % pz), y = (int)((k / pz) % py), x = (int)(k / (pz * py));
        occ_ids[o] = voxel_id(x, y, z);
        first[k] = o;
    }
}

__global__ void k_pt_count(const unsigned* __restrict__ key, const unsigned* __restrict__ first, size_t n, unsigned* __restrict__ cnt) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        if (key[i] != 0xFFFFFFFFu) atomicAdd(cnt + first[key[i]], 1u);
}

__global__ void k_pt_scatter(const unsigned* __restrict__ key, const unsigned* __restrict__ first, const unsigned* __restrict__ noff, size_t n,
                             unsigned* __restrict__ cursor, unsigned* __restrict__ pidx) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        if (key[i] == 0xFFFFFFFFu) continue;
        const unsigned o = first[key[i]];
        pidx[noff[o] + atomicAdd(cursor + o, 1u)] = (unsigned)i;
    }
}

// every voxel's list of point indices ascending (= the order push_back saw them): heap sort, one thread per voxel
__global__ void k_seg_sort(const unsigned* __restrict__ noff, unsigned n_occ, unsigned* __restrict__ pidx) {
    const unsigned o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= n_occ) return;
    unsigned* a = pidx + noff[o];
    const int m = (int)(noff[o + 1] - noff[o]);
    if (m < 2) return;
    auto sift = [&](int root, int end) {
        for (;;) {
            int child = 2 * root + 1;
            if (child > end) break;
            if (child + 1 <= end && a[child] < a[child + 1]) child++;
            if (a[root] >= a[child]) break;
            unsigned t = a[root]; a[root] = a[child]; a[child] = t;
            root = child;
        }
    };
    for (int s = (m - 2) / 2; s >= 0; s--) sift(s, m - 1);
    for (int e = m - 1; e > 0; e--) { unsigned t = a[0]; a[0] = a[e]; a[e] = t; sift(0, e - 1); }
}

__global__ void k_gather_normals(const unsigned* __restrict__ pidx, const float* __restrict__ nrm, size_t m, float* __restrict__ out) {
    for (size_t s = blockIdx.x * (size_t)blockDim.x + threadIdx.x; s < m; s += (size_t)gridDim.x * blockDim.x) {
        const size_t i = pidx[s];
        out[3 * s] = nrm[3 * i]; out[3 * s + 1] = nrm[3 * i + 1]; out[3 * s + 2] = nrm[3 * i + 2];
    }
}

// ---- device-wide exclusive scan of uint32 (blocks of 2048 elements, recursive on the block sums) ----------------
constexpr int SCAN_THREADS = 512, SCAN_ITEMS = 4, SCAN_BLOCK = SCAN_THREADS * SCAN_ITEMS;

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_block(const unsigned* __restrict__ in, unsigned* __restrict__ out, size_t n, unsigned* __restrict__ block_sums) {
    __shared__ unsigned s_warp[SCAN_THREADS / 32];
    const size_t base = (size_t)blockIdx.x * SCAN_BLOCK + (size_t)threadIdx.x * SCAN_ITEMS;
    unsigned v[SCAN_ITEMS], sum = 0;
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; j++) { v[j] = base + j < n ? in[base + j] : 0u; sum += v[j]; }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned x = sum;
    for (int o = 1; o < 32; o <<= 1) { unsigned y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    if (lane == 31) s_warp[warp] = x;
    __syncthreads();
    if (warp == 0) {
        unsigned w = lane < SCAN_THREADS / 32 ? s_warp[lane] : 0u;
        for (int o = 1; o < 32; o <<= 1) { unsigned y = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += y; }
        if (lane < SCAN_THREADS / 32) s_warp[lane] = w;
    }
    __syncthreads();
    unsigned run = (warp ? s_warp[warp - 1] : 0u) + x - sum;
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; j++) { if (base + j < n) out[base + j] = run; run += v[j]; }
    if (threadIdx.x == SCAN_THREADS - 1 && block_sums) block_sums[blockIdx.x] = run;
}

__global__ void k_scan_add(unsigned* __restrict__ out, size_t n, const unsigned* __restrict__ block_offsets) {
    const size_t i = (size_t)blockIdx.x * SCAN_BLOCK + threadIdx.x;
    const unsigned add = block_offsets[blockIdx.x];
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; j++) { const size_t k = i + (size_t)j * SCAN_THREADS; if (k < n) out[k] += add; }
}
