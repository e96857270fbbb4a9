// dmf_distance.cuh -- builds the DMF_GRID_BYTE format: one byte per voxel of the padded index space holding
//     0            the voxel is occupied
//     d in 1..255  the voxel is empty and every voxel within Chebyshev (L-inf) index distance d-1 of it is empty AND
//                  interior (all indices in [1, dim-2]), i.e. d = min(255, distance to the nearest "blocked" voxel),
//                  blocked = occupied, or on the outermost voxel layer, or outside.  Empty voxels of the outermost layer
//                  and of the padding plane hold 1.
// A probe that lands in a voxel with value d >= 2 therefore proves that every later probe of the same ray whose position
// differs by at most d-1 voxels (L-inf) is an in-bounds miss (k_forward_skip).
//
// Chebyshev distance is separable in the max-min sense:
//     min_{x',y',z'} max(|dx|,|dy|,|dz|) = min_{x'} max(|dx|, min_{y'} max(|dy|, min_{z'} |dz|))
// so three 1-D passes are exact: z (two sweeps per line), then y and x (window scan with early exit at |d| >= best).
#pragma once
#include "dmf_device.cuh"

// pass 1: along z, distance to the nearest occupied voxel of the same (x,y) line; 255 = none within 254
__global__ void k_dt_z(const VolDev v, unsigned char* __restrict__ out) {
    const unsigned line = blockIdx.x * blockDim.x + threadIdx.x;             // (x, y) over the padded space
    const unsigned nlines = (unsigned)v.pdim[0] * (unsigned)v.pdim[1];
    if (line >= nlines) return;
    const unsigned nz = (unsigned)v.pdim[2];
    const size_t base = (size_t)line * nz;
    unsigned d = 255;
    for (unsigned z = 0; z < nz; z++) {
        const size_t i = base + z;
        const bool occ = (__ldg(v.bits + (i >> 5)) >> (i & 31)) & 1u;
        d = occ ? 0u : min(255u, d + 1u);
        out[i] = (unsigned char)d;
    }
    d = 255;
    for (unsigned z = nz; z-- > 0;) {
        const size_t i = base + z;
        const unsigned f = out[i];
        d = f == 0 ? 0u : min(255u, d + 1u);
        if (d < f) out[i] = (unsigned char)d;
    }
}

// passes 2 and 3: out(p) = min_j max(|j|, in(p + j*stride)) along one axis (AXIS 1 = y, 0 = x).
// FINAL also folds in the distance to the outermost voxel layer and the encoding rules of the header comment.
template <int AXIS, bool FINAL>
__global__ void k_dt_axis(const VolDev v, const unsigned char* __restrict__ in, unsigned char* __restrict__ out) {
    const size_t n = (size_t)v.pdim[0] * v.pdim[1] * v.pdim[2];
    const unsigned ny = (unsigned)v.pdim[1], nz = (unsigned)v.pdim[2];
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const unsigned z = (unsigned)(i % nz);
        const size_t t = i / nz;
        const unsigned y = (unsigned)(t % ny), x = (unsigned)(t / ny);
        const int c = AXIS == 1 ? (int)y : (int)x;
        const int len = AXIS == 1 ? (int)ny : v.pdim[0];
        const size_t stride = AXIS == 1 ? (size_t)nz : (size_t)ny * nz;
        int best = in[i];
        for (int d = 1; d < best; d++) {
            if (c - d >= 0) best = min(best, max(d, (int)__ldg(in + i - (size_t)d * stride)));
            if (c + d < len) best = min(best, max(d, (int)__ldg(in + i + (size_t)d * stride)));
        }
        if (FINAL) {
            const int dx = v.dim[0], dy = v.dim[1], dz = v.dim[2];
            if ((int)x >= dx || (int)y >= dy || (int)z >= dz) best = 1;                  // padding plane: empty, no clearance
            else if (best != 0) {
                const int b = min(min(min((int)x, dx - 1 - (int)x), min((int)y, dy - 1 - (int)y)), min((int)z, dz - 1 - (int)z));
                best = max(1, min(best, b));                                             // outermost layer: b = 0 -> 1
            }
        }
        out[i] = (unsigned char)best;
    }
}
