// dmf_distance.cuh -- builds the DMF_GRID_BYTE format: one byte per voxel of the padded index space holding
//     0            the voxel is occupied
//     d in 1..255  the voxel is empty and every voxel of the grid within Chebyshev (L-inf) index distance d-1 of it is empty,
//                  i.e. d = min(255, distance to the nearest OCCUPIED voxel).  The volume boundary is NOT a source.
// A probe that lands in a voxel with value d >= 2 therefore proves that every later probe of the same ray whose position
// differs by at most d-1 voxels (L-inf) is a miss IF it is in bounds.  Who proves "in bounds":
//   * the line-first marches (k_forward_line, march_collides' line loop) follow the ray only inside the slab interval in which the
//     line is >= 0.25 voxel inside the volume on every axis, and clamp every jump to the end of that interval -- the ray's
//     DIRECTION is known, so only the face it leaves through can end a jump, and the slab test already knows where that is;
//   * everything else (k_view_start, k_forward_dist, k_forward_skip on bytes, the exact-step skips of march_collides and
//     k_segments_collide) folds the voxel's distance to the outermost voxel layer in with byte_with_border() (dmf_device.cuh).
// (Round 1 and the first half of round 2 made the outermost layer a source of the transform itself.  That cost the line marches
// a geometric ramp of ~log2(distance) probes wherever a ray enters, leaves or starts near the boundary of the volume -- for a
// camera 25 voxels inside the volume looking at an object 150 voxels away: 25, 50, 100 ... instead of one jump.)
// The same passes build the macro-cell clearance field of the bit-grid march (k_forward_skip).
//
// Chebyshev distance is separable in the max-min sense:
//     min_{x',y',z'} max(|dx|,|dy|,|dz|) = min_{x'} max(|dx|, min_{y'} max(|dy|, min_{z'} |dz|))
// so three 1-D passes are exact.  Pass z is a plain nearest-source distance along a line (one warp per line, ballots);
// passes y and x evaluate  g(c) = min_i max(|c - i|, f(i))  along a line in O(n): one thread per line, two sweeps, each with
// a monotonic queue (see dt_sweep).  Round 1 used a per-voxel window scan, O(n * distance): 50 ms at 512^3; this is ~2 ms.
#pragma once
#include "dmf_device.cuh"

// ---- pass z on the voxel grid: out = min(255, distance along z to the nearest source of the same (x,y) line) ------------
// source = occupied.  One warp per line, 32 voxels per step: ballot of the sources, nearest set bit to the left by CLZ; then
// the same from the right end, folded with min.
__global__ void __launch_bounds__(256) k_dt_z(const VolDev v, unsigned char* __restrict__ out) {
    const unsigned lane = threadIdx.x & 31;
    const unsigned nlines = (unsigned)v.pdim[0] * (unsigned)v.pdim[1];
    const unsigned nz = (unsigned)v.pdim[2];
    const unsigned nchunk = (nz + 31) / 32;
    for (unsigned line = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); line < nlines; line += gridDim.x * (blockDim.x >> 5)) {
        const size_t base = (size_t)line * nz;
        unsigned carry = 255u;                                    // distance from the voxel just left of the chunk to its nearest source
        for (unsigned ch = 0; ch < nchunk; ch++) {
            const unsigned z = ch * 32 + lane;
            bool src = false;
            if (z < nz) {
                const size_t i = base + z;
                src = ((__ldg(v.bits + (i >> 5)) >> (i & 31)) & 1u) != 0u;
            }
            const unsigned m = __ballot_sync(0xffffffffu, src);
            const unsigned ml = m & (0xffffffffu >> (31 - lane)); // sources at or left of this lane
            const unsigned d = ml ? lane - (31u - (unsigned)__clz(ml)) : min(255u, carry + lane + 1u);
            if (z < nz) out[base + z] = (unsigned char)min(d, 255u);
            carry = m ? (unsigned)__clz(m) : min(255u, carry + 32u);   // distance from lane 31 of this chunk
        }
        carry = 255u;
        for (unsigned ch = nchunk; ch-- > 0;) {
            const unsigned z = ch * 32 + lane;
            const unsigned f = z < nz ? out[base + z] : 255u;
            const unsigned m = __ballot_sync(0xffffffffu, z < nz && f == 0u);
            const unsigned mr = m & (0xffffffffu << lane);        // sources at or right of this lane
            const unsigned d = mr ? (unsigned)(__ffs(mr) - 1) - lane : min(255u, carry + (32u - lane));
            if (z < nz && d < f) out[base + z] = (unsigned char)d;
            carry = m ? (unsigned)(__ffs(m) - 1) : min(255u, carry + 32u);    // distance from lane 0 of this chunk
        }
    }
}

// One direction of  g(c) = min_{i <= c} max(c - i, f(i))  over a strided line, evaluated for c = 0 .. n-1 in order.
// Every source i costs max(age, f(i)) with age = c - i: flat at f(i) until age reaches f(i), then growing with age.
//   * A source is useless once a younger one has a value f <= its own (dominated for ever): new sources pop such entries
//     off the young end, so the queue holds f strictly increasing from old to young -- and, positions increasing too,
//     the moment pos + f at which an entry turns "age-dominated" increases along the queue as well.
//   * Entries turn age-dominated from the old end, in order.  Of all age-dominated sources only the youngest matters
//     (smallest age, for ever): it is kept as `holder`, the others are dropped.
//   * g(c) = min(c - holder, f of the oldest queued entry).
// Every element is pushed and popped once: O(n) per line.  f < 256 strictly increasing bounds the queue by 256 entries, so
// it is a 256-entry ring indexed by unsigned chars.  VIRTUAL_BORDER adds a source with f = 0 at index -1 (cells outside
// the grid are blocked; the macro-cell field needs this, the voxel grid has no boundary sources at all).
// REVERSE walks c = n-1 .. 0 (sources at i >= c) and folds its result into what the forward sweep stored.
template <bool REVERSE, bool VIRTUAL_BORDER, bool FINAL, class Encode>
__device__ __forceinline__ void dt_sweep(const unsigned char* __restrict__ in, unsigned char* __restrict__ out, size_t base, size_t stride, int n, Encode enc) {
    unsigned ring[256];                                           // (pos << 8) | f
    unsigned char head = 0, tail = 0;                             // oldest entry at head, next free slot at tail
    int holder = VIRTUAL_BORDER ? -1 : -100000;                   // position (in sweep order) of the youngest age-dominated source
    // loads run a few elements ahead of the dependent queue updates
    constexpr int AHEAD = 4;
    unsigned pre[AHEAD];
#pragma unroll
    for (int j = 0; j < AHEAD; j++) { const int c = j; pre[j] = c < n ? in[base + (size_t)(REVERSE ? n - 1 - c : c) * stride] : 255u; }
    for (int c0 = 0; c0 < n; c0 += AHEAD) {
        unsigned cur[AHEAD];
#pragma unroll
        for (int j = 0; j < AHEAD; j++) cur[j] = pre[j];
#pragma unroll
        for (int j = 0; j < AHEAD; j++) { const int c = c0 + AHEAD + j; pre[j] = c < n ? in[base + (size_t)(REVERSE ? n - 1 - c : c) * stride] : 255u; }
#pragma unroll
        for (int j = 0; j < AHEAD; j++) {
            const int c = c0 + j;
            if (c >= n) break;
            const unsigned fc = cur[j];
            while (head != tail && (ring[(unsigned char)(tail - 1)] & 255u) >= fc) tail--;
            ring[tail++] = ((unsigned)c << 8) | fc;
            while (head != tail) {
                const unsigned e = ring[head];
                if ((int)(e >> 8) + (int)(e & 255u) > c) break;
                holder = (int)(e >> 8); head++;
            }
            unsigned g = (unsigned)min(c - holder, 255);
            if (head != tail) g = min(g, ring[head] & 255u);
            const size_t at = base + (size_t)(REVERSE ? n - 1 - c : c) * stride;
            if (REVERSE) { g = min(g, (unsigned)out[at]); if (FINAL) g = enc(at, g); }
            out[at] = (unsigned char)g;
        }
    }
}

struct EncodeNone { __device__ __forceinline__ unsigned operator()(size_t, unsigned g) const { return g; } };

// passes y and x over a dense [n0][n1][n2] byte array (n2 fastest): AXIS 1 walks n1, AXIS 0 walks n0; one thread per line,
// neighbouring threads own neighbouring n2 positions, so every step of a warp reads and writes 32 consecutive bytes.
// AXIS 2 walks n2 itself (uncoalesced; only used for the tiny macro-cell array).
template <int AXIS, bool VIRTUAL_BORDER, bool FINAL, class Encode>
__global__ void __launch_bounds__(128) k_dt_lines(const unsigned char* __restrict__ in, unsigned char* __restrict__ out, int n0, int n1, int n2, Encode enc) {
    const size_t nlines = AXIS == 0 ? (size_t)n1 * n2 : (AXIS == 1 ? (size_t)n0 * n2 : (size_t)n0 * n1);
    for (size_t line = blockIdx.x * (size_t)blockDim.x + threadIdx.x; line < nlines; line += (size_t)gridDim.x * blockDim.x) {
        size_t base, stride; int n;
        if (AXIS == 0) { base = line; stride = (size_t)n1 * n2; n = n0; }
        else if (AXIS == 1) { base = (line / n2) * (size_t)n1 * n2 + (line % n2); stride = (size_t)n2; n = n1; }
        else { base = line * (size_t)n2; stride = 1; n = n2; }
        dt_sweep<false, VIRTUAL_BORDER, false>(in, out, base, stride, n, enc);
        dt_sweep<true, VIRTUAL_BORDER, FINAL>(in, out, base, stride, n, enc);
    }
}

// ---- macro cells (8^3 voxels): clearance field of the bit-grid march (k_forward_skip) ------------------------------------
// seed[m] = 0 if the cell is blocked -- it holds an occupied voxel, or is not entirely inside [0,dim) -- else 255
__global__ void k_macro_seed(const VolDev v, const unsigned* __restrict__ macro_bits, unsigned char* __restrict__ seed) {
    const unsigned n = (unsigned)v.mdim[0] * (unsigned)v.mdim[1] * (unsigned)v.mdim[2];
    for (unsigned m = blockIdx.x * blockDim.x + threadIdx.x; m < n; m += gridDim.x * blockDim.x) {
        const unsigned z = m % (unsigned)v.mdim[2], y = (m / (unsigned)v.mdim[2]) % (unsigned)v.mdim[1], x = m / ((unsigned)v.mdim[2] * (unsigned)v.mdim[1]);
        const bool blocked = ((macro_bits[m >> 5] >> (m & 31)) & 1u) || 8 * ((int)x + 1) > v.dim[0] || 8 * ((int)y + 1) > v.dim[1] || 8 * ((int)z + 1) > v.dim[2];
        seed[m] = blocked ? 0 : 255;
    }
}
// clearance = 8*(D-1) - 0.25 voxels for D >= 2 (D = Chebyshev distance in cells to the nearest blocked cell, capped at 41:
// "farther than 40 cells"), else 0: from any point of the cell one can move that far (L-inf) and stay in empty, interior cells
__global__ void k_macro_clearance(const unsigned char* __restrict__ dist, float* __restrict__ clearance, unsigned n) {
    for (unsigned m = blockIdx.x * blockDim.x + threadIdx.x; m < n; m += gridDim.x * blockDim.x) {
        const int D = min((int)dist[m], 41);
        clearance[m] = D >= 2 ? 8.0f * (float)(D - 1) - 0.25f : 0.0f;
    }
}
