// dmf_forward.cuh -- K1: the forward per-pixel sampled march (one thread per ray) and its helper kernels.
//
// Restates, for all five forward routines of the reference (include/RayTracingEngine.hpp:229-494), the loop
//     for z_depth: for r: for c:  if found[r][c] continue;  p = T * projectPoint(r,c,z_depth);
//                                 if !validPoints(p) continue;  voxel = voxels_[getVoxel(p)];  if voxel: first hit
// with the loop nest inverted (ray outermost, z_depth innermost) -- legal because a ray's samples never depend on
// other rays; only the *order* of the returned id list depends on the nest, and k_order_ids restores it.
#pragma once
#include "dmf_device.cuh"

// Per-launch constants of the march.
struct FwdArgs {
    VolDev vol;
    AngleTest angle;
    const float* __restrict__ poses;   // [n_views][12]
    int view0;                         // first view of this launch (a batch may be marched as several launches, blockIdx.z = view - view0)
    // projectPoint tables (Camera.hpp:24-31), built once per (K, H, W, z0, zdelta, stride) by k_build_tables:
    const float* __restrict__ xtab;    // [S][Wc]  (float)(z_k*((double)c-cx)/fx)
    const float* __restrict__ ytab;    // [S][Hc]  (float)(z_k*((double)r-cy)/fy)
    const float* __restrict__ ztab;    // [S]      (float)z_k,  z_k = (z0 + k*zdelta)*0.001
    int S, Wc, Hc, W, H, cstride, rstride, z0, zdelta;
    // empty-space skipping (k_forward_skip): approximate ray slopes and the macro-cell clearance field
    const float* __restrict__ dcx;     // [Wc] (float)((c-cx)/fx)   -- approximate, only steers the skipping
    const float* __restrict__ dcy;     // [Hc] (float)((r-cy)/fy)
    const float* __restrict__ clearance; // [mdim_x][mdim_y][mdim_z] voxels that can be crossed (L-inf) from anywhere in the cell
                                       //  while provably staying in empty, in-bounds macro cells (0 = evaluate exactly)
    float dcx_max, dcy_max;            // max |dcx|, |dcy| (for the per-view error bound)
    const int* __restrict__ kstart;    // [n_views] probes 0..kstart-1 of EVERY ray of the view are provably in-bounds misses (k_view_start); -1 = view must not skip
    float* veps;                       // [n_views] the view's bound on |reference sample - line point| in voxels (k_view_start)
    // k_forward_line reads everything it needs per view from ONE 64-byte record (k_view_start writes it): rows 0..2 = the pose,
    // row 3 = {e_safe = veps + 2^-10, kstart as int bits, 0, 0}: four 16-byte loads from L1 instead of a shared-memory copy + barrier
    float4* viewrec;                   // [n_views][4]
    // loop constants the host folds once (the kernel runs at 32 registers and would otherwise re-derive them from the constant bank per probe)
    float z0m, zdm, Sf;                // z0 * 0.001, zdelta * 0.001, (float)S
    unsigned pnyz, bias, last;         // pdim_y * pdim_z;  0x4B400000 * (pnyz + pdim_z + 1) (the three shifter offsets, folded);  n_cells - 1
    // k_tile_start: one 64-bit record per view and 4x4 tile of lattice rays -- the sample intervals of the tile's rays (slab test, conservative
    // for all 16) and the result of the tile's cone pre-march; five 12-bit fields, see tile_pack()
    const u64* __restrict__ tile_rec;  // [n_views][tiles_y][tiles_x]
    int tiles_x, tiles_per_view;
    // outputs (may be null)
    int* depth;                        // [n_views][H][W]
    unsigned short* depth16;           // [n_views][H][W]  same, 0xFFFF = none
    float* points;                     // [n_views][H][W][3]
    u64* hit_voxel;                    // [n_views][H][W]
    unsigned* vis;                     // bitset over occupied order, row of view v at vis + v * vis_stride32
    int vis_words32;                   // 32-bit words of visibility per view
    unsigned vis_stride32;             // row pitch (== vis_words32 for a dense [n_views][words] array; the gathered buffer of a sharded sweep interleaves ranks)
    PubTable pub;                      // sharded sweep: push finished rows to the peer GPUs (pub.enabled)
    int* found_any;                    // [n_views]
    int* min_depth;                    // [n_views]  (INT_MAX-initialised; MINIMUM mode)
    // discovery-order bookkeeping (ids requested)
    unsigned* first_key;               // [n_views][n_occ]  min over emitting rays of (k<<21 | lattice index)
    unsigned* ray_key;                 // [n_views][Wc*Hc]  key of each emitting ray, 0xFFFFFFFF otherwise
    int* ray_occ;                      // [n_views][Wc*Hc]  occupied ordinal of each ray's hit
    // Voxel::view / Voxel::good (CLASSIFY / MARK)
    int* first_view;                   // [n_occ] min batch-local view index that hit the voxel (INT_MAX-initialised)
    unsigned* good_bits;               // [ceil(n_occ/32)]
    int* view_mark;                    // [n_occ]
    // carve mode (DMF_FWD_CARVE): bit grid over the padded index space, same layout as vol.bits; every in-bounds sample a ray
    // visits up to and including its first hit sets the bit of the voxel it falls in ("observed": free where not occupied)
    unsigned* observed;
    u64* counters;                     // DMF_CNT_*
};

// xtab/ytab/ztab: exact IEEE double mul/div then round to float, as the tuple<float,float,float> return does.
__global__ void k_build_tables(float* xtab, float* ytab, float* ztab, float* dcx, float* dcy, int S, int Wc, int Hc, int cstride, int rstride,
                               int z0, int zdelta, double fx, double cx, double fy, double cy) {
    int per = Wc + Hc + 1;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < Wc + Hc; j += gridDim.x * blockDim.x) {
        if (j < Wc) dcx[j] = (float)(((double)(j * cstride) - cx) / fx);
        else dcy[j - Wc] = (float)(((double)((j - Wc) * rstride) - cy) / fy);
    }
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < (long long)S * per; i += (long long)gridDim.x * blockDim.x) {
        int k = (int)(i / per), j = (int)(i % per);
        double z = __dmul_rn((double)(z0 + k * zdelta), 0.001);
        if (j < Wc) {
            double c = (double)(j * cstride);
            xtab[(size_t)k * Wc + j] = __double2float_rn(__ddiv_rn(__dmul_rn(z, __dsub_rn(c, cx)), fx));
        } else if (j < Wc + Hc) {
            int jr = j - Wc;
            double r = (double)(jr * rstride);
            ytab[(size_t)k * Hc + jr] = __double2float_rn(__ddiv_rn(__dmul_rn(z, __dsub_rn(r, cy)), fy));
        } else {
            ztab[k] = __double2float_rn(z);
        }
    }
}

#ifndef DMF_VIS_READ_FIRST
#define DMF_VIS_READ_FIRST 0                      // 1: read the visibility word before the atomicOr and skip it if the bit is set (measured 1.6 % slower:
#endif                                            //    the read is one more dependent L2 round trip at the very end of the warp's life)
// Per-ray epilogue shared by k_forward and k_forward_skip: outputs, visibility, marks, discovery keys, counters.
template <int MODE>
__device__ __forceinline__ void forward_epilogue(const FwdArgs& a, unsigned long long* s_cnt, int view, int ci, int ri, bool active, int hit_k,
                                                 int hx, int hy, int hz, float hpx, float hpy, float hpz, float m03, float m13, float m23,
                                                 unsigned n_samples, unsigned n_inb, unsigned n_exact, unsigned n_f64, unsigned n_skip) {
    const VolDev& v = a.vol;
    const int lane = threadIdx.x & 31;
    const bool hit = hit_k >= 0;
    const int z_depth = a.z0 + hit_k * a.zdelta;
    if (active) {
        // 32-bit pixel index: the host entry points keep n_views * H * W below 2^32 per launch
        unsigned pix = ((unsigned)view * (unsigned)a.H + (unsigned)(ri * a.rstride)) * (unsigned)a.W + (unsigned)(ci * a.cstride);
#ifdef DMF_CHECKED
        if (ri * a.rstride >= a.H || ci * a.cstride >= a.W) { atomicAdd(counter_slot(a.counters) + 10, 1ull); pix = 0; }
#endif
        // streaming stores (evict-first): the per-pixel outputs are never read back by the march and must not push the distance
        // bytes out of the L2 (135 MB of bytes against 126 MB of L2: every output line kept is a grid line lost)
        if (a.depth) __stcs(a.depth + pix, hit ? z_depth : -1);
        if (a.depth16) __stcs(a.depth16 + pix, hit ? (unsigned short)z_depth : (unsigned short)0xFFFFu);
        if (a.hit_voxel) __stcs(a.hit_voxel + pix, hit ? voxel_id(hx, hy, hz) : ~0ull);
        if (a.points) {
            float* const pp = a.points + 3ull * pix;
            __stcs(pp, hit ? hpx : 0.f); __stcs(pp + 1, hit ? hpy : 0.f); __stcs(pp + 2, hit ? hpz : 0.f);
        }
    }
    unsigned ties = 0;
    if (MODE == 4) {
        // one atomic per warp at most, and none once the view's minimum is already lower (same-address atomics serialise in L2)
        const int zw = __reduce_min_sync(0xffffffffu, hit ? z_depth : 0x7fffffff);
        if (lane == 0 && zw != 0x7fffffff && zw < *((volatile int*)(a.min_depth + view))) atomicMin(a.min_depth + view, zw);
    } else {
        const unsigned li = ((unsigned)view * (unsigned)a.Hc + (unsigned)ri) * (unsigned)a.Wc + (unsigned)ci;
        bool emit = false;
        int occ = -1;
        if (hit) {
            occ = occupied_ordinal(v, hx, hy, hz);
            DMF_CHECK_IDX(occ, v.n_occ, a.counters);
            if (a.found_any) raise_flag(a.found_any + view);
            bool good = false;
            if (MODE == 1 || MODE == 2) {
                bool need = (z_depth >= 250 && z_depth <= 600);                       // :363, :429
                if (MODE == 2) need = need && !((a.good_bits[occ >> 5] >> (occ & 31)) & 1u);   // :356 (monotone, so a stale read is fine)
                if (need) {
                    float dx, dy, dz;
                    view_direction(v, hpx, hpy, hpz, m03, m13, m23, dx, dy, dz);
                    good = any_normal_faces(v, a.angle, occ, dx, dy, dz, ties);
                }
            }
            if (MODE == 0) emit = true;
            if (MODE == 1) emit = good;
            if (MODE == 2) {
                atomicMin(a.first_view + occ, view);                                  // :354-355, resolved by k_apply_first_view
                if (good) atomicOr(a.good_bits + (occ >> 5), 1u << (occ & 31));       // :366
            }
            if (MODE == 3) a.view_mark[occ] = 1;                                       // :302
        }
        if (a.vis) {
            // neighbouring rays mostly hit the same few voxels: one atomicOr per distinct voxel of the warp, and none if a
            // (possibly stale, the bits only ever get set) cached read already shows the bit
            const unsigned peers = __match_any_sync(__activemask(), emit ? occ : -1);
            if (emit && (unsigned)(threadIdx.x & 31) == (unsigned)(__ffs(peers) - 1)) {
                unsigned* w = a.vis + ((size_t)(unsigned)view * a.vis_stride32 + ((unsigned)occ >> 5));
                const unsigned m = 1u << (occ & 31);
#if DMF_VIS_READ_FIRST
                if (!(*w & m)) atomicOr(w, m);
#else
                atomicOr(w, m);
#endif
            }
        }
        if (active && a.ray_key) {
            unsigned key = emit ? (((unsigned)hit_k << 21) | (unsigned)(ri * a.Wc + ci)) : 0xFFFFFFFFu;
            a.ray_key[li] = key;
            a.ray_occ[li] = occ;
            if (emit) atomicMin(a.first_key + (size_t)view * v.n_occ + occ, key);
        }
    }

    // ---- counters.  Instrumentation the reference does not have: a.counters == null (DMF_FWD_NO_COUNTERS) compiles to one
    // uniform branch.  Otherwise: warp reduce (REDUX), then lane j adds counter j -- ONE vector RED over the few non-zero lanes
    // into one 128-byte line of the 256 counter replicas (round 1 issued up to seven scalar atomics from lane 0, each with its
    // own test and address: ~100 of the ~900 warp-instructions of a ray-warp of k_forward_line).
    // No block barrier: warps whose rays ended early must not sit on a barrier waiting for the longest ray of the block.
    if (a.counters) {
        const unsigned c0 = __reduce_add_sync(0xffffffffu, n_samples), c1 = __reduce_add_sync(0xffffffffu, n_inb);
        const unsigned c2 = __reduce_add_sync(0xffffffffu, hit ? 1u : 0u), c9 = __reduce_add_sync(0xffffffffu, n_skip);
        const unsigned c358 = __reduce_add_sync(0xffffffffu, n_exact | ties | n_f64);      // almost always 0: reduce individually only then
        unsigned c3 = 0, c5 = 0, c8 = 0;
        if (c358) { c3 = __reduce_add_sync(0xffffffffu, n_exact); c5 = __reduce_add_sync(0xffffffffu, ties); c8 = __reduce_add_sync(0xffffffffu, n_f64); }
        unsigned mine = lane == 0 ? c0 : (lane == 1 ? c1 : (lane == 2 ? c2 : (lane == 9 ? c9 : 0u)));
        if (c358) mine = lane == 3 ? c3 : (lane == 5 ? c5 : (lane == 8 ? c8 : mine));
        if (mine) atomicAdd(a.counters + ((((unsigned)view * 37u + (unsigned)ri * 5u + (unsigned)(ci >> 3)) & (DMF_COUNTER_SLOTS - 1u)) * DMF_COUNTER_STRIDE + (unsigned)lane), (unsigned long long)mine);
    }
    (void)s_cnt;
}


constexpr int FWD_THREADS = 256;      // 8 warps: 4 across x 2 down, each warp an 8x4 pixel tile
constexpr int FWD_TILE_W = 32, FWD_TILE_H = 8;
constexpr int FWD_CHUNK = 32;         // z-steps staged in shared memory per round

// MODE: DMF_MODE_*, FMT: DMF_GRID_*, CARVE: also record every visited in-bounds sample's voxel in a.observed
template <int MODE, int FMT, bool CARVE>
__global__ void __launch_bounds__(FWD_THREADS) k_forward(const FwdArgs a) {
    __shared__ float sx[FWD_CHUNK][FWD_TILE_W];
    __shared__ float sy[FWD_CHUNK][FWD_TILE_H];
    __shared__ __align__(16) float sz[FWD_CHUNK][4];   // m02*z, m12*z, m22*z
    __shared__ unsigned long long s_cnt[5];

    const int view = blockIdx.z + a.view0;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int lc = (warp & 3) * 8 + (lane & 7);   // column within the tile
    const int lr = (warp >> 2) * 4 + (lane >> 3); // row within the tile
    const int ci = blockIdx.x * FWD_TILE_W + lc;  // lattice column
    const int ri = blockIdx.y * FWD_TILE_H + lr;  // lattice row
    const bool active = ci < a.Wc && ri < a.Hc;
    if (threadIdx.x < 5) s_cnt[threadIdx.x] = 0ull;

    const float* P = a.poses + 12 * (size_t)view;
    const float m00 = __ldg(P + 0), m01 = __ldg(P + 1), m02 = __ldg(P + 2), m03 = __ldg(P + 3);
    const float m10 = __ldg(P + 4), m11 = __ldg(P + 5), m12 = __ldg(P + 6), m13 = __ldg(P + 7);
    const float m20 = __ldg(P + 8), m21 = __ldg(P + 9), m22 = __ldg(P + 10), m23 = __ldg(P + 11);
    const VolDev& v = a.vol;

    // loop-invariant copies of the per-axis constants (kept in registers / uniform registers by the compiler)
    const float lo0 = v.lo[0], lo1 = v.lo[1], lo2 = v.lo[2], hi0 = v.hi[0], hi1 = v.hi[1], hi2 = v.hi[2];
    const float in0 = v.inv32[0], in1 = v.inv32[1], in2 = v.inv32[2], cc0 = v.c32[0], cc1 = v.c32[1], cc2 = v.c32[2];
    const float er0 = v.err32[0], er1 = v.err32[1], er2 = v.err32[2];
    const unsigned pny = (unsigned)v.pdim[1], pnz = (unsigned)v.pdim[2];
    const unsigned* __restrict__ gbits = v.bits;
    const unsigned char* __restrict__ gbytes = v.bytes;

    bool done = !active;
    int hit_k = -1, hx = 0, hy = 0, hz = 0;
    float hpx = 0.f, hpy = 0.f, hpz = 0.f;
    unsigned n_samples = 0, n_inb = 0, n_exact = 0, n_f64 = 0;

    for (int k0 = 0; k0 < a.S; k0 += FWD_CHUNK) {
        bool stop = done;
        if (MODE == 4) {   // rayTraceAndGetMinimum returns at the first hit plane: later planes cannot matter
            int cur = *((volatile int*)(a.min_depth + view));
            stop = stop || (a.z0 + k0 * a.zdelta > cur);
        }
        if (__syncthreads_and(stop)) break;   // also fences the previous round's shared-memory reads
        const int nk = min(FWD_CHUNK, a.S - k0);
        for (int i = threadIdx.x; i < FWD_CHUNK * FWD_TILE_W; i += FWD_THREADS) {
            int kk = i / FWD_TILE_W, cc = i % FWD_TILE_W, col = blockIdx.x * FWD_TILE_W + cc;
            sx[kk][cc] = (kk < nk && col < a.Wc) ? __ldg(a.xtab + (size_t)(k0 + kk) * a.Wc + col) : 0.f;
        }
        {
            int i = threadIdx.x;   // FWD_CHUNK*FWD_TILE_H == FWD_THREADS
            int kk = i / FWD_TILE_H, rr = i % FWD_TILE_H, row = blockIdx.y * FWD_TILE_H + rr;
            sy[kk][rr] = (kk < nk && row < a.Hc) ? __ldg(a.ytab + (size_t)(k0 + kk) * a.Hc + row) : 0.f;
        }
        if (threadIdx.x < FWD_CHUNK) {
            float z = (threadIdx.x < nk) ? __ldg(a.ztab + k0 + threadIdx.x) : 0.f;
            sz[threadIdx.x][0] = __fmul_rn(m02, z); sz[threadIdx.x][1] = __fmul_rn(m12, z); sz[threadIdx.x][2] = __fmul_rn(m22, z);
        }
        __syncthreads();
        if (!done) {
            const float* psx = &sx[0][lc];
            const float* psy = &sy[0][lr];
            const float4* psz = reinterpret_cast<const float4*>(&sz[0][0]);
            int kk = 0;
#pragma unroll 1
            for (; kk < nk; kk++, psx += FWD_TILE_W, psy += FWD_TILE_H, psz++) {
                const float xf = *psx, yf = *psy;
                const float4 zt = *psz;
                // transformPoints (Camera.hpp:39-45), rule E1: ((m0*x + m1*y) + m2*z) + m3 with separate roundings
                const float px = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m00, xf), __fmul_rn(m01, yf)), zt.x), m03);
                const float py = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m10, xf), __fmul_rn(m11, yf)), zt.y), m13);
                const float pz = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m20, xf), __fmul_rn(m21, yf)), zt.z), m23);
                // validPoints (Volume.hpp:230-233): outside => skip this sample, the ray goes on
                if (!(px > lo0 && px < hi0 && py > lo1 && py < hi1 && pz > lo2 && pz < hi2)) continue;
                n_inb++;
                // getVoxel (Volume.hpp:150-156): float filter, exact double path only near an integer quotient
                bool unsafe = false;
                int ix = voxel_index_f32(px, in0, cc0, er0, unsafe);
                int iy = voxel_index_f32(py, in1, cc1, er1, unsafe);
                int iz = voxel_index_f32(pz, in2, cc2, er2, unsafe);
                if (unsafe) {
                    n_f64++;
                    ix = voxel_index(px, v.vmin[0], v.delta[0], v.inv[0], v.c0[0], v.eps[0], n_exact);
                    iy = voxel_index(py, v.vmin[1], v.delta[1], v.inv[1], v.c0[1], v.eps[1], n_exact);
                    iz = voxel_index(pz, v.vmin[2], v.delta[2], v.inv[2], v.c0[2], v.eps[2], n_exact);
                }
                // voxels_[xid][yid][zid] != nullptr: indices are in [0,dim] here, the padded grid covers index == dim
                unsigned idx = ((unsigned)ix * pny + (unsigned)iy) * pnz + (unsigned)iz;
                DMF_CHECK_IDX(idx, v.n_cells, a.counters);
                if (CARVE) {
                    // read first: after the first few views of a sweep nearly every bit is already set, the word sits in
                    // L1/L2, and a stale 0 only costs a redundant (idempotent) atomicOr
                    unsigned* const ow = a.observed + (idx >> 5);
                    const unsigned om = 1u << (idx & 31);
                    if (!(*ow & om)) atomicOr(ow, om);
                }
                const bool occ_here = FMT == 0 ? ((__ldg(gbits + (idx >> 5)) >> (idx & 31)) & 1u) != 0u : __ldg(gbytes + idx) == 0;
                if (occ_here) {
                    hit_k = k0 + kk; hx = ix; hy = iy; hz = iz; hpx = px; hpy = py; hpz = pz;
                    done = true;
                    kk++;
                    break;
                }
            }
            n_samples += (unsigned)kk;
        }
    }

    forward_epilogue<MODE>(a, s_cnt, view, ci, ri, active, hit_k, hx, hy, hz, hpx, hpy, hpz, m03, m13, m23, n_samples, n_inb, n_exact, n_f64, 0u);
    if (a.pub.enabled) publish_view_row(a.pub, reinterpret_cast<u64*>(a.vis + (size_t)(unsigned)view * a.vis_stride32), view, gridDim.x * gridDim.y);
}

// ---- K1 with empty-space skipping -----------------------------------------------------------------------------
// Same results as k_forward, probe for probe.  Each ray's sample positions lie within eps_q (voxel units, bounded
// below) of the straight line  Q(k) = QA + k*QB  (voxel units), so:
//   * vol.clearance[cell] = C > 0 means: from ANY point of that 8^3-voxel macro cell one can move C voxels (L-inf) and
//     still be in macro cells that are empty and entirely inside the volume (Chebyshev distance field, built at upload,
//     0.25 voxel already subtracted).  Hence the n = floor(C / max|QB|) next samples, and the current one, are
//     in-bounds misses: they are counted (n_inb) and skipped without being evaluated.
//   * wherever the clearance is 0 (next to an occupied macro cell, next to the volume boundary, outside it) the
//     sample is evaluated exactly as in k_forward.
// Error budget: |float sample - line| <= eps_p = 16*2^-24 * (|m0|X + |m1|Y + |m2|Z + |m3|) per axis (5 roundings of the
// reference evaluation + <= 8 of the line's own float evaluation, doubled), eps_q = eps_p/delta.  A view whose eps_q
// exceeds 0.1 voxel on any axis does not skip at all (skip_ok = false).  2*eps_q <= 0.2 < the 0.25 voxel margin.
constexpr int SKIP_THREADS = 128;     // 4 warps: 2 across x 2 down, each warp an 8x4 pixel tile
constexpr int SKIP_TILE_W = 16, SKIP_TILE_H = 8;

template <int MODE, int FMT>
__global__ void __launch_bounds__(SKIP_THREADS, 8) k_forward_skip(const FwdArgs a) {
    __shared__ unsigned long long s_cnt[5];
    const int view = blockIdx.z + a.view0;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ci = blockIdx.x * SKIP_TILE_W + (warp & 1) * 8 + (lane & 7);
    const int ri = blockIdx.y * SKIP_TILE_H + (warp >> 1) * 4 + (lane >> 3);
    const bool active = ci < a.Wc && ri < a.Hc;
    if (threadIdx.x < 5) s_cnt[threadIdx.x] = 0ull;
    __syncthreads();

    const float* P = a.poses + 12 * (size_t)view;
    const float m00 = __ldg(P + 0), m01 = __ldg(P + 1), m02 = __ldg(P + 2), m03 = __ldg(P + 3);
    const float m10 = __ldg(P + 4), m11 = __ldg(P + 5), m12 = __ldg(P + 6), m13 = __ldg(P + 7);
    const float m20 = __ldg(P + 8), m21 = __ldg(P + 9), m22 = __ldg(P + 10), m23 = __ldg(P + 11);
    const VolDev& v = a.vol;
    const float lo0 = v.lo[0], lo1 = v.lo[1], lo2 = v.lo[2], hi0 = v.hi[0], hi1 = v.hi[1], hi2 = v.hi[2];
    const float in0 = v.inv32[0], in1 = v.inv32[1], in2 = v.inv32[2], cc0 = v.c32[0], cc1 = v.c32[1], cc2 = v.c32[2];
    const float er0 = v.err32[0], er1 = v.err32[1], er2 = v.err32[2];
    const unsigned pny = (unsigned)v.pdim[1], pnz = (unsigned)v.pdim[2];
    const unsigned mdx = (unsigned)v.mdim[0], mdy = (unsigned)v.mdim[1], mdz = (unsigned)v.mdim[2];
    const unsigned* __restrict__ gbits = v.bits;
    const unsigned char* __restrict__ gbytes = v.bytes;
    const float* __restrict__ clr = a.clearance;

    // the ray as a line in voxel units (approximate; only steers skipping)
    const int cic = active ? ci : 0, ric = active ? ri : 0;
    const float dcx = __ldg(a.dcx + cic), dcy = __ldg(a.dcy + ric);
    const float g0 = fmaf(m00, dcx, fmaf(m01, dcy, m02)), g1 = fmaf(m10, dcx, fmaf(m11, dcy, m12)), g2 = fmaf(m20, dcx, fmaf(m21, dcy, m22));
    const float z0m = (float)a.z0 * 0.001f, zdm = (float)a.zdelta * 0.001f;
    const float qa0 = fmaf(fmaf(z0m, g0, m03), in0, cc0), qa1 = fmaf(fmaf(z0m, g1, m13), in1, cc1), qa2 = fmaf(fmaf(z0m, g2, m23), in2, cc2);
    const float qb0 = zdm * g0 * in0, qb1 = zdm * g1 * in1, qb2 = zdm * g2 * in2;
    const float qbmax = fmaxf(fabsf(qb0), fmaxf(fabsf(qb1), fabsf(qb2)));
    const float rq = 1.0f / fmaxf(qbmax, 1e-3f);   // <= 1000: keeps (d-1.25)*rq inside the shifter's integer range
    // per-view error bound in voxel units (uniform over the block)
    const float kEps = 9.5367431640625e-07f;   // 16 * 2^-24
    const float e0 = kEps * (fabsf(m00) * a.dcx_max + fabsf(m01) * a.dcy_max + fabsf(m02) + fabsf(m03) + fabsf((float)v.vmin[0])) * fabsf(in0);
    const float e1 = kEps * (fabsf(m10) * a.dcx_max + fabsf(m11) * a.dcy_max + fabsf(m12) + fabsf(m13) + fabsf((float)v.vmin[1])) * fabsf(in1);
    const float e2 = kEps * (fabsf(m20) * a.dcx_max + fabsf(m21) * a.dcy_max + fabsf(m22) + fabsf(m23) + fabsf((float)v.vmin[2])) * fabsf(in2);
    const bool skip_ok = clr != nullptr && fmaxf(e0, fmaxf(e1, e2)) <= 0.1f;   // NaN poses compare false: no skipping

    int hit_k = -1, hx = 0, hy = 0, hz = 0;
    float hpx = 0.f, hpy = 0.f, hpz = 0.f;
    unsigned n_inb = 0, n_exact = 0, n_f64 = 0, n_skip = 0;
    int k = 0;
    float kf = 0.0f;
    const int S = a.S;
    const float* __restrict__ xt = a.xtab + cic;
    const float* __restrict__ yt = a.ytab + ric;
    const float* __restrict__ zt = a.ztab;
    const float kM = 12582912.0f;
    const float fd0 = v.ext[0], fd1 = v.ext[1], fd2 = v.ext[2];   // volume extent in voxel units (>= dim, < dim+1)
    unsigned iter = 0;
    int oob_wait = 0;
    if (active) {
        while (k < S) {
            if (MODE == 4 && (iter++ & 15u) == 0u) {   // rayTraceAndGetMinimum: planes behind the current minimum cannot matter
                const int cur = *((volatile int*)(a.min_depth + view));
                if (a.z0 + k * a.zdelta > cur) break;
            }
            if (FMT == 0 && skip_ok) {
                // bit grid: macro cell of the line point (floor(q/8) through the round-down shifter; no F2I)
                const float q0 = fmaf(kf, qb0, qa0), q1 = fmaf(kf, qb1, qa1), q2 = fmaf(kf, qb2, qa2);
                const unsigned mx = (unsigned)(__float_as_int(__fadd_rd(q0 * 0.125f, kM)) - 0x4B400000);
                const unsigned my = (unsigned)(__float_as_int(__fadd_rd(q1 * 0.125f, kM)) - 0x4B400000);
                const unsigned mz = (unsigned)(__float_as_int(__fadd_rd(q2 * 0.125f, kM)) - 0x4B400000);
                if (mx < mdx && my < mdy && mz < mdz) {
                    const float c = __ldg(clr + ((mx * mdy + my) * mdz + mz));
                    if (c > 0.0f) {
                        int n = min(__float2int_rz(c * rq), S - k - 1);   // samples k .. k+n are provably in-bounds misses
                        n_skip += (unsigned)(n + 1); n_inb += (unsigned)(n + 1);
                        k += n + 1; kf += (float)(n + 1);
                        continue;
                    }
                }
            }
            // ---- exact evaluation of sample k (identical to k_forward) ----
            const float xf = __ldg(xt + (size_t)k * a.Wc), yf = __ldg(yt + (size_t)k * a.Hc), zf = __ldg(zt + k);
            const float px = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m00, xf), __fmul_rn(m01, yf)), __fmul_rn(m02, zf)), m03);
            const float py = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m10, xf), __fmul_rn(m11, yf)), __fmul_rn(m12, zf)), m13);
            const float pz = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m20, xf), __fmul_rn(m21, yf)), __fmul_rn(m22, zf)), m23);
            if (!(px > lo0 && px < hi0 && py > lo1 && py < hi1 && pz > lo2 && pz < hi2)) {
                // validPoints failed: the reference just moves on.  Once per excursion, use the line to jump over the samples
                // that are provably outside (some axis of the line is > 0.25 voxel beyond the volume, eps_q <= 0.1).
                int kn = k + 1;
                if (skip_ok && oob_wait == 0) {
                    // real-valued k interval where the line is inside the volume inflated by 0.25 voxel on every side
                    float t0 = -1e30f, t1 = 1e30f;
                    const float qa[3] = {qa0, qa1, qa2}, qb[3] = {qb0, qb1, qb2}, fd[3] = {fd0, fd1, fd2};
#pragma unroll
                    for (int ax = 0; ax < 3; ax++) {
                        if (fabsf(qb[ax]) > 1e-12f) {
                            const float r = 1.0f / qb[ax];
                            const float ta = (-0.25f - qa[ax]) * r, tb = (fd[ax] + 0.25f - qa[ax]) * r;
                            t0 = fmaxf(t0, fminf(ta, tb)); t1 = fminf(t1, fmaxf(ta, tb));
                        } else if (qa[ax] < -0.25f || qa[ax] > fd[ax] + 0.25f) { t0 = 1e30f; }
                    }
                    // t0,t1 carry a relative error of a few ulps: move the entry 1 sample earlier and the exit 1 later
                    if (!(t0 <= t1) || t1 + 1.0f < kf) kn = S;                               // never (again) inside: all remaining samples fail validPoints
                    else if (t0 - 1.0f > kf + 1.0f) kn = min(S, max(k + 1, __float2int_rd(t0 - 1.0f)));
                    if (kn == k + 1) oob_wait = 4;                                           // inconclusive (grazing the boundary): retry a few samples later
                } else if (oob_wait > 0) oob_wait--;
                kf += (float)(kn - k); k = kn;
                continue;
            }
            oob_wait = 0;
            n_inb++;
            bool unsafe = false;
            int ix = voxel_index_f32(px, in0, cc0, er0, unsafe);
            int iy = voxel_index_f32(py, in1, cc1, er1, unsafe);
            int iz = voxel_index_f32(pz, in2, cc2, er2, unsafe);
            if (unsafe) {
                n_f64++;
                ix = voxel_index(px, v.vmin[0], v.delta[0], v.inv[0], v.c0[0], v.eps[0], n_exact);
                iy = voxel_index(py, v.vmin[1], v.delta[1], v.inv[1], v.c0[1], v.eps[1], n_exact);
                iz = voxel_index(pz, v.vmin[2], v.delta[2], v.inv[2], v.c0[2], v.eps[2], n_exact);
            }
            const unsigned idx = ((unsigned)ix * pny + (unsigned)iy) * pnz + (unsigned)iz;
            if (FMT == 0) {
                if ((__ldg(gbits + (idx >> 5)) >> (idx & 31)) & 1u) {
                    hit_k = k; hx = ix; hy = iy; hz = iz; hpx = px; hpy = py; hpz = pz;
                    k++;
                    break;
                }
                k++; kf += 1.0f;
            } else {
                // distance bytes (dmf_distance.cuh) with the border distance folded in: 0 = occupied; d >= 2 => every voxel within
                // d-1 (L-inf) of this one is an empty interior voxel, so the samples whose line offset stays within d - 1.25 voxels are in-bounds misses
                const unsigned d = byte_with_border(v, __ldg(gbytes + idx), ix, iy, iz);
                if (d == 0u) {
                    hit_k = k; hx = ix; hy = iy; hz = iz; hpx = px; hpy = py; hpz = pz;
                    k++;
                    break;
                }
                int n = 0;
                if (skip_ok && d >= 2u) {
                    const float df = __int_as_float(0x4B000000 | (int)d) - 8388608.0f;        // (float)d without I2F
                    n = min(__float_as_int(__fadd_rd((df - 1.25f) * rq, kM)) - 0x4B400000, S - k - 1);
                    n_skip += (unsigned)n; n_inb += (unsigned)n;
                }
                k += n + 1; kf += (float)(n + 1);
            }
        }
    }
    const unsigned n_samples = active ? (unsigned)min(k, S) : 0u;
    forward_epilogue<MODE>(a, s_cnt, view, ci, ri, active, hit_k, hx, hy, hz, hpx, hpy, hpz, m03, m13, m23, n_samples, n_inb, n_exact, n_f64, n_skip);
    if (a.pub.enabled) publish_view_row(a.pub, reinterpret_cast<u64*>(a.vis + (size_t)(unsigned)view * a.vis_stride32), view, gridDim.x * gridDim.y);
}

// Per view: how many leading probes every ray may skip.  All rays leave from the camera centre t; probe k of any ray is
// within z_k * G voxels (L-inf) of it, G = max_i (|m_i0|*max|dcx| + |m_i1|*max|dcy| + |m_i2|) / delta_i.  If the camera
// sits in a voxel with distance byte d (border distance folded in), everything within d-1 voxels of that voxel is empty and interior, so the probes
// with z_k * G <= d - 1.25 are in-bounds misses for every ray (0.25 = 2 * eps_q + slack, as in k_forward_dist).
__global__ void k_view_start(const FwdArgs a, int n_views, int* __restrict__ kstart) {
    const int view = blockIdx.x * blockDim.x + threadIdx.x;
    if (view >= n_views) return;
    const VolDev& v = a.vol;
    const float* P = a.poses + 12 * (size_t)view;
    const float m[3][4] = {{P[0], P[1], P[2], P[3]}, {P[4], P[5], P[6], P[7]}, {P[8], P[9], P[10], P[11]}};
    int k0 = 0;
    const float tx = m[0][3], ty = m[1][3], tz = m[2][3];
    const float kEps = 9.5367431640625e-07f;   // 16 * 2^-24
    float emax = 0.f, G = 0.f;
    bool finite = true;                                   // fmaxf drops NaN operands, so non-finite poses are tested explicitly
    for (int i = 0; i < 3; i++) {
        const float reach = fabsf(m[i][0]) * a.dcx_max + fabsf(m[i][1]) * a.dcy_max + fabsf(m[i][2]);
        const float e = kEps * (reach + fabsf(m[i][3]) + fabsf((float)v.vmin[i])) * fabsf(v.inv32[i]);
        finite = finite && (e <= 3.0e38f);                // false for NaN and +inf
        emax = fmaxf(emax, e);
        G = fmaxf(G, reach * fabsf(v.inv32[i]) * 1.00001f);
    }
    if (!finite) emax = __int_as_float(0x7fc00000);       // NaN: every "<= bound" test downstream fails
    if (emax <= 0.1f && in_bounds(v, tx, ty, tz) && v.bytes != nullptr) {       // NaN poses fail the comparisons
        unsigned ne = 0;
        const int ix = voxel_index(tx, v.vmin[0], v.delta[0], v.inv[0], v.c0[0], v.eps[0], ne);
        const int iy = voxel_index(ty, v.vmin[1], v.delta[1], v.inv[1], v.c0[1], v.eps[1], ne);
        const int iz = voxel_index(tz, v.vmin[2], v.delta[2], v.inv[2], v.c0[2], v.eps[2], ne);
        if (coords_valid(v, ix, iy, iz)) {
            const float d = (float)byte_with_border(v, v.bytes[linear_index(v, ix, iy, iz)], ix, iy, iz);
            // z_k = (z0 + k*zdelta) mm;  need z_k * 0.001 * G <= d - 1.25  for all k < k0
            const float zmax_mm = (d - 1.25f) / fmaxf(G, 1e-6f) * 1000.0f * 0.9999f;
            if (zmax_mm >= (float)a.z0) k0 = min(a.S, (int)floorf((zmax_mm - (float)a.z0) / (float)a.zdelta) + 1);
        }
    }
    // -1: the view's error bound is too large (or the pose is not finite): k_forward_line must evaluate every sample exactly
    const int ks = (emax <= 0.1f && v.bytes != nullptr) ? max(k0, 0) : -1;
    kstart[view] = ks;
    if (a.veps) a.veps[view] = emax;
    if (a.viewrec) {
        float4* r = a.viewrec + 4 * (size_t)view;
        r[0] = make_float4(m[0][0], m[0][1], m[0][2], m[0][3]); r[1] = make_float4(m[1][0], m[1][1], m[1][2], m[1][3]); r[2] = make_float4(m[2][0], m[2][1], m[2][2], m[2][3]);
        r[3] = make_float4(emax + 0.0009765625f, __int_as_float(ks), 0.f, 0.f);          // eps_q of the view + 2^-10 voxel of slack
    }
}

// ---- K1 on distance bytes: k_forward_dist ---------------------------------------------------------------------
// DMF_GRID_BYTE march.  Every exactly evaluated probe reads its voxel's distance byte d (dmf_distance.cuh, border distance folded in): 0 = hit;
// d >= 2 proves the next floor((d - 1.25) / max|QB|) probes are in-bounds misses (same error budget as k_forward_skip:
// eps_q <= 0.1 voxel per probe, 0.25 voxel margin), so they are counted and skipped.  Two consecutive probes (k, k+1) are
// evaluated per iteration so that their table and grid loads are in flight together; the second is used only if the
// first neither hits nor skips.  EXACT: every axis has err32 == 0 (power-of-two voxel size, vmin == 0): the float
// quotient is exact and floor() is one round-down add.
template <bool EXACT>
__device__ __forceinline__ unsigned probe_index(const VolDev& v, float px, float py, float pz, float in0, float in1, float in2,
                                                float cc0, float cc1, float cc2, float er0, float er1, float er2, unsigned pny, unsigned pnz,
                                                int& ix, int& iy, int& iz, unsigned& n_f64, unsigned& n_exact) {
    const float kM = 12582912.0f;
    if (EXACT) {
        ix = __float_as_int(__fadd_rd(fmaf(px, in0, cc0), kM)) - 0x4B400000;
        iy = __float_as_int(__fadd_rd(fmaf(py, in1, cc1), kM)) - 0x4B400000;
        iz = __float_as_int(__fadd_rd(fmaf(pz, in2, cc2), kM)) - 0x4B400000;
    } else {
        bool unsafe = false;
        ix = voxel_index_f32(px, in0, cc0, er0, unsafe);
        iy = voxel_index_f32(py, in1, cc1, er1, unsafe);
        iz = voxel_index_f32(pz, in2, cc2, er2, unsafe);
        if (unsafe) {
            n_f64++;
            ix = voxel_index(px, v.vmin[0], v.delta[0], v.inv[0], v.c0[0], v.eps[0], n_exact);
            iy = voxel_index(py, v.vmin[1], v.delta[1], v.inv[1], v.c0[1], v.eps[1], n_exact);
            iz = voxel_index(pz, v.vmin[2], v.delta[2], v.inv[2], v.c0[2], v.eps[2], n_exact);
        }
    }
    return ((unsigned)ix * pny + (unsigned)iy) * pnz + (unsigned)iz;
}

template <int MODE, bool EXACT>
__global__ void __launch_bounds__(SKIP_THREADS, 10) k_forward_dist(const FwdArgs a) {
    __shared__ unsigned long long s_cnt[5];
    const int view = blockIdx.z + a.view0;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ci = blockIdx.x * SKIP_TILE_W + (warp & 1) * 8 + (lane & 7);
    const int ri = blockIdx.y * SKIP_TILE_H + (warp >> 1) * 4 + (lane >> 3);
    const bool active = ci < a.Wc && ri < a.Hc;
    if (threadIdx.x < 5) s_cnt[threadIdx.x] = 0ull;
    __syncthreads();

    const float* P = a.poses + 12 * (size_t)view;
    const float m00 = __ldg(P + 0), m01 = __ldg(P + 1), m02 = __ldg(P + 2), m03 = __ldg(P + 3);
    const float m10 = __ldg(P + 4), m11 = __ldg(P + 5), m12 = __ldg(P + 6), m13 = __ldg(P + 7);
    const float m20 = __ldg(P + 8), m21 = __ldg(P + 9), m22 = __ldg(P + 10), m23 = __ldg(P + 11);
    const VolDev& v = a.vol;
    const float lo0 = v.lo[0], lo1 = v.lo[1], lo2 = v.lo[2], hi0 = v.hi[0], hi1 = v.hi[1], hi2 = v.hi[2];
    const float in0 = v.inv32[0], in1 = v.inv32[1], in2 = v.inv32[2], cc0 = v.c32[0], cc1 = v.c32[1], cc2 = v.c32[2];
    const float er0 = v.err32[0], er1 = v.err32[1], er2 = v.err32[2];
    const unsigned pny = (unsigned)v.pdim[1], pnz = (unsigned)v.pdim[2];
    const unsigned char* __restrict__ gbytes = v.bytes;

    // slope of the ray in voxel units per probe (approximate; only steers skipping) and the per-view error bound
    const int cic = active ? ci : 0, ric = active ? ri : 0;
    const float dcx = __ldg(a.dcx + cic), dcy = __ldg(a.dcy + ric);
    const float g0 = fmaf(m00, dcx, fmaf(m01, dcy, m02)), g1 = fmaf(m10, dcx, fmaf(m11, dcy, m12)), g2 = fmaf(m20, dcx, fmaf(m21, dcy, m22));
    const float zdm = (float)a.zdelta * 0.001f;
    const float qbmax = fmaxf(fabsf(zdm * g0 * in0), fmaxf(fabsf(zdm * g1 * in1), fabsf(zdm * g2 * in2)));
    const float kEps = 9.5367431640625e-07f;   // 16 * 2^-24
    const float e0 = kEps * (fabsf(m00) * a.dcx_max + fabsf(m01) * a.dcy_max + fabsf(m02) + fabsf(m03) + fabsf((float)v.vmin[0])) * fabsf(in0);
    const float e1 = kEps * (fabsf(m10) * a.dcx_max + fabsf(m11) * a.dcy_max + fabsf(m12) + fabsf(m13) + fabsf((float)v.vmin[1])) * fabsf(in1);
    const float e2 = kEps * (fabsf(m20) * a.dcx_max + fabsf(m21) * a.dcy_max + fabsf(m22) + fabsf(m23) + fabsf((float)v.vmin[2])) * fabsf(in2);
    const bool skip_ok = fmaxf(e0, fmaxf(e1, e2)) <= 0.1f;   // NaN poses compare false: no skipping
    // probes that may be skipped after a probe with distance byte d:  n(d) = ((4d - 5) * rfix) >> 12 <= (d - 1.25) / qbmax.
    // rfix = floor(2^10 / qbmax) rounded down (<= 2^20 so the product stays below 2^31); 0 disables skipping.
    const int rfix = skip_ok ? __float2int_rd(fminf(__fdividef(1024.0f, fmaxf(qbmax, 1e-3f)) * 0.999999f, 1048576.0f)) : 0;

    int hit_k = -1, hx = 0, hy = 0, hz = 0;
    float hpx = 0.f, hpy = 0.f, hpz = 0.f;
    unsigned n_inb = 0, n_exact = 0, n_f64 = 0, n_skip = 0;
    const int S = a.S;
    int k = active ? min(max(__ldg(a.kstart + view), 0), S) : 0;      // leading probes that no ray of this view can hit (k_view_start)
    n_inb = (unsigned)k; n_skip = (unsigned)k;
    const float* xt = a.xtab + cic;
    const float* yt = a.ytab + ric;
    const float* __restrict__ zt = a.ztab;
    const unsigned uWc = (unsigned)a.Wc, uHc = (unsigned)a.Hc;
    // make the per-thread row pointers opaque register values: otherwise the compiler re-derives base + column from the
    // kernel parameters in every iteration (five 64-bit instructions per load instead of IMAD + IMAD.WIDE)
    asm volatile("" : "+l"(xt), "+l"(yt));
    unsigned iter = 0;
    int oob_wait = 0;
    if (active) {
        while (k < S) {
            if (MODE == 4 && (iter++ & 7u) == 0u) {   // rayTraceAndGetMinimum: planes behind the current minimum cannot matter
                const int cur = *((volatile int*)(a.min_depth + view));
                if (a.z0 + k * a.zdelta > cur) break;
            }
            // ---- probes A = k and B = k+1: loads first, so that both are in flight together ----
            const unsigned ka = (unsigned)k, kb = (unsigned)min(k + 1, S - 1);     // unsigned: one IMAD.WIDE.U32 per address
            const float xa = __ldg(xt + ka * uWc), ya = __ldg(yt + ka * uHc), za = __ldg(zt + ka);
            const float xb = __ldg(xt + kb * uWc), yb = __ldg(yt + kb * uHc), zb = __ldg(zt + kb);
            const float pxa = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m00, xa), __fmul_rn(m01, ya)), __fmul_rn(m02, za)), m03);
            const float pya = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m10, xa), __fmul_rn(m11, ya)), __fmul_rn(m12, za)), m13);
            const float pza = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m20, xa), __fmul_rn(m21, ya)), __fmul_rn(m22, za)), m23);
            if (!(pxa > lo0 && pxa < hi0 && pya > lo1 && pya < hi1 && pza > lo2 && pza < hi2)) {
                // validPoints failed: the reference just moves on.  Use the line to jump over the probes that are provably
                // outside (some axis of the line > 0.25 voxel beyond the volume, eps_q <= 0.1); retried every few probes.
                int kn = k + 1;
                if (skip_ok && oob_wait == 0) {
                    const float z0m = (float)a.z0 * 0.001f, kf = (float)k;
                    const float qa[3] = {fmaf(fmaf(z0m, g0, m03), in0, cc0), fmaf(fmaf(z0m, g1, m13), in1, cc1), fmaf(fmaf(z0m, g2, m23), in2, cc2)};
                    const float qb[3] = {zdm * g0 * in0, zdm * g1 * in1, zdm * g2 * in2};
                    float t0 = -1e30f, t1 = 1e30f;
#pragma unroll
                    for (int ax = 0; ax < 3; ax++) {
                        if (fabsf(qb[ax]) > 1e-12f) {
                            const float r = 1.0f / qb[ax];
                            const float ta = (-0.25f - qa[ax]) * r, tb = (v.ext[ax] + 0.25f - qa[ax]) * r;
                            t0 = fmaxf(t0, fminf(ta, tb)); t1 = fminf(t1, fmaxf(ta, tb));
                        } else if (qa[ax] < -0.25f || qa[ax] > v.ext[ax] + 0.25f) { t0 = 1e30f; }
                    }
                    if (!(t0 <= t1) || t1 + 1.0f < kf) kn = S;                               // never (again) inside the volume
                    else if (t0 - 1.0f > kf + 1.0f) kn = min(S, max(k + 1, __float2int_rd(t0 - 1.0f)));
                    if (kn == k + 1) oob_wait = 4;
                } else if (oob_wait > 0) oob_wait--;
                k = kn;
                continue;
            }
            oob_wait = 0;
            const float pxb = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m00, xb), __fmul_rn(m01, yb)), __fmul_rn(m02, zb)), m03);
            const float pyb = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m10, xb), __fmul_rn(m11, yb)), __fmul_rn(m12, zb)), m13);
            const float pzb = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m20, xb), __fmul_rn(m21, yb)), __fmul_rn(m22, zb)), m23);
            const bool inb_b = (k + 1 < S) && pxb > lo0 && pxb < hi0 && pyb > lo1 && pyb < hi1 && pzb > lo2 && pzb < hi2;
            int ixa, iya, iza, ixb = 0, iyb = 0, izb = 0;
            const unsigned idxa = probe_index<EXACT>(v, pxa, pya, pza, in0, in1, in2, cc0, cc1, cc2, er0, er1, er2, pny, pnz, ixa, iya, iza, n_f64, n_exact);
            const unsigned da = byte_with_border(v, __ldg(gbytes + idxa), ixa, iya, iza);
            unsigned db = 1u;
            if (inb_b) {
                const unsigned idxb = probe_index<EXACT>(v, pxb, pyb, pzb, in0, in1, in2, cc0, cc1, cc2, er0, er1, er2, pny, pnz, ixb, iyb, izb, n_f64, n_exact);
                db = byte_with_border(v, __ldg(gbytes + idxb), ixb, iyb, izb);
            }
            n_inb++;
            if (da == 0u) { hit_k = k; hx = ixa; hy = iya; hz = iza; hpx = pxa; hpy = pya; hpz = pza; k++; break; }
            const int na = min(((int)(4u * da) - 5) * rfix >> 12, S - k - 1);       // da >= 1: (4*da-5) >= -1, and -1*rfix>>12 is -1 or 0
            if (na >= 1 || !inb_b) {
                const int n = max(na, 0);
                n_inb += (unsigned)n; n_skip += (unsigned)n;
                k += n + 1;                                                         // (an out-of-bounds B is re-examined as the next A)
                continue;
            }
            // A neither hit nor skipped: B is the next probe of this ray
            n_inb++;
            if (db == 0u) { hit_k = k + 1; hx = ixb; hy = iyb; hz = izb; hpx = pxb; hpy = pyb; hpz = pzb; k += 2; break; }
            const int nb = max(min(((int)(4u * db) - 5) * rfix >> 12, S - k - 2), 0);
            n_inb += (unsigned)nb; n_skip += (unsigned)nb;
            k += nb + 2;
        }
    }
    const unsigned n_samples = active ? (unsigned)min(k, S) : 0u;
    forward_epilogue<MODE>(a, s_cnt, view, ci, ri, active, hit_k, hx, hy, hz, hpx, hpy, hpz, m03, m13, m23, n_samples, n_inb, n_exact, n_f64, n_skip);
    if (a.pub.enabled) publish_view_row(a.pub, reinterpret_cast<u64*>(a.vis + (size_t)(unsigned)view * a.vis_stride32), view, gridDim.x * gridDim.y);
}

// ---- per-tile prologue of the line march: k_tile_start ----------------------------------------------------------------
// The 16 rays of a 4x4 tile of the lattice leave the camera within a fraction of a degree of each other.  What k_forward_line would
// make each of them work out for itself is done here once per tile, by one thread (1/16 of the rays):
//
// (1) The sample intervals.  With the tile's centre line Qc(k) = QAc + k * QBc (voxel units), every ray's line differs from it by
//     z_k * M * (ddx, ddy, 0) / delta exactly, |ddx| <= hw_x, |ddy| <= hw_y (half the tile's extent in (c-cx)/fx, (r-cy)/fy), so on
//     axis i every ray's line point lies in Qc_i(k) +- (rho_i(k) + eps),  rho_i(k) = A_i * (z0 + k * zdelta),
//     A_i = |1/delta_i| * (|m_i0| hw_x + |m_i1| hw_y), eps = 2 * eps_q(view) for the float evaluation of the two lines.
//     "every ray's line is >= 0.25 voxel inside on axis i" and "some ray's line may be < 0.25 voxel outside" are then linear
//     inequalities in k -- (QBc_i -+ A_i zdelta) k >= / <= const -- and intersecting them gives [kin, kout] (inner, a subset of
//     every ray's own interval) and [k_begin, s_end) (outer, a superset), with the guard g(t) of k_forward_line.  Slopes too
//     small to divide by make the inner interval empty and leave the outer one unconstrained.
// (2) The cone pre-march.  Every sample k' of every ray of the tile lies within rho(k') = max_i rho_i(k') + 3 eps_q(view) of Qc(k')
//     (sample vs its line, its line in float, the centre line in float: eps_q each).  A probe at k reads the byte d of the voxel V
//     of Qc(k): every grid voxel within d-1 of V is empty, and a point within R < d-1 of a point of V is in such a voxel.  Samples
//     k .. k+n of all 16 rays are within rho(k) + n * (max|QBc| + A * zdelta) of Qc(k), so they are misses for
//     n = floor((d - 1 - 2^-8 - rho(k)) / (max|QBc| + A * zdelta)).  The march runs inside [kin, kout] (where Qc(k) is inside the grid),
//     from max(kin, kstart of the view) to the first probe that cannot advance: every sample k' in [kin, kt) of every ray is a miss.
//     Without it a ray spends ~7.3 dependent probes on the bench sweep, half of them far from any surface (the ones that miss L1
//     and L2); with it 2.8.
// Views flagged kstart = -1 (eps_q > 0.1, non-finite poses) get the record "evaluate every sample exactly".
constexpr int TILE_RAYS = 4;
#ifndef DMF_TILE_MIN_BLOCKS
#define DMF_TILE_MIN_BLOCKS 16                    // k_tile_start is a chain of dependent far-field probes: occupancy is what hides them
#endif
__device__ __forceinline__ float rcp_approx(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__host__ __device__ __forceinline__ u64 tile_pack(int k_begin, int s_end, int kin, int kout, int kt) {      // all in [0, 4095]; kout stored + 1
    return (u64)(unsigned)k_begin | ((u64)(unsigned)s_end << 12) | ((u64)(unsigned)kin << 24) | ((u64)(unsigned)(kout + 1) << 36) | ((u64)(unsigned)kt << 48);
}
// s * k >= r  folded into [lo, hi]; INNER: the interval must stay a subset (undecidable => empty), else a superset (=> unconstrained)
template <bool INNER>
__device__ __forceinline__ void tile_constrain(float& lo, float& hi, float s, float r) {      // branch-free: 12 of these per tile
    const bool div = fabsf(s) > 1e-9f;
    const float t = r * rcp_approx(div ? s : 1.0f);                        // relative error < 2^-21: inside tile_guard()
    lo = (div && s > 0.0f) ? fmaxf(lo, t) : lo;
    hi = (div && s < 0.0f) ? fminf(hi, t) : hi;
    // |s k| <= 1e-9 * 4096: the sign of r decides; in doubt, see above
    lo = (!div && (INNER ? !(r < -1e-3f) : (r > 1e-3f))) ? 3e30f : lo;
}
__device__ __forceinline__ float tile_guard(float t) { return fmaf(fabsf(t), 3.814697265625e-06f, 0.015625f); }   // 2^-6 + |t| * 2^-18 samples

__global__ void __launch_bounds__(128, DMF_TILE_MIN_BLOCKS) k_tile_start(const FwdArgs a, u64* __restrict__ tile_rec) {
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (unsigned)a.tiles_per_view) return;
    const unsigned view = blockIdx.y;
    const VolDev& v = a.vol;
    const float4* sp = a.viewrec + 4u * view;
    const float4 r0 = sp[0], r1 = sp[1], r2 = sp[2], r3 = sp[3];
    const int ks = __float_as_int(r3.y);
    const int S = a.S;
    u64 rec = tile_pack(0, S, S, -1, 0);                                             // every sample exactly
    if (ks >= 0) {
        const int tx = (int)(t % (unsigned)a.tiles_x), ty = (int)(t / (unsigned)a.tiles_x);
        const int c0 = tx * TILE_RAYS, c1 = min(c0 + TILE_RAYS - 1, a.Wc - 1), q0 = ty * TILE_RAYS, q1 = min(q0 + TILE_RAYS - 1, a.Hc - 1);
        const float xl = __ldg(a.dcx + c0), xh = __ldg(a.dcx + c1), yl = __ldg(a.dcy + q0), yh = __ldg(a.dcy + q1);
        const float xc = 0.5f * (xl + xh), yc = 0.5f * (yl + yh);
        const float hwx = fmaxf(fabsf(xl - xc), fabsf(xh - xc)) * 1.00001f, hwy = fmaxf(fabsf(yl - yc), fabsf(yh - yc)) * 1.00001f;
        const float m0[3] = {r0.x, r1.x, r2.x}, m1[3] = {r0.y, r1.y, r2.y}, m2[3] = {r0.z, r1.z, r2.z}, m3[3] = {r0.w, r1.w, r2.w};
        float qa[3], qb[3], A3[3];
#pragma unroll
        for (int i = 0; i < 3; i++) {
            const float g = fmaf(m0[i], xc, fmaf(m1[i], yc, m2[i]));
            qa[i] = fmaf(fmaf(a.z0m, g, m3[i]), v.inv32[i], v.c32[i]);
            qb[i] = a.zdm * g * v.inv32[i];
            A3[i] = fabsf(v.inv32[i]) * (fabsf(m0[i]) * hwx + fabsf(m1[i]) * hwy) * 1.0001f;
        }
        // ---- (1) sample intervals, conservative for the 16 rays ----
        const float eps = 2.0f * r3.x;                                               // 2 * (eps_q(view) + 2^-10)
        float ti0 = -1e30f, ti1 = 1e30f, to0 = -1e30f, to1 = 1e30f;
#pragma unroll
        for (int i = 0; i < 3; i++) {
            const float as = A3[i] * a.zdm, bs = A3[i] * a.z0m + eps;                // rho_i(k) + eps = as * k + bs
            tile_constrain<true>(ti0, ti1, qb[i] - as, 0.25f + bs - qa[i]);                       //  Qc - rho - eps >= 0.25
            tile_constrain<true>(ti0, ti1, -(qb[i] + as), -(v.ext[i] - 0.25f - bs - qa[i]));      //  Qc + rho + eps <= ext - 0.25
            tile_constrain<false>(to0, to1, qb[i] + as, -0.25f - bs - qa[i]);                     //  Qc + rho + eps >= -0.25
            tile_constrain<false>(to0, to1, -(qb[i] - as), -(v.ext[i] + 0.25f + bs - qa[i]));     //  Qc - rho - eps <= ext + 0.25
        }
        const float Sf = a.Sf;
        int k_begin = S, s_end = S, kin = S, kout = -1;
        if (to0 <= to1) {
            k_begin = (int)fminf(fmaxf(ceilf(to0 - tile_guard(to0)), 0.0f), Sf);
            s_end = (int)fminf(fmaxf(floorf(to1 + tile_guard(to1)) + 1.0f, 0.0f), Sf);
            if (ti0 <= ti1) {
                kin = (int)fminf(fmaxf(ceilf(ti0 + tile_guard(ti0)), 0.0f), Sf);
                kout = (int)fminf(fmaxf(floorf(ti1 - tile_guard(ti1)), -1.0f), Sf - 1.0f);
            }
        }
        // ---- (2) cone pre-march, inside [kin, kout] only: there every ray's line -- and so the centre line -- is >= 0.25 voxel inside the
        // volume, the probed voxel needs no clamping, and k_forward_line uses the result nowhere else (the samples of [kin, kstart) are
        // misses by k_view_start's proof) ----
        const float A = fmaxf(A3[0], fmaxf(A3[1], A3[2]));
        const float grow = fmaxf(fabsf(qb[0]), fmaxf(fabsf(qb[1]), fabsf(qb[2]))) * 1.0001f + A * a.zdm * 1.0001f;     // growth of the bound per sample
        const float rgrow = 0.9999f / fmaxf(grow, 1e-3f);
        const float rc0 = 1.0f + 0.00390625f + 3.0f * r3.x + A * a.z0m, c1n = -(A * a.zdm);   // room(k) = d - rc0 + c1n * k;  1 + 2^-8 + 3 * (eps_q(view) + 2^-10) + rho(0)
        const float kM = 12582912.0f, koutf = (float)kout;
        const unsigned pnz = (unsigned)v.pdim[2];
        float kf = (float)max(min(ks, S), kin);
        for (int it = 0; it < 64 && kf <= koutf; it++) {
            const unsigned bx = (unsigned)__float_as_int(__fadd_rd(fmaf(kf, qb[0], qa[0]), kM));
            const unsigned by = (unsigned)__float_as_int(__fadd_rd(fmaf(kf, qb[1], qa[1]), kM));
            const unsigned bz = (unsigned)__float_as_int(__fadd_rd(fmaf(kf, qb[2], qa[2]), kM));
            unsigned lidx = bx * a.pnyz + (by * pnz + (bz - a.bias));
            DMF_CHECK_IDX(lidx, v.n_cells, a.counters);
            const unsigned d = __ldg(v.bytes + min(lidx, a.last));                   // (the clamp is a seat belt, never active)
            const float room = fmaf(c1n, kf, (__int_as_float(0x4B000000 | (int)d) - 8388608.0f) - rc0);
            if (!(room >= 0.0f)) break;                                              // (NaN ends the march too)
            kf += (__fadd_rd(room * rgrow, kM) - kM) + 1.0f;                         // floor(room / growth) more samples, and this one
        }
        rec = tile_pack(k_begin, s_end, kin, kout, (int)fminf(kf, Sf));
    }
    tile_rec[view * (unsigned)a.tiles_per_view + t] = rec;
}

// ---- K1 on distance bytes, line-first: k_forward_line ------------------------------------------------------------
// Same results as k_forward / k_forward_dist, probe for probe.  The ray is FIRST followed as the straight line
// Q(k) = QA + k*QB in voxel units (3 FMAs, no table loads): the distance byte d of the line point's voxel decides.
//   d >= 2: every voxel within d-1 (L-inf) of that voxel is empty.  The reference's sample k lies within
//           eps_q <= 0.1 voxel of Q(k) (error budget above k_forward_skip), i.e. in that voxel or one next to it: a
//           miss WITHOUT being evaluated; so are the next floor((d - 1.25) / max|QB|) samples -- as far as they stay inside
//           [kin, kout], where the slab test proves them in bounds (the bytes hold the distance to the nearest OCCUPIED voxel and
//           know nothing of the boundary: the jump is cut at kout + 1).
//   d <= 1: next to an occupied voxel: the sample is evaluated exactly (tables, the reference's float expression, the
//           exact voxel index), as k_forward does -- unless the line point is far enough from every face of its voxel.
// The line is only consulted for k in [kin, kout], where it is >= 0.25 voxel inside the volume on every axis (slab test,
// once per 4x4-ray tile: k_tile_start) so its voxel is addressable; samples that are provably outside (line > 0.25 voxel
// beyond a face) are dropped without evaluation, and the thin bands in between are evaluated exactly.  Views whose eps_q
// exceeds 0.1 voxel (or NaN/inf poses; k_view_start flags them with kstart = -1) evaluate every sample exactly.
// The first line probe of a ray goes to the tile's kt: the cone pre-march of k_tile_start has proven everything before it empty.
// The kernel is issue- and latency-bound at once (ncu: > 80 % issue-active with all 64 warp slots of an SM in use, the top
// stall on the consumer of the distance-byte load), so the code below counts instructions AND registers:
//   * the pose is re-read from the view's 64-byte record with 16-byte loads where the (rare) exact path needs it, instead of
//     pinning 12 registers through the line loop (16 blocks of 128 threads at 32 registers; 12 blocks at 40 registers without
//     spills measured 8 % slower, 10 at 48: 18 %);
//   * "samples advanced after a probe with byte d" is floor(d * rq + c1) from the ray's own slope: 5 FMA-pipe instructions,
//     (float)d from the 0x4B000000 | d trick; d <= 1 yields 2^20, which ends the loop through the same comparison as running
//     past kout;
//   * reciprocals are MUFU approximations: they only steer conservative bounds.
#ifndef DMF_LINE_MIN_BLOCKS
#define DMF_LINE_MIN_BLOCKS 16
#endif
constexpr int LINE_MIN_BLOCKS = DMF_LINE_MIN_BLOCKS;
constexpr float LINE_EXACT_FLAG = 1048576.0f;      // 2^20, far above any sample index or jump (<= 254 * 1000 + 1)

// a 16-byte read of the per-view record at the point of use (volatile: the compiler must not keep the 12 pose floats live through the
// loops -- the kernel runs at 32 registers); the line sits in L1 and every lane reads the same address
__device__ __forceinline__ float4 ldg_f4_volatile(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}

// ---- carve mode on the line (DMF_FWD_CARVE) -----------------------------------------------------------------------
// Every in-bounds sample a ray visits, up to and including its first hit, sets the bit of its voxel in a.observed.  Which
// voxel that is has to be known for EVERY sample, so nothing can be skipped -- but almost nothing has to be evaluated the
// reference's way either: for k in [kin, kout] the sample is in bounds (the line is >= 0.25 voxel inside, eps_q <= 0.1) and
// lies within eps_q(view) of the line point Q(k); if Q(k) is at least e_safe = eps_q(view) + 2^-10 away from every face of
// its voxel, the sample is in that same voxel (the argument k_forward_line uses for its d == 1 probes).  Only samples whose
// line point is closer than that to a face, and those in the thin bands around the volume boundary, take the exact path
// (tables, the reference's float expression, validPoints, exact index).  Runs after the march, when hit_k is known.
__device__ __forceinline__ void observe_voxel(unsigned* __restrict__ obs, unsigned idx) {
    // read first: after the first few views of a sweep nearly every bit is already set and the word sits in L1/L2; a stale 0
    // only costs a redundant (idempotent) atomicOr
    unsigned* const w = obs + (idx >> 5);
    unsigned m;
    asm("shf.l.wrap.b32 %0, 0, 1, %1;" : "=r"(m) : "r"(idx));      // 1u << (idx & 31) in one instruction, opaque to strength "reduction"
    if (!(*w & m)) atomicOr(w, m);
}

// sample k of pixel (ci, ri) evaluated the reference's way (as k_forward does); marks its voxel if it passes validPoints
template <bool EXACT>
__device__ __forceinline__ void carve_exact_sample(const FwdArgs& a, const float4* sp, int ci, int ri, int k) {
    const VolDev& v = a.vol;
    const float xf = __ldg(a.xtab + ((unsigned)k * (unsigned)a.Wc + (unsigned)ci)), yf = __ldg(a.ytab + ((unsigned)k * (unsigned)a.Hc + (unsigned)ri));
    const float zf = __ldg(a.ztab + k);
    const float4 r0 = ldg_f4_volatile(sp), r1 = ldg_f4_volatile(sp + 1), r2 = ldg_f4_volatile(sp + 2);
    const float px = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(r0.x, xf), __fmul_rn(r0.y, yf)), __fmul_rn(r0.z, zf)), r0.w);
    const float py = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(r1.x, xf), __fmul_rn(r1.y, yf)), __fmul_rn(r1.z, zf)), r1.w);
    const float pz = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(r2.x, xf), __fmul_rn(r2.y, yf)), __fmul_rn(r2.z, zf)), r2.w);
    if (!(px > v.lo[0] && px < v.hi[0] && py > v.lo[1] && py < v.hi[1] && pz > v.lo[2] && pz < v.hi[2])) return;   // validPoints failed
    int ix, iy, iz;
    unsigned dummy0 = 0, dummy1 = 0;
    unsigned idx = probe_index<EXACT>(v, px, py, pz, v.inv32[0], v.inv32[1], v.inv32[2], v.c32[0], v.c32[1], v.c32[2],
                                      v.err32[0], v.err32[1], v.err32[2], (unsigned)v.pdim[1], (unsigned)v.pdim[2], ix, iy, iz, dummy0, dummy1);
    DMF_CHECK_IDX(idx, v.n_cells, a.counters);
    observe_voxel(a.observed, idx);
}

// The loop is issue-bound and, within that, bound by the half-rate integer (ALU) pipe -- a 12.20 fixed-point version of the
// line (integer add / mask / compare per axis) measured no faster than the float one because every one of its instructions
// went down that pipe (ncu: ALU 63 % busy, FMA 13 %).  So the line stays in float, on the full-rate FMA pipe:
//   t = Q(k) - 0.5;  s = t + 1.5*2^23 (round to nearest) -> low mantissa bits = rint(t) = n;  h = t - n in [-0.5, 0.5]
//   => Q(k) = n + 0.5 + h: the voxel is n and the distance to its nearer face is 0.5 - |h|;  safe <=> max|h| <= 0.5 - e_safe.
// Four FMA-pipe instructions per axis, then one 3-input max and one compare.  Samples that fail the face test are only
// remembered (one bit each) and evaluated exactly after their chunk of 32, so a warp runs the long exact path a couple of
// times per 32 samples instead of whenever any of its 32 lanes needs it.
#ifndef DMF_CARVE_MLP
#define DMF_CARVE_MLP 8
#endif
#ifndef DMF_CARVE_X2
#define DMF_CARVE_X2 1                            // 1: locate two samples per FFMA2 / FADD2 (sm_100 packed FP32)
#endif
#if DMF_CARVE_X2 && (DMF_CARVE_MLP % 2)
#error "DMF_CARVE_X2 needs an even DMF_CARVE_MLP"
#endif
#ifndef DMF_CARVE_SIGN
#define DMF_CARVE_SIGN 1                          // 1: carve_on_line_sign (face test collected as sign bits, OR tree for the missing bits, full groups only)
#endif
#ifndef DMF_CARVE_CLAMP
#define DMF_CARVE_CLAMP 0                         // carve_on_line_sign: 1 = clamp the located voxel index (never active, see the proof there; costs 3.5 %)
#endif
#if DMF_CARVE_SIGN && (!DMF_CARVE_X2 || (DMF_CARVE_MLP & (DMF_CARVE_MLP - 1)) || DMF_CARVE_MLP > 32)
#error "DMF_CARVE_SIGN needs DMF_CARVE_X2 and a power-of-two DMF_CARVE_MLP <= 32"
#endif
#ifndef DMF_CARVE_MIN_BLOCKS
#define DMF_CARVE_MIN_BLOCKS 8
#endif
constexpr int CARVE_MLP = DMF_CARVE_MLP;          // samples whose observed-word loads are in flight together
constexpr int CARVE_MIN_BLOCKS = DMF_CARVE_MIN_BLOCKS;

template <bool EXACT>
__device__ __forceinline__ void carve_on_line(const FwdArgs& a, const float4* sp, int ci, int ri, int k_first, int k_last, int kin, int kout,
                                              float qa0, float qa1, float qa2, float qb0, float qb1, float qb2, float esafe) {
    const VolDev& v = a.vol;
    const int b0 = max(kin, k_first), b1 = min(kout, k_last);
    if (b0 > b1) {                                           // the line is never safely inside within the visited range
        for (int k = k_first; k <= k_last; k++) carve_exact_sample<EXACT>(a, sp, ci, ri, k);
        return;
    }
    for (int k = k_first; k < b0; k++) carve_exact_sample<EXACT>(a, sp, ci, ri, k);          // entry band around the boundary
    {
        const float kM = 12582912.0f;                                                        // 1.5 * 2^23
        unsigned pnz = (unsigned)v.pdim[2], pnyz = (unsigned)v.pdim[1] * pnz;
        unsigned last = pnyz * (unsigned)v.pdim[0] - 1u;
        unsigned bias = 0x4B400000u * (pnyz + pnz + 1u);                                     // the three "- 0x4B400000" of the shifter, folded
        asm volatile("" : "+r"(pnz), "+r"(pnyz), "+r"(last), "+r"(bias));                    // registers, not re-derived from the constant bank per sample
        const float ta0 = qa0 - 0.5f, ta1 = qa1 - 0.5f, ta2 = qa2 - 0.5f;
        const float hmax = 0.5f - esafe;
#if DMF_CARVE_X2
        const f32x2 qb0p = f2_pack(qb0, qb0), qb1p = f2_pack(qb1, qb1), qb2p = f2_pack(qb2, qb2);
        const f32x2 ta0p = f2_pack(ta0, ta0), ta1p = f2_pack(ta1, ta1), ta2p = f2_pack(ta2, ta2), kMp = f2_pack(kM, kM);
#endif
        unsigned* const obs = a.observed;
        for (int kb = b0; kb <= b1; kb += 32) {
            const int n = min(32, b1 - kb + 1);
            unsigned unsafe = 0u;
            // CARVE_MLP samples at a time: all their voxels are located first, then all their observed words are loaded (back to
            // back, so the loads overlap), then tested.  With one load -> test -> branch per sample the loop was latency-bound
            // (ncu: 41 % of the stall samples on the instruction consuming the load, issue-active 56 %).
            for (int j = 0; j < n; j += CARVE_MLP) {
                unsigned idx[CARVE_MLP], word[CARVE_MLP];
                float hm[CARVE_MLP];                                                     // max |h| over the three axes
                const float kfb = (float)(kb + j);
#if DMF_CARVE_X2
                // two samples per instruction (FFMA2 / FADD2): the same fmaf / add / sub sequence as below, each half
                // rounded exactly like the scalar op, at half the issue slots
                const f32x2 kp = f2_pack(kfb, kfb + 1.0f);
#pragma unroll
                for (int u = 0; u < CARVE_MLP; u += 2) {
                    const f32x2 kf = f2_add(kp, f2_pack((float)u, (float)u));            // exact: small integers
                    const f32x2 t0 = f2_fma(kf, qb0p, ta0p), t1 = f2_fma(kf, qb1p, ta1p), t2 = f2_fma(kf, qb2p, ta2p);
                    const f32x2 s0 = f2_add(t0, kMp), s1 = f2_add(t1, kMp), s2 = f2_add(t2, kMp);
                    const f32x2 h0 = f2_sub(t0, f2_sub(s0, kMp)), h1 = f2_sub(t1, f2_sub(s1, kMp)), h2 = f2_sub(t2, f2_sub(s2, kMp));
                    float h0a, h0b, h1a, h1b, h2a, h2b, s0a, s0b, s1a, s1b, s2a, s2b;
                    f2_unpack(h0, h0a, h0b); f2_unpack(h1, h1a, h1b); f2_unpack(h2, h2a, h2b);
                    f2_unpack(s0, s0a, s0b); f2_unpack(s1, s1a, s1b); f2_unpack(s2, s2a, s2b);
                    hm[u] = fmaxf(fabsf(h0a), fmaxf(fabsf(h1a), fabsf(h2a)));
                    hm[u + 1] = fmaxf(fabsf(h0b), fmaxf(fabsf(h1b), fabsf(h2b)));
                    idx[u] = min((unsigned)__float_as_int(s0a) * pnyz + ((unsigned)__float_as_int(s1a) * pnz + ((unsigned)__float_as_int(s2a) - bias)), last);
                    idx[u + 1] = min((unsigned)__float_as_int(s0b) * pnyz + ((unsigned)__float_as_int(s1b) * pnz + ((unsigned)__float_as_int(s2b) - bias)), last);
                }
#else
#pragma unroll
                for (int u = 0; u < CARVE_MLP; u++) {
                    const float kf = kfb + (float)u;                                     // exact: small integers
                    const float t0 = fmaf(kf, qb0, ta0), t1 = fmaf(kf, qb1, ta1), t2 = fmaf(kf, qb2, ta2);
                    const float s0 = __fadd_rn(t0, kM), s1 = __fadd_rn(t1, kM), s2 = __fadd_rn(t2, kM);
                    const float h0 = __fsub_rn(t0, __fsub_rn(s0, kM)), h1 = __fsub_rn(t1, __fsub_rn(s1, kM)), h2 = __fsub_rn(t2, __fsub_rn(s2, kM));
                    hm[u] = fmaxf(fabsf(h0), fmaxf(fabsf(h1), fabsf(h2)));
                    idx[u] = min((unsigned)__float_as_int(s0) * pnyz + ((unsigned)__float_as_int(s1) * pnz + ((unsigned)__float_as_int(s2) - bias)), last);
                }
#endif
                const unsigned live = (j + CARVE_MLP <= n) ? ((1u << CARVE_MLP) - 1u) : ((1u << (n - j)) - 1u);   // samples of this group inside the chunk
                unsigned okm = 0u, bit[CARVE_MLP];
#pragma unroll
                for (int u = 0; u < CARVE_MLP; u++)
                    if (hm[u] <= hmax) okm |= 1u << u;
                unsafe |= (~okm & live) << j;
#pragma unroll
                for (int u = 0; u < CARVE_MLP; u++) word[u] = obs[idx[u] >> 5];         // (a clamped, in-grid address even when the sample is not used)
                // one branch per group, not per sample: after the first views of a sweep no bit is missing any more
                unsigned need = 0u;
#pragma unroll
                for (int u = 0; u < CARVE_MLP; u++) {
                    asm("shf.l.wrap.b32 %0, 0, 1, %1;" : "=r"(bit[u]) : "r"(idx[u]));   // 1u << (idx & 31)
                    if (!(word[u] & bit[u])) need |= 1u << u;
                }
                need &= okm & live;
                if (need) {
#pragma unroll
                    for (int u = 0; u < CARVE_MLP; u++)
                        if ((need >> u) & 1u) atomicOr(obs + (idx[u] >> 5), bit[u]);
                }
            }
            while (unsafe) {
                const int j = __ffs(unsafe) - 1;
                unsafe &= unsafe - 1u;
                carve_exact_sample<EXACT>(a, sp, ci, ri, kb + j);
            }
        }
    }
    for (int k = b1 + 1; k <= k_last; k++) carve_exact_sample<EXACT>(a, sp, ci, ri, k);      // exit band
}

#if DMF_CARVE_SIGN
// carve_on_line with fewer instructions on the half-rate integer pipe (ncu on the packed-FP32 build: ALU pipe 65 % busy, issue
// 72 %): the same samples located by the same arithmetic, but
//  * the face test "max|h| <= hmax" is taken from the SIGN of hmax - max|h| (the sign of a rounded float difference is the sign
//    of the exact one; equal operands give +0) and funnel-shifted into a per-group bit string: one SHF per sample instead of
//    compare + select + add.  The newest sample sits in bit 0, so bit b of a group is its sample b ^ (CARVE_MLP - 1);
//  * the missing observed bits are collected as words (bit & ~word, one LOP3) and OR-ed into one "anything missing?" test;
//  * groups are always full: a range whose length is not a multiple of CARVE_MLP starts with one group at b0 and continues
//    at b0 + (length % CARVE_MLP), so a few samples are located twice (marking is idempotent) and no per-sample "inside the
//    chunk" mask exists.  A safe range shorter than one group is evaluated exactly.
// (Branching on "a face test failed" per group does NOT pay: 1.2 % of the samples fail, i.e. some lane of nearly every warp --
// measured 10.3 ms against 9.5 ms.  The failures stay deferred to the end of the chunk as in carve_on_line.)
template <bool EXACT>
__device__ __forceinline__ void carve_on_line_sign(const FwdArgs& a, const float4* sp, int ci, int ri, int k_first, int k_last, int kin, int kout,
                                                   float qa0, float qa1, float qa2, float qb0, float qb1, float qb2, float esafe) {
    const VolDev& v = a.vol;
    const int b0 = max(kin, k_first), b1 = min(kout, k_last);
    const int len = b1 - b0 + 1;
    if (len < CARVE_MLP) {                                   // no (or a very short) safely-inside range: everything exactly
        for (int k = k_first; k <= k_last; k++) carve_exact_sample<EXACT>(a, sp, ci, ri, k);
        return;
    }
    for (int k = k_first; k < b0; k++) carve_exact_sample<EXACT>(a, sp, ci, ri, k);          // entry band around the boundary
    {
        const float kM = 12582912.0f;                                                        // 1.5 * 2^23
        unsigned pnz = (unsigned)v.pdim[2], pnyz = (unsigned)v.pdim[1] * pnz;
        unsigned bias = 0x4B400000u * (pnyz + pnz + 1u);                                     // the three "- 0x4B400000" of the shifter, folded
#if DMF_CARVE_CLAMP
        unsigned last = pnyz * (unsigned)v.pdim[0] - 1u;
        asm volatile("" : "+r"(last));
#endif
        asm volatile("" : "+r"(pnz), "+r"(pnyz), "+r"(bias));                                // registers, not re-derived from the constant bank per sample
        const float ta0 = qa0 - 0.5f, ta1 = qa1 - 0.5f, ta2 = qa2 - 0.5f;
        const float hmax = 0.5f - esafe;
        const f32x2 qb0p = f2_pack(qb0, qb0), qb1p = f2_pack(qb1, qb1), qb2p = f2_pack(qb2, qb2);
        const f32x2 ta0p = f2_pack(ta0, ta0), ta1p = f2_pack(ta1, ta1), ta2p = f2_pack(ta2, ta2), kMp = f2_pack(kM, kM);
        unsigned* const obs = a.observed;
        const int head = len % CARVE_MLP;                    // != 0: one group at b0 first, the aligned rest starts at b0 + head
        int kb = b0, n = head ? CARVE_MLP : min(32, len), next = head ? b0 + head : b0 + n;
        for (;;) {
            unsigned unsafe = 0u;                            // bit j + b: sample kb + j + (b ^ (CARVE_MLP - 1)) failed the face test
            for (int j = 0; j < n; j += CARVE_MLP) {         // n is a multiple of CARVE_MLP
                unsigned idx[CARVE_MLP], word[CARVE_MLP], miss[CARVE_MLP];
                unsigned fail = 0u;
                const float kfb = (float)(kb + j);
                const f32x2 kp = f2_pack(kfb, kfb + 1.0f);
#pragma unroll
                for (int u = 0; u < CARVE_MLP; u += 2) {
                    const f32x2 kf = f2_add(kp, f2_pack((float)u, (float)u));            // exact: small integers
                    const f32x2 t0 = f2_fma(kf, qb0p, ta0p), t1 = f2_fma(kf, qb1p, ta1p), t2 = f2_fma(kf, qb2p, ta2p);
                    const f32x2 s0 = f2_add(t0, kMp), s1 = f2_add(t1, kMp), s2 = f2_add(t2, kMp);
                    const f32x2 h0 = f2_sub(t0, f2_sub(s0, kMp)), h1 = f2_sub(t1, f2_sub(s1, kMp)), h2 = f2_sub(t2, f2_sub(s2, kMp));
                    float h0a, h0b, h1a, h1b, h2a, h2b, s0a, s0b, s1a, s1b, s2a, s2b;
                    f2_unpack(h0, h0a, h0b); f2_unpack(h1, h1a, h1b); f2_unpack(h2, h2a, h2b);
                    f2_unpack(s0, s0a, s0b); f2_unpack(s1, s1a, s1b); f2_unpack(s2, s2a, s2b);
                    const float da = __fsub_rn(hmax, fmaxf(fabsf(h0a), fmaxf(fabsf(h1a), fabsf(h2a))));   // < 0  <=>  face test failed
                    const float db = __fsub_rn(hmax, fmaxf(fabsf(h0b), fmaxf(fabsf(h1b), fabsf(h2b))));
                    fail = __funnelshift_l(__float_as_uint(da), fail, 1);                // (fail << 1) | sign(da)
                    fail = __funnelshift_l(__float_as_uint(db), fail, 1);
                    idx[u] = (unsigned)__float_as_int(s0a) * pnyz + ((unsigned)__float_as_int(s1a) * pnz + ((unsigned)__float_as_int(s2a) - bias));
                    idx[u + 1] = (unsigned)__float_as_int(s0b) * pnyz + ((unsigned)__float_as_int(s1b) * pnz + ((unsigned)__float_as_int(s2b) - bias));
#if DMF_CARVE_CLAMP
                    idx[u] = min(idx[u], last); idx[u + 1] = min(idx[u + 1], last);      // a seat belt, never active: the samples are inside
#endif
                    DMF_CHECK_IDX(idx[u], v.n_cells, a.counters); DMF_CHECK_IDX(idx[u + 1], v.n_cells, a.counters);
                }
                unsafe |= fail << j;
#pragma unroll
                for (int u = 0; u < CARVE_MLP; u++) word[u] = obs[idx[u] >> 5];
                unsigned missing = 0u;
#pragma unroll
                for (int u = 0; u < CARVE_MLP; u++) {
                    unsigned bit;
                    asm("shf.l.wrap.b32 %0, 0, 1, %1;" : "=r"(bit) : "r"(idx[u]));      // 1u << (idx & 31)
                    miss[u] = bit & ~word[u];                                            // the bit if it is not set yet, else 0
                    missing |= miss[u];
                }
                if (missing) {                                                           // rare once the sweep has seen the voxels
#pragma unroll
                    for (int u = 0; u < CARVE_MLP; u++)
                        if (miss[u] && !((fail >> (CARVE_MLP - 1 - u)) & 1u)) atomicOr(obs + (idx[u] >> 5), miss[u]);
                }
            }
            while (unsafe) {
                const int b = __ffs(unsafe) - 1;
                unsafe &= unsafe - 1u;
                carve_exact_sample<EXACT>(a, sp, ci, ri, kb + (b ^ (CARVE_MLP - 1)));
            }
            kb = next;
            if (kb > b1) break;
            n = min(32, b1 - kb + 1); next = kb + n;
        }
    }
    for (int k = b1 + 1; k <= k_last; k++) carve_exact_sample<EXACT>(a, sp, ci, ri, k);      // exit band
}
#endif

template <int MODE, bool EXACT, bool CARVE>
__global__ void __launch_bounds__(SKIP_THREADS, CARVE ? CARVE_MIN_BLOCKS : LINE_MIN_BLOCKS) k_forward_line(const FwdArgs a) {
    // No shared memory, no barrier: everything a block needs per view comes from the view's 64-byte record (k_view_start), read
    // with four 16-byte loads that hit L1; "samples advanced after a probe with byte d" is computed per probe from the RAY's own
    // slope (5 FMA-pipe instructions) instead of a per-block 256-entry table (round 1: table build + two barriers + a shared
    // atomicMax per ray, and an LDS in the dependent chain of every probe).
    const int view = blockIdx.z + a.view0;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int ci = blockIdx.x * SKIP_TILE_W + (warp & 1) * 8 + (lane & 7);
    int ri = blockIdx.y * SKIP_TILE_H + (warp >> 1) * 4 + (lane >> 3);
    const bool active = ci < a.Wc && ri < a.Hc;
    if (!active) { ci = 0; ri = 0; }
#ifdef DMF_LINE_STATS
    const long long t_begin = clock64();
    long long t_line = 0, t_exact = 0;
#endif
    const float4* sp = a.viewrec + 4u * (unsigned)view;
    const VolDev& v = a.vol;
    const float kM = 12582912.0f;
    const int S = a.S;
    // ---- per ray: the line; per tile (k_tile_start): the sample intervals and the first sample the cone pre-march could not prove a miss ----
    const u64 rec = __ldg(a.tile_rec + ((unsigned)view * (unsigned)a.tiles_per_view + (unsigned)(ri / TILE_RAYS) * (unsigned)a.tiles_x + (unsigned)(ci / TILE_RAYS)));
    float qa0, qa1, qa2, qb0, qb1, qb2;
    float rq, c1;                                     // samples advanced after a probe with byte d >= 2: floor(d * rq + c1)
    int k = (int)((unsigned)rec & 0xFFFu), s_end = (int)(((unsigned)rec >> 12) & 0xFFFu), kin = (int)((unsigned)(rec >> 24) & 0xFFFu);
    int kout = (int)((unsigned)(rec >> 36) & 0xFFFu) - 1;
    const int kt = (int)((unsigned)(rec >> 48) & 0xFFFu);
    const int k_slab = k;                             // CARVE: first sample that can be in bounds (before the view-wide k0 skip)
    unsigned n_inb = 0, n_exact = 0, n_f64 = 0, n_skip = 0;
    {
        const float4 r0 = ldg_f4_volatile(sp), r1 = ldg_f4_volatile(sp + 1), r2 = ldg_f4_volatile(sp + 2);
        const float dcx = __ldg(a.dcx + ci), dcy = __ldg(a.dcy + ri);
        const float g0 = fmaf(r0.x, dcx, fmaf(r0.y, dcy, r0.z)), g1 = fmaf(r1.x, dcx, fmaf(r1.y, dcy, r1.z)), g2 = fmaf(r2.x, dcx, fmaf(r2.y, dcy, r2.z));
        const float z0m = a.z0m, zdm = a.zdm;
        qa0 = fmaf(fmaf(z0m, g0, r0.w), v.inv32[0], v.c32[0]); qa1 = fmaf(fmaf(z0m, g1, r1.w), v.inv32[1], v.c32[1]); qa2 = fmaf(fmaf(z0m, g2, r2.w), v.inv32[2], v.c32[2]);
        qb0 = zdm * g0 * v.inv32[0]; qb1 = zdm * g1 * v.inv32[1]; qb2 = zdm * g2 * v.inv32[2];
        const float qbmax = fmaxf(fabsf(qb0), fmaxf(fabsf(qb1), fabsf(qb2)));
        rq = rcp_approx(fmaxf(qbmax, 1e-3f)) * 0.999999f;                           // <= 1000; rounded down: never over-skips
        c1 = fmaf(-1.25f, rq, 1.0f);
        const int ks = __float_as_int(ldg_f4_volatile(sp + 3).y);                    // -1: no skipping for this view; else leading probes no ray can hit
        if (ks >= 0) {
            const int k0 = min(ks, S);
            k = max(k, k0);
            n_inb = n_skip = (unsigned)k0;                                            // k0 > 0 only when the camera sits inside the volume
        }
    }
    if (!active) { k = s_end = S; n_inb = n_skip = 0u; }      // (lanes outside a ragged lattice must not count the view-wide k0 either)
    // opaque register copies: otherwise the compiler re-derives the pixel from %tid / %ctaid on every exact evaluation
    asm volatile("" : "+r"(ci), "+r"(ri));

    int hit_k = -1, hx = 0, hy = 0, hz = 0;
    float hpx = 0.f, hpy = 0.f, hpz = 0.f;
    const unsigned char* __restrict__ gbytes = v.bytes;
    const unsigned pny = (unsigned)v.pdim[1], pnz = (unsigned)v.pdim[2];
    const unsigned pnyz = a.pnyz, bias = a.bias, last = a.last;                      // folded on the host (FwdArgs)
    unsigned iter = 0;
    bool stop = false;
#ifdef DMF_LINE_STATS
    const long long t_loop0 = clock64();
#endif
    while (k < s_end && !stop) {
        if (k >= kin && k <= kout) {
            // ---- follow the line ----
#ifdef DMF_LINE_STATS
            const long long tl0 = clock64();
#endif
            // k .. kout are in bounds (slab test); those before kt are misses (cone pre-march of the tile): the first probe goes to kt,
            // or to kout if the tile's march got even further (one probe that could be saved, but no second loop shape)
            float kf = (float)max(k, min(kt, kout));
            const float koutf = (float)kout;
            for (;;) {
                if (MODE == 4 && (iter++ & 7u) == 0u) {   // rayTraceAndGetMinimum: planes behind the current minimum cannot matter
                    const int cur = *((volatile int*)(a.min_depth + view));
                    if (a.z0 + (int)kf * a.zdelta > cur) { stop = true; break; }
                }
                const unsigned bx = (unsigned)__float_as_int(__fadd_rd(fmaf(kf, qb0, qa0), kM));
                const unsigned by = (unsigned)__float_as_int(__fadd_rd(fmaf(kf, qb1, qa1), kM));
                const unsigned bz = (unsigned)__float_as_int(__fadd_rd(fmaf(kf, qb2, qa2), kM));
#ifdef DMF_CHECKED
                unsigned lidx = bx * pnyz + (by * pnz + (bz - bias));
                DMF_CHECK_IDX(lidx, v.n_cells, a.counters);
                const unsigned d = __ldg(gbytes + lidx);
#else
                const unsigned d = __ldg(gbytes + min(bx * pnyz + (by * pnz + (bz - bias)), last));   // (the clamp is a seat belt, never active)
#endif
#ifdef DMF_LINE_STATS
                n_f64++;                                   // diagnostic build: F64_PATH counts line probes, EXACT_DIV exact ones
#endif
                // this probe + the skipped ones: floor((d - 1.25) / max|QB|) + 1 for d >= 2; 2^20 when the probe must be evaluated exactly
                float adv = d >= 2u ? __fadd_rd(fmaf(__int_as_float(0x4B000000 | (int)d) - 8388608.0f, rq, c1), kM) - kM : LINE_EXACT_FLAG;
                if (d == 1u) {
                    // Next to an occupied voxel, but this voxel itself is empty.  If the line point is
                    // at least e_safe (> eps_q) away from every face of its voxel, the reference's sample is in the same voxel:
                    // an in-bounds miss (the line is >= 0.25 voxel inside the volume here).  Otherwise evaluate exactly.
                    // (Resolving d == 0 here as well -- "the sample is in this occupied voxel: a hit, work out its exact position after the
                    // march" -- was measured 4 % SLOWER: ~1.2 % of the face tests fail, i.e. some lane of most warps still takes the exact
                    // path, so the warp pays for both.)
                    const float q0 = fmaf(kf, qb0, qa0), q1 = fmaf(kf, qb1, qa1), q2 = fmaf(kf, qb2, qa2);
                    const float f0 = q0 - (__fadd_rd(q0, kM) - kM), f1 = q1 - (__fadd_rd(q1, kM) - kM), f2 = q2 - (__fadd_rd(q2, kM) - kM);
                    const float e = ldg_f4_volatile(sp + 3).x;     // e_safe = eps_q of the view + 2^-10 voxel of slack
                    if (fminf(f0, fminf(f1, f2)) >= e && fmaxf(f0, fmaxf(f1, f2)) <= 1.0f - e) adv = 1.0f;
                }
                kf += adv;
                if (!(kf <= koutf)) break;
            }
#ifdef DMF_LINE_STATS
            t_line += clock64() - tl0;
#endif
            const bool need_exact = kf >= LINE_EXACT_FLAG;
            if (need_exact) kf -= LINE_EXACT_FLAG;
            // the bytes know nothing of the volume's boundary: samples up to kout are in bounds by the slab test, later ones are not proven
            const int k2 = min((int)kf, kout + 1);
            n_inb += (unsigned)(k2 - k); n_skip += (unsigned)(k2 - k);
            k = k2;
            if (!need_exact) continue;
        }
        // ---- exact evaluation of sample k (identical to k_forward) ----
        if (MODE == 4 && (iter++ & 7u) == 0u) {
            const int cur = *((volatile int*)(a.min_depth + view));
            if (a.z0 + k * a.zdelta > cur) break;
        }
#ifdef DMF_LINE_STATS
        n_exact++;
        const long long te0 = clock64();
#endif
        const float xf = __ldg(a.xtab + ((unsigned)k * (unsigned)a.Wc + (unsigned)ci)), yf = __ldg(a.ytab + ((unsigned)k * (unsigned)a.Hc + (unsigned)ri));
        const float zf = __ldg(a.ztab + k);
        const float4 r0 = ldg_f4_volatile(sp), r1 = ldg_f4_volatile(sp + 1), r2 = ldg_f4_volatile(sp + 2);
        const float px = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(r0.x, xf), __fmul_rn(r0.y, yf)), __fmul_rn(r0.z, zf)), r0.w);
        const float py = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(r1.x, xf), __fmul_rn(r1.y, yf)), __fmul_rn(r1.z, zf)), r1.w);
        const float pz = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(r2.x, xf), __fmul_rn(r2.y, yf)), __fmul_rn(r2.z, zf)), r2.w);
        if (!(px > v.lo[0] && px < v.hi[0] && py > v.lo[1] && py < v.hi[1] && pz > v.lo[2] && pz < v.hi[2])) { k++; continue; }   // validPoints failed
        n_inb++;
        int ix, iy, iz;
        unsigned idx = probe_index<EXACT>(v, px, py, pz, v.inv32[0], v.inv32[1], v.inv32[2], v.c32[0], v.c32[1], v.c32[2],
                                          v.err32[0], v.err32[1], v.err32[2], pny, pnz, ix, iy, iz, n_f64, n_exact);
        DMF_CHECK_IDX(idx, v.n_cells, a.counters);
        const unsigned de = __ldg(gbytes + idx);
#ifdef DMF_LINE_STATS
        if (de != 77u) t_exact += clock64() - te0;        // (depends on the load: the wait is inside the bracket)
#endif
        if (de == 0u) { hit_k = k; hx = ix; hy = iy; hz = iz; hpx = px; hpy = py; hpz = pz; k++; break; }
        k++;
    }
#ifdef DMF_LINE_STATS
    const long long t_loop1 = clock64();
#endif
    if (CARVE && MODE != 4 && active) {
        const int k_last = hit_k >= 0 ? hit_k : s_end - 1;                          // samples >= s_end are provably outside the volume
#if DMF_CARVE_SIGN
        carve_on_line_sign<EXACT>(a, sp, ci, ri, k_slab, k_last, kin, kout, qa0, qa1, qa2, qb0, qb1, qb2, ldg_f4_volatile(sp + 3).x);
#else
        carve_on_line<EXACT>(a, sp, ci, ri, k_slab, k_last, kin, kout, qa0, qa1, qa2, qb0, qb1, qb2, ldg_f4_volatile(sp + 3).x);
#endif
    }
    const unsigned n_samples = active ? (unsigned)((hit_k >= 0 || stop || (MODE == 4 && k < s_end)) ? min(k, S) : S) : 0u;
    float t0 = 0.f, t1 = 0.f, t2 = 0.f;
    if (MODE == 1 || MODE == 2) { const float* P = a.poses + 12u * (unsigned)view; t0 = __ldg(P + 3); t1 = __ldg(P + 7); t2 = __ldg(P + 11); }
    forward_epilogue<MODE>(a, nullptr, view, ci, ri, active, hit_k, hx, hy, hz, hpx, hpy, hpz, t0, t1, t2, n_samples, n_inb, n_exact, n_f64, n_skip);
    if (MODE != 4 && a.pub.enabled) publish_view_row(a.pub, reinterpret_cast<u64*>(a.vis + (size_t)(unsigned)view * a.vis_stride32), view, gridDim.x * gridDim.y);
#ifdef DMF_LINE_STATS
    if (lane == 0) {
        unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        u64* slot = a.counters + (size_t)smid * DMF_COUNTER_STRIDE;
        const unsigned long long dt = (unsigned long long)(clock64() - t_begin);
        atomicAdd(slot + 13, dt); atomicAdd(slot + 14, (unsigned long long)t_line); atomicAdd(slot + 15, (unsigned long long)t_exact);
        atomicAdd(slot + 11, (unsigned long long)(t_loop0 - t_begin)); atomicAdd(slot + 10, (unsigned long long)(t_loop1 - t_loop0));   // prologue, whole march loop
        if (warp == 0) atomicAdd(slot + 12, 1ull);
    }
#endif
}

// CLASSIFY: `if(voxel->view==0) voxel->view=view` over a batch of views in call order (:354-355)
__global__ void k_apply_first_view(int* view_mark, int* first_view, int n_occ, int view_id0) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_occ) return;
    int f = first_view[i];
    if (f != 0x7fffffff) {
        if (view_mark[i] == 0) view_mark[i] = view_id0 + f;
        first_view[i] = 0x7fffffff;
    }
}

__global__ void k_fill_u32(unsigned* p, size_t n, unsigned val) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = val;
}

__global__ void k_finish_min_depth(int* min_depth, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && min_depth[i] == 0x7fffffff) min_depth[i] = -1;
}

// ---- discovery order -------------------------------------------------------------------------------------
// The reference appends an id the first time a (z_depth, r, c)-ordered scan emits it (:484-488, :432-436).
// A ray "wins" its voxel if its key (k<<21 | lattice index) equals first_key[occ]; the returned list is the
// winners sorted by key.  Keys are unique and the compaction below keeps lattice (r,c) order, so a stable
// 2-pass (5+5 bit) LSD radix sort on k alone suffices (k_ord_*).

// Stage 1 (multi-block, fully parallel over the lattice): stable compaction of the winners' keys in lattice order.
//   k_win_count   blk_cnt[view][b] = winners among lattice entries [b*WIN_BLOCK, (b+1)*WIN_BLOCK)
//   k_win_offsets per view: exclusive scan of blk_cnt -> blk_off, total -> n_win[view]
//   k_win_compact tmp[view][blk_off + rank inside the block] = key
constexpr int WIN_THREADS = 256, WIN_ITEMS = 2, WIN_BLOCK = WIN_THREADS * WIN_ITEMS;      // (8 items per thread left a single view with 150 blocks on 148 SMs: 13 us of dependent gathers)

__device__ __forceinline__ bool is_winner(const unsigned* rk, const int* ro, const unsigned* fk, int i, int R, unsigned& key) {
    if (i >= R) return false;
    key = rk[i];
    return key != 0xFFFFFFFFu && fk[ro[i]] == key;
}

__global__ void __launch_bounds__(WIN_THREADS) k_win_count(const unsigned* ray_key, const int* ray_occ, const unsigned* first_key,
                                                           unsigned* blk_cnt, int R, int n_occ, int nb) {
    __shared__ unsigned s_warp[WIN_THREADS / 32];
    const int view = blockIdx.y;
    const unsigned* rk = ray_key + (size_t)view * R; const int* ro = ray_occ + (size_t)view * R; const unsigned* fk = first_key + (size_t)view * n_occ;
    unsigned cnt = 0, key;
    for (int j = 0; j < WIN_ITEMS; j++) cnt += is_winner(rk, ro, fk, blockIdx.x * WIN_BLOCK + j * WIN_THREADS + threadIdx.x, R, key) ? 1u : 0u;
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) { unsigned t = 0; for (int w = 0; w < WIN_THREADS / 32; w++) t += s_warp[w]; blk_cnt[(size_t)view * nb + blockIdx.x] = t; }
}

__global__ void __launch_bounds__(1024) k_win_offsets(const unsigned* blk_cnt, unsigned* blk_off, int* n_win, int nb) {
    __shared__ unsigned s_warp[32];
    __shared__ unsigned s_carry;
    const int view = blockIdx.x;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < nb; base += 1024) {
        const int i = base + threadIdx.x;
        const unsigned val = i < nb ? blk_cnt[(size_t)view * nb + i] : 0u;
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        unsigned x = val;
        for (int o = 1; o < 32; o <<= 1) { unsigned y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        if (lane == 31) s_warp[warp] = x;
        __syncthreads();
        if (warp == 0) { unsigned w = s_warp[lane]; for (int o = 1; o < 32; o <<= 1) { unsigned y = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += y; } s_warp[lane] = w; }
        __syncthreads();
        const unsigned excl = s_carry + (warp ? s_warp[warp - 1] : 0u) + x - val;
        if (i < nb) blk_off[(size_t)view * nb + i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = excl + val;
        __syncthreads();
    }
    if (threadIdx.x == 0) n_win[view] = (int)s_carry;
}

// blk_off == nullptr: every block derives its offset from the raw counts blk_cnt itself (the sum over the blocks before it) and
// the last block writes the view's total to n_win -- one launch less than k_win_offsets + this (single-view calls)
__global__ void __launch_bounds__(WIN_THREADS) k_win_compact(const unsigned* ray_key, const int* ray_occ, const unsigned* first_key,
                                                             const unsigned* blk_off, unsigned* tmp, int R, int n_occ, int nb,
                                                             const unsigned* blk_cnt = nullptr, int* n_win = nullptr) {
    __shared__ unsigned s_warp[WIN_THREADS / 32];
    __shared__ unsigned s_base0;
    const int view = blockIdx.y;
    const unsigned* rk = ray_key + (size_t)view * R; const int* ro = ray_occ + (size_t)view * R; const unsigned* fk = first_key + (size_t)view * n_occ;
    unsigned* out = tmp + (size_t)view * R;
    if (!blk_off) {
        unsigned bef = 0;
        for (int b = threadIdx.x; b < (int)blockIdx.x; b += WIN_THREADS) bef += blk_cnt[(size_t)view * nb + b];
        bef = __reduce_add_sync(0xffffffffu, bef);
        if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = bef;
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned t = 0;
            for (int q = 0; q < WIN_THREADS / 32; q++) t += s_warp[q];
            s_base0 = t;
            if ((int)blockIdx.x == nb - 1) n_win[view] = (int)(t + blk_cnt[(size_t)view * nb + blockIdx.x]);
        }
        __syncthreads();
    }
    unsigned base = blk_off ? blk_off[(size_t)view * nb + blockIdx.x] : s_base0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int j = 0; j < WIN_ITEMS; j++) {       // rows of WIN_THREADS consecutive lattice entries keep the lattice order
        unsigned key = 0;
        const bool w = is_winner(rk, ro, fk, blockIdx.x * WIN_BLOCK + j * WIN_THREADS + threadIdx.x, R, key);
        const unsigned ballot = __ballot_sync(0xffffffffu, w);
        if (lane == 0) s_warp[warp] = __popc(ballot);
        __syncthreads();
        unsigned before = 0, total = 0;
        for (int q = 0; q < WIN_THREADS / 32; q++) { const unsigned c = s_warp[q]; if (q < warp) before += c; total += c; }
        if (w) out[base + before + __popc(ballot & ((1u << lane) - 1u))] = key;
        base += total;
        __syncthreads();
    }
}

// Stage 2, multi-block (round 1 sorted each view in ONE block: 110-195 us for the ~24 k winners of one VGA view, the longest
// kernel of a single-view call): a stable LSD radix sort of the compacted keys on their z-plane bits, two
// passes of 5 bits (keys are unique and arrive in lattice order, so stability on k alone gives (k, r, c) order).  Per pass:
//   k_ord_hist    per block of ORD_TILE keys: histogram of the 32 digit values           -> hist[view][digit][block]
//   k_ord_scan    per view: exclusive scan of hist in (digit-major, block-minor) order   -> same array, in place
//   k_ord_scatter per block: stable rank of every key among equal digits of the block (warp match + per-warp counts), scatter
constexpr int ORD_TILE = 1024;                    // keys per block = threads per block
__global__ void __launch_bounds__(ORD_TILE) k_ord_hist(const unsigned* __restrict__ keys, const int* __restrict__ n_win, unsigned* __restrict__ hist, int R, int nblk, int shift) {
    __shared__ unsigned s_h[32];
    const int view = blockIdx.y, nw = n_win[view];
    if ((int)(blockIdx.x * ORD_TILE) >= nw) { if (threadIdx.x < 32) hist[((size_t)view * 32 + threadIdx.x) * nblk + blockIdx.x] = 0u; return; }
    if (threadIdx.x < 32) s_h[threadIdx.x] = 0u;
    __syncthreads();
    const int i = blockIdx.x * ORD_TILE + threadIdx.x;
    const unsigned digit = i < nw ? (keys[(size_t)view * R + i] >> shift) & 31u : 32u;
    const unsigned peers = __match_any_sync(0xffffffffu, digit);
    if (digit < 32u && (threadIdx.x & 31) == (unsigned)(__ffs(peers) - 1)) atomicAdd(&s_h[digit], (unsigned)__popc(peers));
    __syncthreads();
    if (threadIdx.x < 32) hist[((size_t)view * 32 + threadIdx.x) * nblk + blockIdx.x] = s_h[threadIdx.x];
}
__global__ void __launch_bounds__(1024) k_ord_scan(unsigned* __restrict__ hist, int nblk) {
    __shared__ unsigned s_warp[32];
    __shared__ unsigned s_carry;
    unsigned* h = hist + (size_t)blockIdx.x * 32 * nblk;
    const int n = 32 * nblk;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
        const int i = base + threadIdx.x;
        const unsigned val = i < n ? h[i] : 0u;
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        unsigned x = val;
        for (int o = 1; o < 32; o <<= 1) { unsigned y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        if (lane == 31) s_warp[warp] = x;
        __syncthreads();
        if (warp == 0) { unsigned w = s_warp[lane]; for (int o = 1; o < 32; o <<= 1) { unsigned y = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += y; } s_warp[lane] = w; }
        __syncthreads();
        const unsigned excl = s_carry + (warp ? s_warp[warp - 1] : 0u) + x - val;
        if (i < n) h[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = excl + val;
        __syncthreads();
    }
}
// FINAL: instead of the sorted keys write the winners' occupied ordinals (ray_occ of the lattice index in the key's low 21 bits)
// FOLD: hist holds the RAW block histograms and every block derives its own 32 offsets from them (warp d sums digit d's row:
// all blocks for the digit's total, the blocks before this one for its share) -- for the few hundred blocks of a VGA view this is
// cheaper than a separate one-block scan kernel and the launch gap around it (single-view calls are bound by launches, not work)
template <bool FINAL, bool FOLD>
__global__ void __launch_bounds__(ORD_TILE) k_ord_scatter(const unsigned* __restrict__ keys, const int* __restrict__ n_win, const unsigned* __restrict__ hist,
                                                          unsigned* __restrict__ dst, const int* __restrict__ ray_occ, int* __restrict__ out_occ, int R, int nblk, int shift) {
    __shared__ unsigned s_cnt[32][33];            // [warp][digit] -> exclusive prefix over the warps of this block
    __shared__ unsigned s_tot[32], s_base[32];
    const int view = blockIdx.y, nw = n_win[view];
    if ((int)(blockIdx.x * ORD_TILE) >= nw) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (FOLD) {
        const unsigned* row = hist + ((size_t)view * 32 + warp) * nblk;          // warp d <-> digit d
        unsigned tot = 0, bef = 0;
        for (int b = lane; b < nblk; b += 32) { const unsigned h = row[b]; tot += h; if (b < (int)blockIdx.x) bef += h; }
        tot = __reduce_add_sync(0xffffffffu, tot); bef = __reduce_add_sync(0xffffffffu, bef);
        if (lane == 0) { s_tot[warp] = tot; s_base[warp] = bef; }
        __syncthreads();
        if (warp == 0) {
            const unsigned t = s_tot[lane];
            unsigned x = t;
            for (int o = 1; o < 32; o <<= 1) { unsigned y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
            s_base[lane] += x - t;                                                  // digits before this one + this digit's earlier blocks
        }
        __syncthreads();
    }
    const int i = blockIdx.x * ORD_TILE + threadIdx.x;
    const unsigned key = i < nw ? keys[(size_t)view * R + i] : 0u;
    const unsigned digit = i < nw ? (key >> shift) & 31u : 32u;
    const unsigned peers = __match_any_sync(0xffffffffu, digit);
    const unsigned rank_in_warp = (unsigned)__popc(peers & ((1u << lane) - 1u));
    s_cnt[warp][lane] = 0u;
    __syncwarp();
    if (digit < 32u && lane == __ffs(peers) - 1) s_cnt[warp][digit] = (unsigned)__popc(peers);
    __syncthreads();
    {   // thread (d = warp, w = lane): exclusive scan over the warps w of digit d's counts
        const unsigned c = s_cnt[lane][warp];
        unsigned x = c;
        for (int o = 1; o < 32; o <<= 1) { unsigned y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        __syncthreads();
        s_cnt[lane][warp] = x - c;
    }
    __syncthreads();
    if (digit < 32u) {
        const unsigned pos = (FOLD ? s_base[digit] : hist[((size_t)view * 32 + digit) * nblk + blockIdx.x]) + s_cnt[warp][digit] + rank_in_warp;
        if (FINAL) out_occ[(size_t)view * R + pos] = ray_occ[(size_t)view * R + (key & 0x1FFFFFu)];
        else dst[(size_t)view * R + pos] = key;
    }
}

// compact per-view winner lists into one contiguous uint64 id array, view after view (entries beyond `cap` are dropped:
// the host sees the total in offsets[n_views] and reports the overflow)
// n_ids != nullptr: the offsets are not there yet -- every block sums the counts of the views before its own (a chunk has at most a
// few thousand views) and block x == 0 of each view writes them for the host (k_ids_offsets folded in: one launch less)
__global__ void k_gather_ids(const int* out_occ, long long* offsets, const u64* occ_ids, u64* ids, int R, long long cap, const int* n_ids = nullptr) {
    const int view = blockIdx.y;
    __shared__ long long s_b;
    if (n_ids) {
        long long acc = 0;
        for (int v = threadIdx.x; v < view; v += blockDim.x) acc += n_ids[v];
        for (int o = 16; o; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
        if (threadIdx.x == 0) s_b = 0;
        __syncthreads();
        if ((threadIdx.x & 31) == 0 && acc) atomicAdd((unsigned long long*)&s_b, (unsigned long long)acc);
        __syncthreads();
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            offsets[view] = s_b;
            if (view == (int)gridDim.y - 1) offsets[view + 1] = s_b + n_ids[view];
        }
    }
    const long long b = n_ids ? s_b : offsets[view], n = n_ids ? (long long)n_ids[view] : offsets[view + 1] - b;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        if (b + i < cap) ids[b + i] = occ_ids[out_occ[(size_t)view * R + i]];
}
