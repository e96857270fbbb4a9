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
    // projectPoint tables (Camera.hpp:24-31), built once per (K, H, W, z0, zdelta, stride) by k_build_tables:
    const float* __restrict__ xtab;    // [S][Wc]  (float)(z_k*((double)c-cx)/fx)
    const float* __restrict__ ytab;    // [S][Hc]  (float)(z_k*((double)r-cy)/fy)
    const float* __restrict__ ztab;    // [S]      (float)z_k,  z_k = (z0 + k*zdelta)*0.001
    int S, Wc, Hc, W, H, cstride, rstride, z0, zdelta;
    // outputs (may be null)
    int* depth;                        // [n_views][H][W]
    float* points;                     // [n_views][H][W][3]
    u64* hit_voxel;                    // [n_views][H][W]
    unsigned* vis;                     // [n_views][vis_words32] bitset over occupied order
    int vis_words32;
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
    u64* counters;                     // DMF_CNT_*
};

// xtab/ytab/ztab: exact IEEE double mul/div then round to float, as the tuple<float,float,float> return does.
__global__ void k_build_tables(float* xtab, float* ytab, float* ztab, int S, int Wc, int Hc, int cstride, int rstride,
                               int z0, int zdelta, double fx, double cx, double fy, double cy) {
    int per = Wc + Hc + 1;
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

constexpr int FWD_THREADS = 256;      // 8 warps: 4 across x 2 down, each warp an 8x4 pixel tile
constexpr int FWD_TILE_W = 32, FWD_TILE_H = 8;
constexpr int FWD_CHUNK = 32;         // z-steps staged in shared memory per round

// MODE: DMF_MODE_*, FMT: DMF_GRID_*
template <int MODE, int FMT>
__global__ void __launch_bounds__(FWD_THREADS) k_forward(const FwdArgs a) {
    __shared__ float sx[FWD_CHUNK][FWD_TILE_W];
    __shared__ float sy[FWD_CHUNK][FWD_TILE_H];
    __shared__ float sz[FWD_CHUNK][4];   // m02*z, m12*z, m22*z
    __shared__ unsigned long long s_cnt[4];

    const int view = blockIdx.z;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int lc = (warp & 3) * 8 + (lane & 7);   // column within the tile
    const int lr = (warp >> 2) * 4 + (lane >> 3); // row within the tile
    const int ci = blockIdx.x * FWD_TILE_W + lc;  // lattice column
    const int ri = blockIdx.y * FWD_TILE_H + lr;  // lattice row
    const bool active = ci < a.Wc && ri < a.Hc;
    if (threadIdx.x < 4) s_cnt[threadIdx.x] = 0ull;

    const float* P = a.poses + 12 * (size_t)view;
    const float m00 = __ldg(P + 0), m01 = __ldg(P + 1), m02 = __ldg(P + 2), m03 = __ldg(P + 3);
    const float m10 = __ldg(P + 4), m11 = __ldg(P + 5), m12 = __ldg(P + 6), m13 = __ldg(P + 7);
    const float m20 = __ldg(P + 8), m21 = __ldg(P + 9), m22 = __ldg(P + 10), m23 = __ldg(P + 11);
    const VolDev& v = a.vol;

    bool done = !active;
    int hit_k = -1, hx = 0, hy = 0, hz = 0;
    float hpx = 0.f, hpy = 0.f, hpz = 0.f;
    unsigned n_samples = 0, n_inb = 0, n_exact = 0;

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
            for (int kk = 0; kk < nk; kk++) {
                const float xf = sx[kk][lc], yf = sy[kk][lr];
                // transformPoints (Camera.hpp:39-45), rule E1
                const float px = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m00, xf), __fmul_rn(m01, yf)), sz[kk][0]), m03);
                const float py = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m10, xf), __fmul_rn(m11, yf)), sz[kk][1]), m13);
                const float pz = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m20, xf), __fmul_rn(m21, yf)), sz[kk][2]), m23);
                n_samples++;
                if (!in_bounds(v, px, py, pz)) continue;          // validPoints: skip, do not terminate
                n_inb++;
                const int ix = voxel_index(px, v.vmin[0], v.delta[0], v.inv[0], v.c0[0], v.eps[0], n_exact);
                const int iy = voxel_index(py, v.vmin[1], v.delta[1], v.inv[1], v.c0[1], v.eps[1], n_exact);
                const int iz = voxel_index(pz, v.vmin[2], v.delta[2], v.inv[2], v.c0[2], v.eps[2], n_exact);
                if (!coords_valid(v, ix, iy, iz)) continue;       // the reference reads out of bounds here (UB): empty
                if (occupied<FMT>(v, ix, iy, iz)) {
                    hit_k = k0 + kk; hx = ix; hy = iy; hz = iz; hpx = px; hpy = py; hpz = pz;
                    done = true;
                    break;
                }
            }
        }
    }

    // ---- per-ray epilogue ------------------------------------------------------------------------------
    const bool hit = hit_k >= 0;
    const int z_depth = a.z0 + hit_k * a.zdelta;
    if (active) {
        const size_t pix = ((size_t)view * a.H + (size_t)ri * a.rstride) * a.W + (size_t)ci * a.cstride;
        if (a.depth) a.depth[pix] = hit ? z_depth : -1;
        if (a.hit_voxel) a.hit_voxel[pix] = hit ? voxel_id(hx, hy, hz) : ~0ull;
        if (a.points) {
            a.points[3 * pix] = hit ? hpx : 0.f; a.points[3 * pix + 1] = hit ? hpy : 0.f; a.points[3 * pix + 2] = hit ? hpz : 0.f;
        }
    }
    unsigned ties = 0;
    if (MODE == 4) {
        if (hit) atomicMin(a.min_depth + view, z_depth);
    } else {
        const size_t li = (size_t)view * a.Wc * a.Hc + (size_t)ri * a.Wc + ci;
        bool emit = false;
        int occ = -1;
        if (hit) {
            occ = occupied_ordinal(v, hx, hy, hz);
            if (a.found_any) a.found_any[view] = 1;
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
            if (emit && a.vis) {
                unsigned* w = a.vis + (size_t)view * a.vis_words32 + (occ >> 5);
                unsigned m = 1u << (occ & 31);
                if (!(*((volatile unsigned*)w) & m)) atomicOr(w, m);
            }
        }
        if (active && a.ray_key) {
            unsigned key = emit ? (((unsigned)hit_k << 21) | (unsigned)(ri * a.Wc + ci)) : 0xFFFFFFFFu;
            a.ray_key[li] = key;
            a.ray_occ[li] = occ;
            if (emit) atomicMin(a.first_key + (size_t)view * v.n_occ + occ, key);
        }
    }

    // ---- counters: warp reduce -> shared -> 4 global atomics per block ------------------------------
    unsigned long long c0 = n_samples, c1 = n_inb, c2 = hit ? 1u : 0u, c3 = n_exact, c5 = ties;
    for (int o = 16; o; o >>= 1) {
        c0 += __shfl_down_sync(0xffffffffu, c0, o); c1 += __shfl_down_sync(0xffffffffu, c1, o);
        c2 += __shfl_down_sync(0xffffffffu, c2, o); c3 += __shfl_down_sync(0xffffffffu, c3, o);
        c5 += __shfl_down_sync(0xffffffffu, c5, o);
    }
    if (lane == 0) {
        atomicAdd(&s_cnt[0], c0); atomicAdd(&s_cnt[1], c1); atomicAdd(&s_cnt[2], c2); atomicAdd(&s_cnt[3], c3);
        if (c5) atomicAdd(a.counters + 5, c5);
    }
    __syncthreads();
    if (threadIdx.x < 4 && s_cnt[threadIdx.x]) atomicAdd(a.counters + threadIdx.x, s_cnt[threadIdx.x]);
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
// winners sorted by key.  Keys are unique and the input is already in lattice (r,c) order, so a stable
// 2-pass (5+5 bit) LSD counting sort on k alone suffices.  One block per view; every thread owns a contiguous
// run of the input so stability needs no intra-block ranking.
constexpr int ORD_THREADS = 512;

__device__ __forceinline__ unsigned block_exclusive_scan_512(unsigned val, unsigned* s_warp, unsigned& total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned x = val;
    for (int o = 1; o < 32; o <<= 1) { unsigned y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    if (lane == 31) s_warp[warp] = x;
    __syncthreads();
    if (warp == 0) {
        unsigned w = (lane < ORD_THREADS / 32) ? s_warp[lane] : 0;
        for (int o = 1; o < 32; o <<= 1) { unsigned y = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += y; }
        s_warp[lane] = w;   // inclusive
    }
    __syncthreads();
    unsigned base = warp ? s_warp[warp - 1] : 0;
    total = s_warp[ORD_THREADS / 32 - 1];
    __syncthreads();
    return base + x - val;
}

// one stable counting-sort pass on digit = (key >> shift) & 31 over items src[0..n); PASS 0 filters winners.
template <int PASS>
__device__ void order_pass(const unsigned* src, const int* ray_occ, const unsigned* first_key, unsigned* dst, int n, int shift,
                           unsigned (*s_off)[ORD_THREADS], unsigned* s_warp, unsigned& n_out) {
    const int t = threadIdx.x;
    const int chunk = (n + ORD_THREADS - 1) / ORD_THREADS;
    const int b = min(n, t * chunk), e = min(n, b + chunk);
    for (int d = 0; d < 32; d++) s_off[d][t] = 0;
    for (int i = b; i < e; i++) {
        unsigned key = src[i];
        bool take = PASS ? true : (key != 0xFFFFFFFFu && first_key[ray_occ[i]] == key);
        if (take) s_off[(key >> shift) & 31][t]++;
    }
    __syncthreads();
    // exclusive scan over the (digit-major, thread-minor) linearisation: thread t owns entries [32t, 32t+32)
    unsigned* lin = &s_off[0][0];
    unsigned sum = 0;
    for (int j = 0; j < 32; j++) sum += lin[32 * t + j];
    unsigned total;
    unsigned base = block_exclusive_scan_512(sum, s_warp, total);
    for (int j = 0; j < 32; j++) { unsigned c = lin[32 * t + j]; lin[32 * t + j] = base; base += c; }
    __syncthreads();
    for (int i = b; i < e; i++) {
        unsigned key = src[i];
        bool take = PASS ? true : (key != 0xFFFFFFFFu && first_key[ray_occ[i]] == key);
        if (take) dst[s_off[(key >> shift) & 31][t]++] = key;
    }
    n_out = total;
    __syncthreads();
}

// grid = n_views.  tmp_a/tmp_b: [n_views][R] scratch.  out_occ: [n_views][R] winners' occupied ordinals in
// discovery order; n_ids[view] their count.
__global__ void __launch_bounds__(ORD_THREADS) k_order_ids(const unsigned* ray_key, const int* ray_occ, const unsigned* first_key,
                                                           unsigned* tmp_a, unsigned* tmp_b, int* out_occ, int* n_ids, int R, int n_occ) {
    extern __shared__ unsigned s_dyn[];
    unsigned (*s_off)[ORD_THREADS] = (unsigned (*)[ORD_THREADS])s_dyn;
    __shared__ unsigned s_warp[32];
    const int view = blockIdx.x;
    const unsigned* rk = ray_key + (size_t)view * R;
    const int* ro = ray_occ + (size_t)view * R;
    const unsigned* fk = first_key + (size_t)view * n_occ;
    unsigned* ta = tmp_a + (size_t)view * R;
    unsigned* tb = tmp_b + (size_t)view * R;
    unsigned nw = 0, nw2 = 0;
    order_pass<0>(rk, ro, fk, ta, R, 21, s_off, s_warp, nw);            // low 5 bits of k
    __threadfence_block();
    order_pass<1>(ta, nullptr, nullptr, tb, (int)nw, 26, s_off, s_warp, nw2);   // high 5 bits of k
    __threadfence_block();
    for (int i = threadIdx.x; i < (int)nw; i += ORD_THREADS) out_occ[(size_t)view * R + i] = ro[tb[i] & 0x1FFFFFu];
    if (threadIdx.x == 0) n_ids[view] = (int)nw;
}

// compact per-view winner lists into one contiguous uint64 id array at host-computed offsets
__global__ void k_gather_ids(const int* out_occ, const long long* offsets, const u64* occ_ids, u64* ids, int R) {
    const int view = blockIdx.y;
    const long long b = offsets[view], n = offsets[view + 1] - b;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        ids[b + i] = occ_ids[out_occ[(size_t)view * R + i]];
}
